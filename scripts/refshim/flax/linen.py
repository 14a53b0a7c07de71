"""Minimal eager ``flax.linen`` (see ../README.md): Module with dataclass fields, ``setup`` / ``@compact`` submodule
naming (attribute names, ``Name_i`` for lists, ``Class_N`` for inline construction), ``param``, ``init`` / ``apply``;
``Dense``, ``DenseGeneral``, ``LayerNorm``, ``MultiHeadDotProductAttention`` restated from the flax 0.12 API docs.
Arrays are torch float64 tensors.  Test infrastructure."""
from __future__ import annotations

import dataclasses
import functools
import math
from typing import Any

import torch

_stack: list = []          # modules whose method / setup is executing (innermost last)


def compact(fn):
    fn._compact = True
    return fn


class _Root:
    def __init__(self, params, initializing):
        self.params = params
        self.initializing = initializing


class _Initializers:
    @staticmethod
    def ones(key, shape, dtype=None):
        return torch.ones(tuple(shape), dtype=torch.float64)

    @staticmethod
    def zeros(key, shape, dtype=None):
        return torch.zeros(tuple(shape), dtype=torch.float64)

    @staticmethod
    def normal(stddev=1e-2, dtype=None):
        def init(key, shape, dtype=None):
            from jax import random

            return stddev * random.normal(key, tuple(shape))

        return init

    @staticmethod
    def constant(value, dtype=None):
        def init(key, shape, dtype=None):
            return torch.full(tuple(shape), float(value), dtype=torch.float64)

        return init

    @staticmethod
    def lecun_normal():
        def init(key, shape, dtype=None):
            from jax import random

            fan_in = 1
            for s in tuple(shape)[:-1]:
                fan_in *= s
            return random.normal(key, tuple(shape)) / math.sqrt(max(fan_in, 1))

        return init


initializers = _Initializers()


def tanh(x):
    return torch.tanh(x)


def softmax(x, axis=-1):
    return torch.softmax(x, dim=axis)


_RESERVED = {"setup", "init", "apply", "param", "is_initializing", "clone"}


def _wrap(fn):
    is_compact = getattr(fn, "_compact", False)

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        st = self.__dict__.get("_st")
        if st is None or (st["root"] is None and self.parent is None):
            return fn(self, *args, **kwargs)      # unbound module: plain method call (init_params, evaluate, ...)
        self._bind()
        self._run_setup()
        if is_compact:
            st["autonames"] = {}
        _stack.append(self)
        try:
            return fn(self, *args, **kwargs)
        finally:
            _stack.pop()

    return wrapper


class Module:
    def __init_subclass__(cls, kw_only=False, **kwargs):
        super().__init_subclass__(**kwargs)
        ann = dict(cls.__dict__.get("__annotations__", {}))
        ann.pop("parent", None)
        ann.pop("name", None)
        ann["parent"] = Any
        ann["name"] = Any
        cls.__annotations__ = ann
        cls.parent = dataclasses.field(default=None, kw_only=True, repr=False)
        cls.name = dataclasses.field(default=None, kw_only=True)
        for nm, attr in list(cls.__dict__.items()):
            if callable(attr) and isinstance(attr, type(_wrap)) and (not nm.startswith("__") or nm == "__call__") \
                    and nm not in _RESERVED:
                setattr(cls, nm, _wrap(attr))
        dataclasses.dataclass(cls, eq=False, repr=False, kw_only=bool(kw_only))

    # ---- construction ------------------------------------------------------------------------------------------
    def __post_init__(self):
        object.__setattr__(self, "_st", {"root": None, "scope": None, "setup_done": False, "in_setup": False,
                                         "autonames": {}})
        if self.parent is None and _stack:
            par = _stack[-1]
            object.__setattr__(self, "parent", par)
            if self.name is None and not par._st["in_setup"]:
                n = par._st["autonames"].get(type(self).__name__, 0)
                par._st["autonames"][type(self).__name__] = n + 1
                object.__setattr__(self, "name", f"{type(self).__name__}_{n}")

    def __setattr__(self, key, value):
        st = self.__dict__.get("_st")
        if st is not None and st["in_setup"]:
            def adopt(m, nm):
                if isinstance(m, Module) and m.name is None:
                    object.__setattr__(m, "parent", self)
                    object.__setattr__(m, "name", nm)

            adopt(value, key)
            if isinstance(value, (list, tuple)):
                for i, m in enumerate(value):
                    adopt(m, f"{key}_{i}")
        object.__setattr__(self, key, value)

    # ---- binding -----------------------------------------------------------------------------------------------
    def _bind(self):
        st = self._st
        if st["scope"] is not None:
            return
        par = self.parent
        par._bind()
        st["root"] = par._st["root"]
        if self.name is None:
            raise RuntimeError(f"refshim: submodule {type(self).__name__} used before it was named")
        sc = par._st["scope"]
        if st["root"].initializing:
            st["scope"] = sc.setdefault(self.name, {})
        else:
            st["scope"] = sc.get(self.name, {})   # a parameter-free submodule (features, LogDet) has no entry

    def _run_setup(self):
        st = self._st
        if st["setup_done"]:
            return
        st["setup_done"] = True
        if hasattr(self, "setup"):
            st["in_setup"] = True
            _stack.append(self)
            try:
                self.setup()
            finally:
                _stack.pop()
                st["in_setup"] = False

    def is_initializing(self):
        return bool(self._st["root"] and self._st["root"].initializing)

    def param(self, name, init_fn, *init_args):
        sc = self._st["scope"]
        if name in sc:
            return sc[name]
        if not self._st["root"].initializing:
            raise KeyError(f"refshim: parameter {name!r} missing in {type(self).__name__} ({self.name})")
        from jax import random

        sc[name] = init_fn(random.split(None, 1)[0], *init_args)
        return sc[name]

    def clone(self):
        vals = {f.name: getattr(self, f.name) for f in dataclasses.fields(self) if f.name not in ("parent", "name")}
        saved = list(_stack)
        _stack.clear()
        try:
            return type(self)(**vals)
        finally:
            _stack.extend(saved)

    def _run(self, root, method, args, kwargs):
        m = self.clone()
        m._st["root"] = root
        m._st["scope"] = root.params
        fn = getattr(type(m), method.__name__) if method is not None else type(m).__call__
        saved = list(_stack)
        _stack.clear()
        try:
            return fn(m, *args, **kwargs)
        finally:
            _stack.clear()
            _stack.extend(saved)

    def init(self, rngs, *args, method=None, **kwargs):
        root = _Root({}, True)
        self._run(root, method, args, kwargs)

        def prune(d):   # parameter-free submodules leave no entry, as in flax
            out = {}
            for k, v in d.items():
                if isinstance(v, dict):
                    v = prune(v)
                    if v:
                        out[k] = v
                else:
                    out[k] = v
            return out

        return {"params": prune(root.params)}

    def apply(self, variables, *args, method=None, **kwargs):
        root = _Root(variables["params"], False)
        return self._run(root, method, args, kwargs)


# ---------------------------------------------------------------------------------------------------------------
class Dense(Module):
    """``y = x @ kernel + bias``; ``kernel (in, features)`` LeCun-normal, ``bias (features,)`` zeros."""

    features: int
    use_bias: bool = True
    kernel_init: Any = None
    bias_init: Any = None

    @compact
    def __call__(self, x):
        kinit = self.kernel_init or initializers.lecun_normal()
        binit = self.bias_init or initializers.zeros
        kernel = self.param("kernel", kinit, (x.shape[-1], self.features))
        y = x @ kernel
        if self.use_bias:
            y = y + self.param("bias", binit, (self.features,))
        return y


class DenseGeneral(Module):
    """Contraction of the trailing ``axis`` dims of x with ``kernel (*in_dims, *features)``; bias ``(*features)``."""

    features: Any
    axis: Any = -1
    use_bias: bool = True
    kernel_init: Any = None
    bias_init: Any = None

    @compact
    def __call__(self, x):
        feats = tuple(self.features) if isinstance(self.features, (list, tuple)) else (int(self.features),)
        axes = tuple(self.axis) if isinstance(self.axis, (list, tuple)) else (int(self.axis),)
        n_in = len(axes)
        assert sorted(a % x.ndim for a in axes) == list(range(x.ndim - n_in, x.ndim)), "refshim: trailing axes only"
        in_dims = tuple(x.shape[-n_in:])

        def kinit(key, shape, dtype=None):   # flax flattens (prod(in), prod(out)) for the fan-in computation
            fan_in = 1
            for s in in_dims:
                fan_in *= s
            from jax import random

            return random.normal(key, tuple(shape)) / math.sqrt(fan_in)

        kernel = self.param("kernel", self.kernel_init or kinit, in_dims + feats)
        y = torch.tensordot(x, kernel, dims=n_in)
        if self.use_bias:
            y = y + self.param("bias", self.bias_init or initializers.zeros, feats)
        return y


class LayerNorm(Module):
    """Over the last axis; flax's default fast variance ``E[x^2] - E[x]^2`` (clipped at 0), ``scale`` ones, ``bias``
    zeros."""

    epsilon: float = 1e-6
    use_bias: bool = True
    use_scale: bool = True

    @compact
    def __call__(self, x):
        mean = x.mean(dim=-1, keepdim=True)
        var = torch.clamp((x * x).mean(dim=-1, keepdim=True) - mean * mean, min=0.0)
        y = (x - mean) * torch.rsqrt(var + self.epsilon)
        f = x.shape[-1]
        if self.use_scale:
            y = y * self.param("scale", initializers.ones, (f,))
        if self.use_bias:
            y = y + self.param("bias", initializers.zeros, (f,))
        return y


class MultiHeadDotProductAttention(Module):
    """Self-attention as flax applies it to one input: ``query`` / ``key`` / ``value`` DenseGeneral ``(in, H, dh)``,
    ``softmax(q k^T / sqrt(dh))``, ``out`` DenseGeneral over ``(H, dh)``."""

    num_heads: int
    qkv_features: Any = None
    out_features: Any = None
    deterministic: Any = None
    use_bias: bool = True

    @compact
    def __call__(self, x):
        feat = self.qkv_features or x.shape[-1]
        out_f = self.out_features or x.shape[-1]
        dh = feat // self.num_heads
        q = DenseGeneral((self.num_heads, dh), use_bias=self.use_bias, name="query")(x)
        k = DenseGeneral((self.num_heads, dh), use_bias=self.use_bias, name="key")(x)
        v = DenseGeneral((self.num_heads, dh), use_bias=self.use_bias, name="value")(x)
        q = q / math.sqrt(dh)
        w = torch.softmax(torch.einsum("...qhd,...khd->...hqk", q, k), dim=-1)
        o = torch.einsum("...hqk,...khd->...qhd", w, v)
        return DenseGeneral(out_f, axis=(-2, -1), use_bias=self.use_bias, name="out")(o)
