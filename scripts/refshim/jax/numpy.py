"""jax.numpy subset over torch float64 (see ../README.md).  Test infrastructure."""
import builtins
import math

import numpy as _np
import torch

ndarray = torch.Tensor
pi = math.pi
float32 = torch.float64   # everything is evaluated in float64: parameter dtypes declared float32 are widened
float64 = torch.float64
complex64 = torch.complex128
complex128 = torch.complex128
int32 = torch.int64
int64 = torch.int64
bool_ = torch.bool
newaxis = None
inf = math.inf


def _dt(dtype):
    if dtype is None:
        return None
    if dtype in (float, builtins.float):
        return torch.float64
    if dtype in (int, builtins.int):
        return torch.int64
    if dtype in (bool,):
        return torch.bool
    if dtype in (complex,):
        return torch.complex128
    if isinstance(dtype, torch.dtype):
        return torch.float64 if dtype == torch.float32 else dtype
    return torch.float64


def asarray(x, dtype=None):
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = _np.asarray(x)
        t = torch.as_tensor(a)
    if dtype is not None:
        return t.to(_dt(dtype))
    if t.dtype in (torch.float32, torch.float16):
        t = t.to(torch.float64)
    if t.dtype == torch.complex64:
        t = t.to(torch.complex128)
    if t.dtype == torch.int32:
        t = t.to(torch.int64)
    return t


array = asarray


def _ax(axis):
    if isinstance(axis, list):
        return tuple(axis)
    return axis


def zeros(shape, dtype=None):
    return torch.zeros(shape if not isinstance(shape, int) else (shape,), dtype=_dt(dtype) or torch.float64)


def ones(shape, dtype=None):
    return torch.ones(shape if not isinstance(shape, int) else (shape,), dtype=_dt(dtype) or torch.float64)


def zeros_like(x, dtype=None):
    return torch.zeros_like(asarray(x), dtype=_dt(dtype))


def ones_like(x, dtype=None):
    return torch.ones_like(asarray(x), dtype=_dt(dtype))


def eye(n, dtype=None):
    return torch.eye(n, dtype=_dt(dtype) or torch.float64)


def arange(*args, dtype=None):
    t = torch.arange(*args)
    return t.to(_dt(dtype)) if dtype is not None else t


def shape(x):
    return tuple(asarray(x).shape)


def reshape(x, newshape):
    return asarray(x).reshape(tuple(newshape) if not isinstance(newshape, int) else (newshape,))


def transpose(x, axes=None):
    x = asarray(x)
    if axes is None:
        return x.permute(*reversed(range(x.ndim)))
    return x.permute(*axes)


def broadcast_to(x, shape):
    return asarray(x).broadcast_to(tuple(shape))


def concatenate(xs, axis=0):
    return torch.cat([asarray(x) for x in xs], dim=axis)


def stack(xs, axis=0):
    return torch.stack([asarray(x) for x in xs], dim=axis)


def split(x, indices_or_sections, axis=0):
    if isinstance(indices_or_sections, int):
        return list(torch.tensor_split(x, indices_or_sections, dim=axis))
    return list(torch.tensor_split(x, [int(i) for i in indices_or_sections], dim=axis))


def meshgrid(*xs, indexing="xy"):
    return list(torch.meshgrid(*[asarray(x) for x in xs], indexing=indexing))


def _red(fn):
    def f(x, axis=None, keepdims=False):
        x = asarray(x)
        if axis is None:
            return fn(x)
        return fn(x, dim=_ax(axis), keepdim=keepdims)

    return f


sum = _red(torch.sum)  # noqa: A001
mean = _red(torch.mean)


def _minmax(fn):
    def f(x, axis=None, keepdims=False):
        x = asarray(x)
        if axis is None:
            return fn(x)
        return fn(x, dim=_ax(axis), keepdim=keepdims)

    return f


max = amax = _minmax(torch.amax)  # noqa: A001
min = amin = _minmax(torch.amin)  # noqa: A001


def argmin(x, axis=None):
    return torch.argmin(asarray(x)) if axis is None else torch.argmin(asarray(x), dim=axis)


def all(x, axis=None):  # noqa: A001
    x = asarray(x)
    return torch.all(x) if axis is None else torch.all(x, dim=axis)


def allclose(a, b, rtol=1e-5, atol=1e-8):
    return bool(torch.allclose(asarray(a).double(), asarray(b).double(), rtol=rtol, atol=atol))


def _un(fn):
    def f(x):
        x = asarray(x)
        if not (x.is_floating_point() or x.is_complex()):
            x = x.to(torch.float64)
        return fn(x)

    return f


exp = _un(torch.exp)
log = _un(torch.log)
log1p = _un(torch.log1p)
sqrt = _un(torch.sqrt)
sin = _un(torch.sin)
cos = _un(torch.cos)
tanh = _un(torch.tanh)
floor = _un(torch.floor)
ceil = _un(torch.ceil)
sign = _un(torch.sign)
square = _un(torch.square)
arctan = _un(torch.arctan)
arccos = _un(torch.arccos)
arcsin = _un(torch.arcsin)


def tile(x, reps):
    return asarray(x).repeat(*reps) if not isinstance(reps, int) else asarray(x).repeat(reps)


def abs(x):  # noqa: A001
    return torch.abs(asarray(x))


def mod(a, b):
    return torch.remainder(asarray(a), b)


def where(c, a, b):
    a = a if isinstance(a, torch.Tensor) else torch.as_tensor(a, dtype=(b.dtype if isinstance(b, torch.Tensor) else torch.float64))
    b = b if isinstance(b, torch.Tensor) else torch.as_tensor(b, dtype=a.dtype)
    return torch.where(c, a, b)


def dot(a, b):
    a, b = asarray(a), asarray(b)
    if a.dtype != b.dtype:
        dt = torch.promote_types(a.dtype, b.dtype)
        a, b = a.to(dt), b.to(dt)
    if a.ndim == 0 or b.ndim == 0:
        return a * b
    return a @ b


def matmul(a, b):
    return dot(a, b)


def einsum(eq, *ops):
    ops = [asarray(o) for o in ops]
    dt = ops[0].dtype
    for o in ops[1:]:
        dt = torch.promote_types(dt, o.dtype)
    return torch.einsum(eq, *[o.to(dt) for o in ops])


def triu(x, k=0):
    return torch.triu(asarray(x), diagonal=k)


def tril(x, k=0):
    return torch.tril(asarray(x), diagonal=k)


def triu_indices(n, k=0, m=None):
    idx = torch.triu_indices(n, m if m is not None else n, offset=k)
    return idx[0], idx[1]


def diagonal(x, offset=0, axis1=0, axis2=1):
    return torch.diagonal(asarray(x), offset=offset, dim1=axis1, dim2=axis2)


def diag(x, k=0):
    return torch.diag(asarray(x), diagonal=k)


def take_along_axis(x, idx, axis):
    return torch.take_along_dim(asarray(x), idx, dim=axis)


def iscomplexobj(x):
    return isinstance(x, complex) or (isinstance(x, torch.Tensor) and x.is_complex())


def isscalar(x):
    return _np.isscalar(x)


def real(x):
    return asarray(x).real


def imag(x):
    return asarray(x).imag


def conj(x):
    return asarray(x).conj()


class _Linalg:
    @staticmethod
    def norm(x, ord=None, axis=None, keepdims=False):  # noqa: A002
        x = asarray(x)
        if axis is None:
            return torch.linalg.norm(x)
        return torch.linalg.vector_norm(x, dim=_ax(axis), keepdim=keepdims)

    inv = staticmethod(lambda a: torch.linalg.inv(asarray(a)))
    pinv = staticmethod(lambda a: torch.linalg.pinv(asarray(a)))
    det = staticmethod(lambda a: torch.linalg.det(asarray(a)))

    @staticmethod
    def slogdet(a):
        s, l = torch.linalg.slogdet(asarray(a))
        return s, l


linalg = _Linalg()


# ---- array methods jax arrays have and torch tensors lack ------------------------------------------------------
def _astype(self, dtype):
    return self.to(_dt(dtype))


class _AtIndexer:
    def __init__(self, t):
        self.t = t

    def __getitem__(self, idx):
        return _AtSetter(self.t, idx)


class _AtSetter:
    def __init__(self, t, idx):
        self.t, self.idx = t, idx

    def set(self, v):
        out = self.t.clone()
        out[self.idx] = v
        return out

    def add(self, v):
        out = self.t.clone()
        out[self.idx] = out[self.idx] + v
        return out


if not hasattr(torch.Tensor, "astype"):
    torch.Tensor.astype = _astype
if not hasattr(torch.Tensor, "at"):
    torch.Tensor.at = property(lambda self: _AtIndexer(self))

# jax promotes mixed integer / float operands of ``@``; torch.matmul refuses them
if not getattr(torch.Tensor, "_refshim_matmul", False):
    _orig_matmul = torch.Tensor.__matmul__

    def _promoting_matmul(self, other):
        if isinstance(other, torch.Tensor) and self.dtype != other.dtype:
            dt = torch.promote_types(self.dtype, other.dtype)
            return _orig_matmul(self.to(dt), other.to(dt))
        return _orig_matmul(self, other)

    torch.Tensor.__matmul__ = _promoting_matmul
    torch.Tensor._refshim_matmul = True

# ``a // b`` under jax autodiff has a zero derivative; torch's floor_divide has no derivative formula at all.
# floor(a / b) is the same value with a zero gradient.
if not getattr(torch.Tensor, "_refshim_floordiv", False):
    def _floordiv(self, other):
        return torch.floor(torch.true_divide(self, other))

    def _rfloordiv(self, other):
        return torch.floor(torch.true_divide(other, self))

    torch.Tensor.__floordiv__ = _floordiv
    torch.Tensor.__rfloordiv__ = _rfloordiv
    torch.Tensor._refshim_floordiv = True
