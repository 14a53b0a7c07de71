"""Dense forward-Laplacian recurrences (oracle; test infrastructure only).

Restates the propagation rules of the reference's forward-Laplacian interpreter for the dense
representation: every tracked tensor carries ``(x, J, L)`` with ``J[k] = dx/dr_k`` (basis axis
leading, reference ``laplacian/types.py:29-53``) and ``L = sum_k d2x/dr_k^2``.

Rules and where they come from in the reference (``src/jaqmc/laplacian/primitives``):

* affine / linear maps (reshape, slice, sum, mean, concat ...): ``dot_general.py:377-407``,
  ``arithmetic.py:191-266``, ``reductions.py:25-61``, ``shape.py:578-626``
* bilinear products: ``dot_general.py:410-449``, ``core.py:450-501``
* unary elementwise ``f``: ``elementwise.py:42-72`` (``f' J``, ``f' L + f'' sum_k J_k^2``)
* division: ``arithmetic.py:472-550``
* ``slogdet``: ``slogdet.py:25-72``
* seed: ``seed.py:75-121`` (identity Jacobian, zero Laplacian)

Every function also accepts plain ``torch.Tensor`` arguments (then it is just the primal op), so
that one restatement of each network serves both derivative routes: these recurrences and the
brute-force autograd Hessian trace (``tests/laplacian/helpers.py:45-125`` in the reference).
"""

from __future__ import annotations

import torch


class Lap:
    """``(x, jac, lap)`` triple; ``jac`` has the tracked-coordinate axis leading."""

    __slots__ = ("x", "jac", "lap")

    def __init__(self, x, jac, lap):
        self.x, self.jac, self.lap = x, jac, lap

    # ---- shape helpers -------------------------------------------------
    @property
    def shape(self):
        return self.x.shape

    @property
    def K(self):
        return self.jac.shape[0]

    @property
    def dtype(self):
        return self.x.dtype

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        return Lap(self.x[idx], self.jac[(slice(None),) + idx], self.lap[idx])

    # ---- arithmetic ----------------------------------------------------
    def __neg__(self):
        return Lap(-self.x, -self.jac, -self.lap)

    def _pad(self, nd):
        """Unsqueeze leading primal dims so broadcasting never mixes the basis axis with a primal axis."""
        extra = nd - self.x.dim()
        if extra <= 0:
            return self
        x = self.x.reshape((1,) * extra + tuple(self.x.shape))
        return Lap(x, self.jac.reshape((self.K, *x.shape)), self.lap.reshape(x.shape))

    def __add__(self, o):
        if isinstance(o, Lap):
            nd = max(self.x.dim(), o.x.dim())
            a, b = self._pad(nd), o._pad(nd)
            x = a.x + b.x
            return Lap(x, (a.jac + b.jac).expand((a.K, *x.shape)), (a.lap + b.lap).expand(x.shape))
        o = torch.as_tensor(o, dtype=_promote(self.x, o))
        a = self._pad(o.dim())
        x = a.x + o
        return Lap(x, a.jac.expand((a.K, *x.shape)).to(x.dtype), a.lap.expand(x.shape).to(x.dtype))

    __radd__ = __add__

    def __sub__(self, o):
        return self + (-o)

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        if isinstance(o, Lap):
            nd = max(self.x.dim(), o.x.dim())
            a, b = self._pad(nd), o._pad(nd)
            x = a.x * b.x
            jac = a.jac * b.x + a.x * b.jac
            lap = a.lap * b.x + a.x * b.lap + 2.0 * (a.jac * b.jac).sum(0)
            return Lap(x, jac, lap)
        o = torch.as_tensor(o, dtype=_promote(self.x, o))
        a = self._pad(o.dim())
        return Lap(a.x * o, a.jac * o, a.lap * o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Lap):
            return self * reciprocal(o)
        o = torch.as_tensor(o, dtype=_promote(self.x, o))
        return self * (1.0 / o)

    def __rtruediv__(self, o):
        return reciprocal(self) * o


def _promote(x, o):
    if isinstance(o, torch.Tensor):
        return torch.promote_types(x.dtype, o.dtype)
    if isinstance(o, complex):
        return torch.promote_types(x.dtype, torch.complex64)
    return x.dtype


def is_lap(a) -> bool:
    return isinstance(a, Lap)


def seed(x: torch.Tensor) -> Lap:
    """Identity-Jacobian seed for the coordinates ``x`` (reference ``seed.py:75-121``)."""
    k = x.numel()
    jac = torch.eye(k, dtype=x.dtype).reshape(k, *x.shape)
    return Lap(x, jac, torch.zeros_like(x))


def const(x: torch.Tensor, like: Lap) -> Lap:
    return Lap(x, torch.zeros((like.K, *x.shape), dtype=x.dtype), torch.zeros_like(x))


def value(a):
    return a.x if isinstance(a, Lap) else a


# ---- unary elementwise -------------------------------------------------------
def _unary(a, f, df, d2f):
    if not isinstance(a, Lap):
        return f(a)
    y = f(a.x)
    d1 = df(a.x, y)
    d2 = d2f(a.x, y)
    return Lap(y, d1 * a.jac, d1 * a.lap + d2 * (a.jac * a.jac).sum(0))


def tanh(a):
    return _unary(a, torch.tanh, lambda x, y: 1 - y * y, lambda x, y: -2 * y * (1 - y * y))


def exp(a):
    return _unary(a, torch.exp, lambda x, y: y, lambda x, y: y)


def log(a):
    return _unary(a, torch.log, lambda x, y: 1 / x, lambda x, y: -1 / (x * x))


def log1p(a):
    return _unary(a, torch.log1p, lambda x, y: 1 / (1 + x), lambda x, y: -1 / ((1 + x) ** 2))


def sqrt(a):
    return _unary(a, torch.sqrt, lambda x, y: 0.5 / y, lambda x, y: -0.25 / (y * x))


def rsqrt(a):
    return _unary(a, torch.rsqrt, lambda x, y: -0.5 * y / x, lambda x, y: 0.75 * y / (x * x))


def sin(a):
    return _unary(a, torch.sin, lambda x, y: torch.cos(x), lambda x, y: -y)


def cos(a):
    return _unary(a, torch.cos, lambda x, y: -torch.sin(x), lambda x, y: -y)


def square(a):
    return _unary(a, lambda x: x * x, lambda x, y: 2 * x, lambda x, y: 2 * torch.ones_like(x))


def reciprocal(a):
    return _unary(a, lambda x: 1 / x, lambda x, y: -y * y, lambda x, y: 2 * y * y * y)


def abs_(a):
    """|x| for real x: multiply J and L by sign(x) (reference ``elementwise.py:118-127``)."""
    return _unary(a, torch.abs, lambda x, y: torch.sign(x), lambda x, y: torch.zeros_like(x))


def erfc(a):
    c = 2.0 / torch.pi**0.5
    return _unary(
        a,
        torch.erfc,
        lambda x, y: -c * torch.exp(-x * x),
        lambda x, y: 2 * c * x * torch.exp(-x * x),
    )


# ---- linear structure ----------------------------------------------------------
def linear_map(fn, a):
    """Apply a *linear* torch function (reshape/slice/sum/mean/transpose/...) to all three parts."""
    if not isinstance(a, Lap):
        return fn(a)
    return Lap(fn(a.x), torch.stack([fn(j) for j in a.jac]) if a.K else fn(a.jac), fn(a.lap))


def matmul(a, w: torch.Tensor):
    """``a @ w`` with untracked ``w`` (one GEMM over rows ``{x, J_1..J_K, L}``)."""
    if not isinstance(a, Lap):
        return a @ w.to(a.dtype)
    w = w.to(a.dtype)
    return Lap(a.x @ w, a.jac @ w, a.lap @ w)


def dense(a, kernel: torch.Tensor, bias: torch.Tensor | None = None):
    """flax ``nn.Dense`` / ``nn.DenseGeneral``: contract the last axis of ``a`` with axis 0 of ``kernel``."""
    k2 = kernel.reshape(kernel.shape[0], -1)
    out_shape = tuple(kernel.shape[1:])
    y = matmul(a, k2)
    y = linear_map(lambda t: t.reshape(*t.shape[:-1], *out_shape), y)
    if bias is not None:
        y = y + bias.to(value(y).dtype)
    return y


def bilinear(fn, a, b):
    """Bilinear torch function of two possibly-tracked operands (reference ``dot_general.py:410-449``)."""
    la, lb = isinstance(a, Lap), isinstance(b, Lap)
    if not la and not lb:
        return fn(a, b)
    if la and not lb:
        return Lap(fn(a.x, b), torch.stack([fn(j, b) for j in a.jac]), fn(a.lap, b))
    if lb and not la:
        return Lap(fn(a, b.x), torch.stack([fn(a, j) for j in b.jac]), fn(a, b.lap))
    x = fn(a.x, b.x)
    jac = torch.stack([fn(ja, b.x) + fn(a.x, jb) for ja, jb in zip(a.jac, b.jac)])
    cross = sum(fn(ja, jb) for ja, jb in zip(a.jac, b.jac))
    lap = fn(a.lap, b.x) + fn(a.x, b.lap) + 2.0 * cross
    return Lap(x, jac, lap)


def einsum(eq: str, a, b):
    return bilinear(lambda p, q: torch.einsum(eq, p, q), a, b)


def cat(parts, dim: int):
    """Concatenate tracked and untracked parts (constants get a zero Jacobian)."""
    ref = next((p for p in parts if isinstance(p, Lap)), None)
    if ref is None:
        return torch.cat(parts, dim=dim)
    dt = ref.dtype
    for p in parts:
        dt = torch.promote_types(dt, value(p).dtype)
    ps = [p if isinstance(p, Lap) else const(p.to(dt), ref) for p in parts]
    d = dim if dim >= 0 else dim + ps[0].x.dim()
    return Lap(
        torch.cat([p.x.to(dt) for p in ps], dim=d),
        torch.cat([p.jac.to(dt) for p in ps], dim=d + 1),
        torch.cat([p.lap.to(dt) for p in ps], dim=d),
    )


def sum_(a, dim, keepdim=False):
    return linear_map(lambda t: t.sum(dim=dim, keepdim=keepdim), a)


def mean(a, dim, keepdim=False):
    return linear_map(lambda t: t.mean(dim=dim, keepdim=keepdim), a)


def reshape(a, *shape):
    return linear_map(lambda t: t.reshape(*shape), a)


def transpose(a, *perm):
    return linear_map(lambda t: t.permute(*perm), a)


def broadcast_to(a, shape):
    return linear_map(lambda t: t.expand(shape), a)


def stop_gradient(a):
    return value(a).detach() if isinstance(value(a), torch.Tensor) else value(a)


def to_complex(a):
    cd = torch.complex128
    if isinstance(a, Lap):
        return Lap(a.x.to(cd), a.jac.to(cd), a.lap.to(cd))
    return a.to(cd)


# ---- slogdet -----------------------------------------------------------------------
def logdet(a):
    """Batched ``log det`` of ``(..., n, n)`` matrices as ``(sign_or_phase, log|det|)``.

    Forward-Laplacian rule (reference ``laplacian/primitives/slogdet.py:46-72``):
    ``ld_J[k] = tr(A^-1 A_J[k])``, ``ld_L = tr(A^-1 A_L) - sum_k tr((A^-1 A_J[k])^2)``.
    Real input: the sign is untracked.  Complex input: the returned log-determinant is the complex
    ``log|det| + i*arg(det)`` carried as one holomorphic quantity (its real / imaginary parts are
    exactly the reference's ``logabs`` and phase-angle derivatives ``slogdet.py:65-72``).
    """
    if not isinstance(a, Lap):
        sign, logabs = torch.linalg.slogdet(a)
        if a.is_complex():
            return sign, logabs.to(a.dtype) + 1j * torch.angle(sign)
        return sign, logabs
    sign, logabs = torch.linalg.slogdet(a.x)
    inv = torch.linalg.inv(a.x)
    m = inv @ a.jac  # (K, ..., n, n)
    tr_j = torch.diagonal(m, dim1=-2, dim2=-1).sum(-1)  # (K, ...)
    tr_l = torch.diagonal(inv @ a.lap, dim1=-2, dim2=-1).sum(-1)
    tr_sq = (m * m.transpose(-1, -2)).sum((-1, -2)).sum(0)
    if a.x.is_complex():
        ld = logabs.to(a.x.dtype) + 1j * torch.angle(sign)
        return sign, Lap(ld, tr_j, tr_l - tr_sq)
    return sign, Lap(logabs, tr_j, tr_l - tr_sq)


# ---- brute-force route -----------------------------------------------------------------
def brute_force(fn, x: torch.Tensor):
    """Value, gradient and Laplacian of scalar ``fn(x)`` via autograd Jacobian + full Hessian.

    Mirrors the reference's test oracle (``tests/laplacian/helpers.py:45-125``).  Complex outputs
    are differentiated as real and imaginary parts.
    """
    shape = x.shape
    flat = x.reshape(-1).clone()

    def f_real(v):
        return torch.real(fn(v.reshape(shape)))

    def f_imag(v):
        return torch.imag(fn(v.reshape(shape)))

    val = fn(x)
    g = torch.func.jacrev(f_real)(flat)
    h = torch.func.hessian(f_real)(flat)
    lap = torch.trace(h)
    if val.is_complex():
        gi = torch.func.jacrev(f_imag)(flat)
        hi = torch.func.hessian(f_imag)(flat)
        g = g + 1j * gi
        lap = lap + 1j * torch.trace(hi)
    return val, g, lap
