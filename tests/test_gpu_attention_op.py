"""``jaqmc_b200_attention_fl`` on the GPU: every attention kernel (CUDA-core block kernel, warp-per-component kernel,
mma.sync 3xTF32 tensor-core kernel with truncating (3) and round-to-nearest (4) operand splits, and its sparse
logit-Jacobian phase for one-electron queries / keys (5)) on the SAME random augmented operands against a float64 statement of the
forward-Laplacian rule of ``softmax(q k^T / sqrt(d)) v``.  Operand-level, so that the tile-shape edge cases
(n = 17, 32, 33, 48 ...) are tested without the conditioning of a wavefunction entering: the network-level parity tests
in test_gpu_attention_nets.py cannot separate a kernel error from an ill-conditioned walker.

The float64 rule below is itself pinned (``test_rule_matches_autograd_of_the_oracle``, CPU) against second-order
autograd through the oracle's ``attention_core`` (reference ``_attention.py:20-34``).

Tolerances were fixed before the first run, the same for every kernel: value and Jacobian rows 1e-5, Laplacian row 5e-5
of the largest reference magnitude of that row type (cf. test_gpu_dense.py: 5e-6 / 3e-5 for one dense layer)."""

import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import networks as ON

F64 = torch.float64
TOL_VJ, TOL_L = 1e-5, 5e-5


def rule_f64(q, k, v):
    """q, k, v: (n, C, H, d) float64 dense augmented operands, C = 3n+2.  Returns out (n, C, H, d)."""
    n, Cc, H, d = q.shape
    K = Cc - 2
    s = 1.0 / math.sqrt(d)
    q0, k0, v0 = q[:, 0], k[:, 0], v[:, 0]
    qJ, kJ, vJ = q[:, 1:1 + K], k[:, 1:1 + K], v[:, 1:1 + K]
    qL, kL, vL = q[:, -1], k[:, -1], v[:, -1]
    S = torch.einsum("ihd,jhd->hij", q0, k0) * s
    w = torch.softmax(S, dim=-1)
    SJ = (torch.einsum("ichd,jhd->chij", qJ, k0) + torch.einsum("ihd,jchd->chij", q0, kJ)) * s
    SL = (torch.einsum("ihd,jhd->hij", qL, k0) + torch.einsum("ihd,jhd->hij", q0, kL)
          + 2.0 * torch.einsum("ichd,jchd->hij", qJ, kJ)) * s
    cJ = SJ - (w[None] * SJ).sum(-1, keepdim=True)
    wJ = w[None] * cJ
    wL = (wJ * cJ).sum(0) + w * (SL - (w * SL).sum(-1, keepdim=True) - (wJ * SJ).sum(-1, keepdim=True).sum(0))
    out = torch.empty_like(v)
    out[:, 0] = torch.einsum("hij,jhd->ihd", w, v0)
    out[:, 1:1 + K] = torch.einsum("chij,jhd->ichd", wJ, v0) + torch.einsum("hij,jchd->ichd", w, vJ)
    out[:, -1] = (torch.einsum("hij,jhd->ihd", wL, v0) + torch.einsum("hij,jhd->ihd", w, vL)
                  + 2.0 * torch.einsum("chij,jchd->ihd", wJ, vJ))
    return out


def test_rule_matches_autograd_of_the_oracle():
    """q(x) = q0 + sum_k qJ_k x_k + sum_k qL x_k^2 / (2K) has Jacobian qJ and Laplacian qL at x = 0."""
    g = torch.Generator().manual_seed(3)
    n, H, d = 3, 2, 4
    K = 3 * n
    ops = [torch.randn(n, K + 2, H, d, generator=g, dtype=F64) for _ in range(3)]

    def embed(t, x):
        return t[:, 0] + torch.einsum("ichd,c->ihd", t[:, 1:1 + K], x) + t[:, -1] * (x * x).sum() / (2.0 * K)

    def f(x):
        return ON.attention_core(embed(ops[0], x), embed(ops[1], x), embed(ops[2], x))

    x0 = torch.zeros(K, dtype=F64)
    jac = torch.autograd.functional.jacobian(f, x0)                      # (n, H, d, K)
    lap = torch.zeros(n, H, d, dtype=F64)
    for a in range(n):
        for b in range(H):
            for c in range(d):
                hes = torch.autograd.functional.hessian(lambda x: f(x)[a, b, c], x0)
                lap[a, b, c] = torch.diagonal(hes).sum()
    out = rule_f64(*ops)
    assert torch.allclose(out[:, 0], f(x0), atol=1e-12)
    assert torch.allclose(out[:, 1:1 + K], jac.permute(0, 3, 1, 2), atol=1e-12)
    assert torch.allclose(out[:, -1], lap, atol=1e-11)


def _operands(n, H, d, W, local_qk, seed):
    g = torch.Generator().manual_seed(seed)
    K = 3 * n
    amp = torch.cat([torch.ones(1), 0.3 * torch.ones(K), torch.ones(1)]).to(F64)[None, None, :, None, None]
    dense = [torch.randn(W, n, K + 2, H, d, generator=g, dtype=F64) * amp for _ in range(3)]
    stored = list(dense)
    if local_qk:
        for t in range(2):
            loc = torch.randn(W, n, 5, H, d, generator=g, dtype=F64)
            loc[:, :, 1:4] *= 0.3
            full = torch.zeros_like(dense[t])
            full[:, :, 0], full[:, :, -1] = loc[:, :, 0], loc[:, :, 4]
            for i in range(n):
                full[:, i, 1 + 3 * i:4 + 3 * i] = loc[:, i, 1:4]
            dense[t], stored[t] = full, loc
    # the kernels see float32 operands: the reference starts from the same rounded values
    dense = [t.float().double() for t in dense]
    stored = [t.float().contiguous() for t in stored]
    return dense, stored


def _run(kernel, stored, n, H, d, W, local_qk):
    from jaqmc_b200._lib import cuda_library

    lib = cuda_library()
    dev = torch.device("cuda", 0)
    q, k, v = [t.reshape(W, n, t.shape[2], H * d).to(dev).contiguous() for t in stored]
    out = torch.full((W, n, 3 * n + 2, H * d), float("nan"), device=dev)
    p = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_float))
    qc = 5 if local_qk else 3 * n + 2
    rc = lib.jaqmc_b200_attention_fl(p(q), p(k), p(v), p(out), W, n, H, d, qc, qc, kernel,
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return rc, out.reshape(W, n, 3 * n + 2, H, d).double().cpu()


def _errors(out, ref):
    K = ref.shape[2] - 2
    rows = {"value": (slice(0, 1), TOL_VJ), "jacobian": (slice(1, 1 + K), TOL_VJ), "laplacian": (slice(K + 1, K + 2), TOL_L)}
    res = {}
    for name, (sl, tol) in rows.items():
        res[name] = ((out[:, :, sl] - ref[:, :, sl]).abs().max() / ref[:, :, sl].abs().max()).item(), tol
    return res


# (n, kernels that must support the shape)
SHAPES = [(5, (1, 2, 3, 4)), (14, (1, 2, 3, 4)), (15, (1, 3, 4)), (16, (1, 3, 4)), (17, (1, 3, 4)), (18, (1, 3, 4)), (24, (1, 3, 4)), (30, (1, 3, 4)), (32, (1, 3, 4)),
          (33, (1, 3, 4)), (40, (1, 3, 4)), (42, (1, 3, 4)), (47, (1, 3, 4)), (48, (1, 3, 4))]


@pytest.mark.gpu
@pytest.mark.parametrize("local_qk", [False, True], ids=["dense_qk", "local_qk"])
@pytest.mark.parametrize("n,kernels", SHAPES, ids=["n%d" % s[0] for s in SHAPES])
def test_attention_kernels_match_float64_rule(n, kernels, local_qk):
    H, d, W = 2, 64, 3
    dense, stored = _operands(n, H, d, W, local_qk, seed=100 + n)
    ref = torch.stack([rule_f64(dense[0][w], dense[1][w], dense[2][w]) for w in range(W)])
    if local_qk and 3 in kernels:
        kernels = kernels + (5,)      # the tensor-core kernel's sparse phase for one-electron q / k
    outs = {}
    for kern in kernels + (0,):
        rc, out = _run(kern, stored, n, H, d, W, local_qk)
        assert rc == 0, (kern, rc)
        assert torch.isfinite(out).all(), kern
        errs = _errors(out, ref)
        print("n=%d kernel=%d %s" % (n, kern, {a: "%.2e" % b[0] for a, b in errs.items()}))
        for name, (e, tol) in errs.items():
            assert e < tol, (n, kern, name, e)
        outs[kern] = out
    # the library's own choice is one of the forced kernels, bit for bit
    assert any(torch.equal(outs[0], outs[kern]) for kern in kernels)


@pytest.mark.gpu
def test_forced_kernel_rejects_unsupported_shapes():
    for n, kern in ((20, 2), (16, 2), (50, 3)):   # n = 15, 16: the warp kernel's staging exceeds shared memory
        dense, stored = _operands(n, 1, 64, 1, False, seed=1)
        rc, _ = _run(kern, stored, n, 1, 64, 1, False)
        assert rc != 0, (n, kern)
