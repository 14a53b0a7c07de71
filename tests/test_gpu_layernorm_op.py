"""``jaqmc_b200_layernorm_fl`` on the GPU: the three LayerNorm kernels on the same random augmented input against a
float64 statement of the forward-Laplacian rule, itself pinned (CPU test) against second-order autograd through the
oracle's ``layer_norm``.  Tolerances fixed before the first run, the same for every kernel: value and Jacobian rows 1e-5,
Laplacian row 5e-5 of the largest reference magnitude of that row type."""

import ctypes as C

import pytest
import torch

from oracle import networks as ON

F64 = torch.float64
TOL_VJ, TOL_L = 1e-5, 5e-5
EPS = 1e-5


def rule_f64(x, scale, bias, eps):
    """x (C, F) float64 with C = K + 2 rows {value, K Jacobian rows, Laplacian}."""
    K = x.shape[0] - 2
    x0, J, L = x[0], x[1:1 + K], x[-1]
    mu = x0.mean()
    xc = x0 - mu
    var = (xc * xc).mean()
    s = (var + eps) ** -0.5
    Jc = J - J.mean(-1, keepdim=True)
    Lc = L - L.mean()
    varJ = 2.0 * (xc[None] * Jc).mean(-1)
    varL = 2.0 * (xc * Lc).mean() + 2.0 * (Jc * Jc).mean(-1).sum()
    sJ = -0.5 * s ** 3 * varJ
    sL = -0.5 * s ** 3 * varL + 0.75 * s ** 5 * (varJ * varJ).sum()
    out = torch.empty_like(x)
    out[0] = xc * s * scale + bias
    out[1:1 + K] = (Jc * s + xc[None] * sJ[:, None]) * scale
    out[-1] = (Lc * s + xc * sL + 2.0 * (Jc * sJ[:, None]).sum(0)) * scale
    return out


def test_rule_matches_autograd_of_the_oracle():
    g = torch.Generator().manual_seed(5)
    K, F = 6, 8
    x = torch.randn(K + 2, F, generator=g, dtype=F64)
    p = {"scale": torch.randn(F, generator=g, dtype=F64), "bias": torch.randn(F, generator=g, dtype=F64)}

    def f(t):
        h = x[0] + torch.einsum("kf,k->f", x[1:1 + K], t) + x[-1] * (t * t).sum() / (2.0 * K)
        return ON.layer_norm(p, h, EPS)

    t0 = torch.zeros(K, dtype=F64)
    out = rule_f64(x, p["scale"], p["bias"], EPS)
    jac = torch.autograd.functional.jacobian(f, t0)          # (F, K)
    lap = torch.stack([torch.diagonal(torch.autograd.functional.hessian(lambda t: f(t)[j], t0)).sum() for j in range(F)])
    assert torch.allclose(out[0], f(t0), atol=1e-12)
    assert torch.allclose(out[1:1 + K], jac.T, atol=1e-12)
    assert torch.allclose(out[-1], lap, atol=1e-11)


def _run(kernel, x, scale, bias):
    from jaqmc_b200._lib import cuda_library

    lib = cuda_library()
    dev = torch.device("cuda", 0)
    xd, sd, bd = x.float().to(dev).contiguous(), scale.float().to(dev), bias.float().to(dev)
    out = torch.full_like(xd, float("nan"))
    p = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_float))
    G, Cc, F = xd.shape
    rc = lib.jaqmc_b200_layernorm_fl(p(xd), p(sd), p(bd), p(out), G, Cc, F, C.c_float(EPS), kernel,
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return rc, out.double().cpu()


# (components, features, kernels that must support the shape)
SHAPES = [(1, 256, (1, 2, 3)), (5, 64, (1, 2)), (11, 128, (1, 2, 3)), (44, 256, (1, 2, 3)), (128, 256, (1, 2, 3)),
          (23, 512, (1, 2, 3)), (146, 256, (1, 2, 3)), (8, 96, (1, 2))]


@pytest.mark.gpu
@pytest.mark.parametrize("Cc,F,kernels", SHAPES, ids=["C%d_F%d" % s[:2] for s in SHAPES])
def test_layernorm_kernels_match_float64_rule(Cc, F, kernels):
    g = torch.Generator().manual_seed(Cc * 1000 + F)
    G = 5
    x = (torch.randn(G, Cc, F, generator=g, dtype=F64) + 0.5).float().double()
    if Cc > 1:
        x[:, 1:-1] *= 0.3
        x = x.float().double()
    scale = (1.0 + 0.2 * torch.randn(F, generator=g, dtype=F64)).float().double()
    bias = (0.1 * torch.randn(F, generator=g, dtype=F64)).float().double()
    if Cc > 1:
        ref = torch.stack([rule_f64(x[i], scale, bias, EPS) for i in range(G)])
    else:
        xc = x[:, 0] - x[:, 0].mean(-1, keepdim=True)
        ref = (xc * ((xc * xc).mean(-1, keepdim=True) + EPS) ** -0.5 * scale + bias)[:, None]
    outs = {}
    for kern in kernels + (0,):
        rc, out = _run(kern, x, scale, bias)
        assert rc == 0, (kern, rc)
        assert torch.isfinite(out).all(), kern
        rows = [("value", slice(0, 1), TOL_VJ)]
        if Cc > 1:
            rows += [("jacobian", slice(1, Cc - 1), TOL_VJ), ("laplacian", slice(Cc - 1, Cc), TOL_L)]
        errs = {nm: ((out[:, sl] - ref[:, sl]).abs().max() / ref[:, sl].abs().max()).item() for nm, sl, _ in rows}
        print("C=%d F=%d kernel=%d %s" % (Cc, F, kern, {a: "%.2e" % b for a, b in errs.items()}))
        for nm, _, tol in rows:
            assert errs[nm] < tol, (Cc, F, kern, nm, errs[nm])
        outs[kern] = out
    assert any(torch.equal(outs[0], outs[kern]) for kern in kernels)


@pytest.mark.gpu
def test_forced_kernel_rejects_unsupported_shapes():
    x = torch.randn(2, 5, 96, dtype=F64)
    rc, _ = _run(3, x, torch.ones(96, dtype=F64), torch.zeros(96, dtype=F64))
    assert rc != 0
