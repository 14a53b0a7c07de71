"""GPU tests of the parameter-gradient path: ``jaqmc_b200_ferminet_logpsi_vjp`` against ``torch.autograd`` through the
float64 oracle (FermiNet-N2 at the default widths included: the reverse dense layers run on the tcgen05 kernels), and
``LossAndGrad`` against the oracle's restatement of the reference's reduce / finalize (estimator/loss_grad.py:96-128)."""

import numpy as np
import pytest
import torch

import helpers as H
import test_emu_gradients as E
from oracle import estimators as OE
from oracle import networks as ON

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _rt():
    from jaqmc_b200._runtime import runtime

    return runtime(DEV)


@pytest.mark.parametrize("mol,ndets,hs,hd,kw", [
    ("Li", 3, (16, 16, 16), (8, 8, 8), {}),
    ("LiH", 4, (64, 64, 64), (32, 32, 32), {}),
    ("H", 2, (8, 8), (4, 4), {}),
    ("LiH", 3, (32, 32), (8, 8), dict(envelope="null")),
    ("Ar", 4, (32, 32), (8, 8), {}),        # 18 electrons: warp-per-matrix inversion
    ("Zn", 2, (32, 32), (8, 8), dict(tol=5e-4)),   # 30 electrons (measured 1.1e-4; N2 below uses the same bound)
])
def test_vjp_matches_autograd_small(mol, ndets, hs, hd, kw):
    E.check_vjp(_rt(), mol, ndets, hs, hd, 6, device=DEV, **kw)


def test_vjp_matches_autograd_n2_full_network():
    """The headline network (N2, 256 x 4 / 32 x 4, 16 determinants)."""
    E.check_vjp(_rt(), "N2", 16, (256,) * 4, (32,) * 4, 6, device=DEV, tol=5e-4)


def test_vjp_is_deterministic_and_linear_in_the_cotangent():
    """Size-independent properties at 4096 walkers: bit-identical repeats; vjp(a c1 + b c2) = a vjp(c1) + b vjp(c2)."""
    from jaqmc_b200.data import MoleculeData
    from jaqmc_b200.wavefunction import FermiNetWavefunction

    atoms, charges, nspins = H.molecule("Li")
    wf = FermiNetWavefunction(nspins=nspins, ndets=16)
    W = 4096
    el = H.synthetic_walkers(atoms, charges, nspins, W, seed=4).float().to(DEV)
    data = MoleculeData(el, atoms.float().to(DEV), charges.float().to(DEV))
    params = wf.init_params(data, 3)
    g = torch.Generator(device=DEV).manual_seed(0)
    c1 = torch.randn(W, generator=g, device=DEV) / W
    c2 = torch.randn(W, generator=g, device=DEV) / W
    g1, lp1 = wf.logpsi_vjp(params, data, c1)
    g1b, lp1b = wf.logpsi_vjp(params, data, c1)
    g2, _ = wf.logpsi_vjp(params, data, c2)
    g12, _ = wf.logpsi_vjp(params, data, 0.5 * c1 - 2.0 * c2)
    assert torch.equal(lp1, lp1b)
    dl = (lp1 - wf.logpsi(params, data)).abs()      # a different inversion kernel than the sampling path's
    assert dl.median() < 1e-5 and dl.quantile(0.99) < 1e-3
    for a, b, c, d in zip(ON.tree_leaves(g1), ON.tree_leaves(g1b), ON.tree_leaves(g2), ON.tree_leaves(g12)):
        assert torch.equal(a, b)
        assert torch.isfinite(a).all()
        want = 0.5 * a - 2.0 * c
        scale = want.abs().max() + 1e-12
        assert ((d - want).abs().max() / scale) < 2e-4


def test_loss_and_grad_matches_reference_formula():
    """``LossAndGrad`` (MAD clipping, the default) against ``oracle.loss_and_grad`` fed with per-walker autograd scores."""
    from jaqmc_b200.data import MoleculeData
    from jaqmc_b200.estimator import LossAndGrad
    from jaqmc_b200.wavefunction import FermiNetWavefunction

    atoms, charges, nspins = H.molecule("LiH")
    hs, hd, ndets = [32, 32], [8, 8], 4
    wf = FermiNetWavefunction(nspins=nspins, ndets=ndets, hidden_dims_single=hs, hidden_dims_double=hd)
    p64 = H.round_f32(ON.init_ferminet_params(nspins, atoms.shape[0], ndets, hs, hd, seed=2))
    W, pool = 24, 40
    el = H.synthetic_walkers(atoms, charges, nspins, pool, seed=6)
    params = H.to_f32(p64, DEV)
    leaves = ON.tree_leaves(p64)
    for t in leaves:
        t.requires_grad_(True)
    scores = [[] for _ in leaves]
    for w in range(pool):
        _, lp = ON.ferminet_logpsi(p64, el[w], atoms, nspins)
        gs = torch.autograd.grad(lp, leaves)
        for k, g_ in enumerate(gs):
            scores[k].append(g_)
    for t in leaves:
        t.requires_grad_(False)
    scores = [torch.stack(s) for s in scores]
    # random (not equilibrated) walkers: some sit next to a node of psi, where the score is huge and ill-conditioned
    # (float32 inversion of a near-singular orbital matrix); keep the W best-conditioned ones of the pool
    keep = sum(s.reshape(pool, -1).abs().amax(dim=1) for s in scores).argsort()[:W]
    el = el[keep]
    scores = [s[keep] for s in scores]
    data = MoleculeData(el.float().to(DEV), atoms.float().to(DEV), charges.float().to(DEV))
    out = wf.local_energy(params, data)
    # the outlier the clipping has to catch goes on a well-conditioned walker (smallest score): a walker next to a
    # node has a huge, ill-conditioned score whose float32 error would dominate the comparison once it carries the
    # largest weight
    size = sum(s.reshape(W, -1).abs().amax(dim=1) for s in scores)
    e_loc = out["e_loc"].clone()
    e_loc[int(size.argmin())] += 500.0
    for method, scale in (("mad", 5.0), ("iqr", 1.5), ("none", 1.0)):
        got = LossAndGrad(f_log_psi=wf, clip_method=method, clip_scale=scale).evaluate(params, data, {"total_energy": e_loc})
        loss, want = OE.loss_and_grad(scores, e_loc.cpu().double(), method, scale)
        assert abs(float(got["loss"]) - float(loss)) < 1e-4 * abs(float(loss))
        for g_, w_ in zip(ON.tree_leaves(got["grads"]), want):
            sc = float(w_.abs().max()) + 1e-12
            assert float((g_.cpu().double() - w_).abs().max()) / sc < 2e-3, (method, sc)
    with pytest.raises(ValueError):
        LossAndGrad(f_log_psi=wf).evaluate(params, data, {"total_energy": torch.zeros(W, 2, device=DEV)})
