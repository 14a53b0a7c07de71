"""Sampling path on the CPU emulation build of the kernels: the periodic Metropolis-Hastings step
(``jaqmc_b200_mh_step_pbc``) against ``oracle.mh_update(wrap=...)`` (reference geometry/pbc.py:97-111,187-201 and
sampler/mcmc.py:96-137), and the stand-alone propose / accept building blocks."""

import numpy as np
import torch

import helpers as H
import test_emu_solid as S
from oracle import estimators as OE


def test_mh_step_pbc_matches_oracle_emu():
    rt = H.emu_runtime()
    W, Sn = 12, 3
    wf, sysh, el, logpsi_fn, (sim, cell_atoms, cell_charges), f32 = S._setup("fcc_lih_221", W, seed=3)
    lat = torch.as_tensor(sim, dtype=torch.float64)
    inv = torch.linalg.inv(lat)

    def wrap(x):
        fr = x @ inv
        return (fr - torch.floor(fr)) @ lat

    def blp(x):
        return torch.stack([2.0 * logpsi_fn(x[w]).real for w in range(x.shape[0])])

    g = torch.Generator().manual_seed(2)
    n = el.shape[1]
    normals = torch.randn(Sn, W, n, 3, generator=g, dtype=torch.float64).float()
    uniforms = torch.rand(Sn, W, generator=g, dtype=torch.float64).float().clamp_min(1e-7)
    stddev = 0.25
    sd32 = torch.tensor(stddev, dtype=torch.float32)
    x = torch.from_numpy(el).double()
    lp = blp(x)
    acc_ref, margin = [], []
    for s in range(Sn):
        # the oracle's update on the float32 proposal the kernel forms (step = float32 sum minus x, stddev 1)
        step = (x.float() + normals[s] * sd32).double() - x
        x_new, lp_new, cond, ratio = OE.mh_update(blp, x, lp, step, uniforms[s].double(), 1.0, wrap=wrap)
        acc_ref.append(cond)
        margin.append((ratio - torch.log(uniforms[s].double())).abs())
        x, lp = x_new, lp_new
    acc_ref, margin = torch.stack(acc_ref), torch.stack(margin)
    e32 = torch.from_numpy(el).contiguous().clone()
    logpsi = torch.empty(W)
    n_acc, accepted = rt.mh_step(wf, sysh, e32, logpsi, normals.contiguous(), uniforms.contiguous(),
                                 torch.tensor([stddev]), logpsi_valid=False, record_accepts=True, wrap_lattice=sim)
    accepted = accepted.bool()
    ok = torch.ones(W, dtype=torch.bool)
    for s in range(Sn):
        diff = (accepted[s] != acc_ref[s]) & ok
        assert (margin[s][diff] < 2e-5).all()
        ok &= ~diff
    assert ok.sum() >= W - 1
    assert int(n_acc) == int(accepted.sum())
    d = (e32.double() - x) @ inv
    d = d - torch.round(d)
    assert (d[ok] @ lat).abs().max() < 2e-5
    moved = accepted.any(dim=0)
    fr = e32.double()[moved] @ inv
    assert (fr > -1e-5).all() and (fr < 1 + 1e-5).all()
    np.testing.assert_allclose(logpsi[ok].numpy(), (0.5 * lp)[ok].numpy(), atol=5e-5)


def test_propose_and_accept_building_blocks():
    rt = H.emu_runtime()
    g = torch.Generator().manual_seed(0)
    W, n = 9, 4
    x1 = torch.randn(W, n, 3, generator=g)
    nrm = torch.randn(W, n, 3, generator=g)
    sd = torch.tensor([0.37])
    x2 = rt.mh_propose(x1, nrm, sd)
    np.testing.assert_allclose(x2.numpy(), (x1 + nrm * sd).numpy(), rtol=0, atol=2.5e-7)   # fma vs mul + add
    lp1 = torch.randn(W, generator=g)
    lp2 = torch.randn(W, generator=g)
    u = torch.rand(W, generator=g).clamp_min(1e-7)
    n_acc = torch.zeros(1)
    acc = torch.empty(W, dtype=torch.uint8)
    xa, la = x1.clone(), lp1.clone()
    rt.mh_accept(xa, x2, la, lp2, u, n_acc, acc)
    cond = (lp2 - lp1) > torch.log(u)
    assert torch.equal(acc.bool(), cond)
    assert torch.equal(xa, torch.where(cond[:, None, None], x2, x1))
    assert torch.equal(la, torch.where(cond, lp2, lp1))
    assert int(n_acc) == int(cond.sum())
