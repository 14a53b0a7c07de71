"""Parameter-gradient path (``jaqmc_b200_ferminet_logpsi_vjp``; reference estimator/loss_grad.py:70-128) on the CPU
emulation build against ``torch.autograd`` through the float64 oracle: for a per-walker cotangent c the kernels must
return  sum_w c_w d log|psi|(x_w) / d theta  for every leaf of the parameter tree."""

import numpy as np
import pytest
import torch

import helpers as H
from jaqmc_b200 import _marshal as M
from oracle import networks as ON


def oracle_vjp(p64, el, atoms, nspins, cot, envelope="abs_isotropic"):
    leaves = ON.tree_leaves(p64)
    for t in leaves:
        t.requires_grad_(True)
    total = 0.0
    lps = []
    for w in range(el.shape[0]):
        _, lp = ON.ferminet_logpsi(p64, el[w], atoms, nspins, envelope)
        total = total + cot[w] * lp
        lps.append(float(lp))
    grads = torch.autograd.grad(total, leaves, allow_unused=True)
    for t in leaves:
        t.requires_grad_(False)
    return [g if g is not None else torch.zeros_like(t) for g, t in zip(grads, leaves)], np.asarray(lps)


def check_vjp(rt, mol, ndets, hs, hd, W, device="cpu", envelope="abs_isotropic", split=True, seed=0, tol=2e-4):
    atoms, charges, nspins = H.molecule(mol)
    p64 = H.round_f32(ON.init_ferminet_params(nspins, atoms.shape[0], ndets, hs, hd, seed=seed + 1))
    if envelope == "null":
        p64["params"]["envelope_layer"] = {}
    if not split:
        p = p64["params"]
        p["orbital_layer"] = {"DenseGeneral_0": p["orbital_layer"]["SplitChannelDense_0"]["DenseGeneral_0"]}
        if "_env_up" in p["envelope_layer"]:
            p["envelope_layer"] = {"_env": p["envelope_layer"]["_env_up"]}
    el = H.synthetic_walkers(atoms, charges, nspins, W, seed=seed)
    g = torch.Generator().manual_seed(seed + 5)
    cot = torch.randn(W, generator=g, dtype=torch.float64).float().double()
    want, lp_ref = oracle_vjp(p64, el, atoms, nspins, cot, envelope)
    p32 = H.to_f32(p64, device)
    wf = M.ferminet_handle(p32, nspins, atoms.shape[0], ndets, hs, hd, envelope, split)
    grads = ON.tree_map(lambda t: torch.zeros_like(t), p32)
    gh = M.ferminet_handle(grads, nspins, atoms.shape[0], ndets, hs, hd, envelope, split)
    sysh = M.system_handle(atoms.float().to(device), None)
    lp, sg = rt.ferminet_logpsi_vjp(wf, gh, sysh, el.float().contiguous().to(device), cot.float().to(device))
    np.testing.assert_allclose(lp.cpu().numpy(), lp_ref, rtol=5e-6, atol=3e-5)
    got = ON.tree_leaves(grads)
    names = [k for k, _ in sorted(_flat(p64["params"]).items())]
    assert len(got) == len(want) == len(names)
    worst = (0.0, "")
    for nm, g_, w_ in zip(names, got, want):
        scale = float(w_.abs().max()) + 1e-12
        err = float((g_.cpu().double() - w_).abs().max()) / scale
        worst = max(worst, (err, nm))
        assert err < tol, (nm, err, scale)
    print("vjp %s %s: largest leaf error %.2e (%s), tolerance %.0e" % (mol, tuple(hs), worst[0], worst[1], tol))
    return grads


def _flat(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(_flat(v, f"{prefix}{k}/"))
        else:
            out[f"{prefix}{k}"] = v
    return out


@pytest.mark.parametrize("mol,ndets,hs,hd,kw", [
    ("Li", 3, (16, 16, 16), (8, 8, 8), {}),                       # residual layers in both streams
    ("LiH", 2, (16, 12), (4, 8), dict(split=False)),               # no residual (widths change), shared orbitals / envelope
    ("H", 2, (8, 8), (4, 4), {}),                                  # one spin channel
    ("He", 4, (12, 12, 12), (6, 6, 6), dict(envelope="isotropic")),
    ("LiH", 3, (16, 16), (8, 8), dict(envelope="null")),
    ("Li", 2, (16,), (8,), {}),                                    # one layer: no two-electron layer at all
    ("Ar", 2, (16, 16), (8, 8), {}),                               # 18 electrons: the warp-per-matrix inversion (17 ... 32)
])
def test_ferminet_logpsi_vjp_matches_autograd(mol, ndets, hs, hd, kw):
    check_vjp(H.emu_runtime(), mol, ndets, hs, hd, 5, **kw)
