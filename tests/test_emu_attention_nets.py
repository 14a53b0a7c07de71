"""LapNet and Psiformer kernel arithmetic + host orchestration, checked on the CPU against the float64 oracle
(host emulation build of the CUDA sources, tests/emu; test infrastructure only)."""

import numpy as np
import pytest
import torch

import helpers as H
from jaqmc_b200 import _marshal as M
from oracle import networks as ON

LAPNET_CASES = {
    # name: (molecule, ndets, layers, heads, head_dim, local_updates, jastrow)
    "li": ("Li", 3, 2, 2, 8, 2, True),
    "lih_nojastrow": ("LiH", 2, 3, 2, 4, 1, False),
    "h_single_channel": ("H", 2, 2, 1, 8, 2, True),
    "he_one_layer": ("He", 2, 1, 2, 4, 2, True),
}
PSIFORMER_CASES = {
    # name: (molecule, ndets, layers, heads, head_dim, mlp_hidden, layer_norm_mode, jastrow)
    "li_pre": ("Li", 3, 2, 2, 8, (16,), "pre", True),
    "lih_post": ("LiH", 2, 2, 2, 4, (12,), "post", True),
    "he_null_two_hidden": ("He", 2, 1, 2, 4, (8, 12), "null", False),
    "h_single_channel": ("H", 2, 2, 1, 8, (8,), "pre", True),
}


def _strip(p64, nspins):
    """Trees of single-channel systems hold one DenseGeneral_0 / _env (orbital.py:73, envelope.py:82-96)."""
    return p64


def _lapnet(case, W, seed=0):
    mol, ndets, layers, heads, dh, nlu, jas = LAPNET_CASES[case]
    atoms, charges, nspins = H.molecule(mol)
    p64 = H.round_f32(ON.init_lapnet_params(nspins, atoms.shape[0], ndets, layers, heads, dh, nlu, seed=seed + 3,
                                            jastrow=jas))
    el = H.synthetic_walkers(atoms, charges, nspins, W, seed=seed)
    wf = M.lapnet_handle(H.to_f32(p64), nspins, atoms.shape[0], ndets, layers, heads, dh, nlu, jastrow=jas)
    sysh = M.system_handle(atoms.float(), charges.float())
    return wf, sysh, el, atoms, charges, (lambda e: ON.lapnet_logpsi(p64, e, atoms, nspins, heads))


def _psiformer(case, W, seed=0):
    mol, ndets, layers, heads, dh, mlp, lnm, jas = PSIFORMER_CASES[case]
    atoms, charges, nspins = H.molecule(mol)
    p64 = H.round_f32(ON.init_psiformer_params(nspins, atoms.shape[0], ndets, layers, heads, dh, mlp, seed=seed + 5,
                                               jastrow=jas))
    el = H.synthetic_walkers(atoms, charges, nspins, W, seed=seed)
    wf = M.psiformer_handle(H.to_f32(p64), nspins, atoms.shape[0], ndets, layers, heads, dh, mlp, lnm, jastrow=jas)
    sysh = M.system_handle(atoms.float(), charges.float())
    return wf, sysh, el, atoms, charges, (lambda e: ON.psiformer_logpsi(p64, e, atoms, nspins, layer_norm_mode=lnm))


def _check(rt, wf, sysh, el, atoms, charges, fn):
    e32 = el.float().contiguous()
    out = {k: v.numpy() for k, v in rt.local_energy(wf, sysh, e32).items()}
    ref = H.oracle_batch(fn, el, atoms, charges, track=True)
    assert np.array_equal(out["sign"], ref["sign"])
    H.assert_fp32_parity(out, ref, el)
    np.testing.assert_allclose(out["e_kin"], -0.5 * (out["lap"] + (out["grad"] ** 2).sum(-1)), rtol=1e-5, atol=1e-5)
    lp, sg = rt.logpsi(wf, sysh, e32)  # value-only path (the MCMC forward)
    assert np.array_equal(sg.numpy(), ref["sign"])
    np.testing.assert_allclose(lp.numpy(), out["logpsi"], rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize("case", list(LAPNET_CASES))
def test_lapnet_local_energy_matches_oracle(case):
    _check(H.emu_runtime(), *_lapnet(case, 4))


@pytest.mark.parametrize("case", list(PSIFORMER_CASES))
def test_psiformer_local_energy_matches_oracle(case):
    _check(H.emu_runtime(), *_psiformer(case, 4))
