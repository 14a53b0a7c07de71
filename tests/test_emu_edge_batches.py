"""Ragged / degenerate walker batches through the C ABI (CPU emulation build): an empty batch, one walker, odd counts.
A walker's result must not depend on which batch it was evaluated in (the reference vmaps one-walker functions)."""

import numpy as np
import pytest
import torch

import helpers as H
from jaqmc_b200 import _marshal as M
from oracle import networks as ON


def _ferminet(device="cpu", hs=(16, 16), hd=(8, 8), ndets=3, mol="LiH"):
    atoms, charges, nspins = H.molecule(mol)
    p = H.to_f32(H.round_f32(ON.init_ferminet_params(nspins, atoms.shape[0], ndets, hs, hd, seed=4)), device)
    wf = M.ferminet_handle(p, nspins, atoms.shape[0], ndets, hs, hd, "abs_isotropic", True)
    sysh = M.system_handle(atoms.float().to(device), charges.float().to(device))
    return wf, sysh, atoms, charges, nspins


def check_sub_batches(rt, wf, sysh, el, sizes, exact=True):
    full = {k: v.cpu().numpy() for k, v in rt.local_energy(wf, sysh, el).items()}
    lp_full, sg_full = (t.cpu().numpy() for t in rt.logpsi(wf, sysh, el))
    for w in sizes:
        part = {k: v.cpu().numpy() for k, v in rt.local_energy(wf, sysh, el[:w].contiguous()).items()}
        lp, sg = (t.cpu().numpy() for t in rt.logpsi(wf, sysh, el[:w].contiguous()))
        for k, v in part.items():
            assert v.shape[0] == w, (k, v.shape)
            if exact:
                assert np.array_equal(v, full[k][:w]), (w, k)
            else:
                np.testing.assert_allclose(v, full[k][:w], rtol=1e-6, atol=1e-6)
        assert np.array_equal(sg, sg_full[:w])
        assert np.array_equal(lp, lp_full[:w]) if exact else np.allclose(lp, lp_full[:w], rtol=1e-6, atol=1e-6)


def test_sub_batches_give_identical_walkers():
    rt = H.emu_runtime()
    wf, sysh, atoms, charges, nspins = _ferminet()
    el = H.synthetic_walkers(atoms, charges, nspins, 9, seed=2).float().contiguous()
    check_sub_batches(rt, wf, sysh, el, (0, 1, 2, 5, 9))


def test_empty_batch_mh_step_is_a_no_op():
    rt = H.emu_runtime()
    wf, sysh, atoms, charges, nspins = _ferminet()
    n = sum(nspins)
    el = torch.zeros(0, n, 3)
    lp = torch.zeros(0)
    n_acc, _ = rt.mh_step(wf, sysh, el, lp, torch.zeros(3, 0, n, 3), torch.zeros(3, 0), torch.full((1,), 0.1), logpsi_valid=False)
    assert float(n_acc) == 0.0
