"""Edge spin configurations of the reference's wavefunction tests (tests/wavefunction/molecule_wavefunction_test.py:148-190):
all electrons in one spin channel, either one -- (1, 0), (0, 2), (3, 0) -- through the CUDA pipeline (CPU emulation build here,
GPU in test_gpu_edge_spins.py) against the float64 oracle."""

import numpy as np
import pytest
import torch

import helpers as H
from jaqmc_b200 import _marshal as M
from oracle import networks as ON

F64 = torch.float64
CASES = [((1, 0), 1.0), ((0, 2), 2.0), ((3, 0), 3.0), ((0, 1), 1.0)]


def setup(nspins, charge, kind, device="cpu", W=4, seed=0, ndets=3):
    atoms = torch.zeros(1, 3, dtype=F64)
    charges = torch.tensor([charge], dtype=F64)
    el = H.synthetic_walkers(atoms, charges, nspins, W, seed=seed)
    if kind == "ferminet":
        hs, hd = (16, 16), (8, 8)
        p64 = H.round_f32(ON.init_ferminet_params(nspins, 1, ndets, hs, hd, seed=seed + 1))
        wf = M.ferminet_handle(H.to_f32(p64, device), nspins, 1, ndets, hs, hd, "abs_isotropic", True)
        fn = lambda e: ON.ferminet_logpsi(p64, e, atoms, nspins)  # noqa: E731
    elif kind == "psiformer":
        p64 = H.round_f32(ON.init_psiformer_params(nspins, 1, 3, 2, 2, 8, (16,), seed=seed + 5))
        wf = M.psiformer_handle(H.to_f32(p64, device), nspins, 1, 3, 2, 2, 8, (16,), "pre")
        fn = lambda e: ON.psiformer_logpsi(p64, e, atoms, nspins, "pre")  # noqa: E731
    else:
        p64 = H.round_f32(ON.init_lapnet_params(nspins, 1, 3, 2, 2, 8, 2, seed=seed + 3))
        wf = M.lapnet_handle(H.to_f32(p64, device), nspins, 1, 3, 2, 2, 8, 2)
        fn = lambda e: ON.lapnet_logpsi(p64, e, atoms, nspins, 2)  # noqa: E731
    sysh = M.system_handle(atoms.float().to(device), charges.float().to(device))
    return wf, sysh, el, atoms, charges, fn


def check(rt, nspins, charge, kind, device="cpu", ndets=3):
    wf, sysh, el, atoms, charges, fn = setup(nspins, charge, kind, device, ndets=ndets)
    e32 = el.float().contiguous().to(device)
    out = {k: v.cpu().numpy() for k, v in rt.local_energy(wf, sysh, e32).items()}
    ref = H.oracle_batch(fn, el, atoms, charges)
    assert np.array_equal(out["sign"], ref["sign"])
    H.assert_fp32_parity(out, ref, el)
    lp, sg = rt.logpsi(wf, sysh, e32)
    assert np.array_equal(sg.cpu().numpy(), ref["sign"])
    _, l_scale = H.fp32_scales(ref, el)
    assert (np.abs(lp.cpu().numpy() - ref["logpsi"]) / l_scale).max() < 1e-5


@pytest.mark.parametrize("kind", ["ferminet", "psiformer", "lapnet"])
@pytest.mark.parametrize("nspins,charge", CASES, ids=["%d_%d" % c[0] for c in CASES])
def test_single_channel_spin_configurations(nspins, charge, kind):
    check(H.emu_runtime(), nspins, charge, kind)
