"""Kernel arithmetic + host orchestration of the FermiNet path, checked on the CPU.

The kernels are compiled for the host by tests/emu/build_emu.py (block-stride emulation, test infrastructure
only) and driven through the product's own marshalling layer; the float64 oracle is the checker.
"""

import numpy as np
import pytest
import torch

import helpers as H
from jaqmc_b200 import _marshal as M
from oracle import networks as ON

CASES = {
    # name: (molecule, ndets, hidden_single, hidden_double, spin_split, envelope)
    "li_small": ("Li", 3, (16, 16, 16), (8, 8, 8), True, "abs_isotropic"),
    "lih_nosplit": ("LiH", 2, (16, 16), (4, 8), False, "abs_isotropic"),
    "h_single_channel": ("H", 2, (8, 8), (4, 4), True, "abs_isotropic"),
    "he_iso": ("He", 4, (12, 12, 12, 12), (6, 6, 6, 6), True, "isotropic"),
    "li_one_layer": ("Li", 2, (16,), (8,), True, "abs_isotropic"),
    "lih_null_envelope": ("LiH", 3, (16, 16), (8, 8), True, "null"),
    "ar_18_electrons": ("Ar", 2, (8, 8), (4, 4), True, "abs_isotropic"),
}


def _setup(case, W, seed=0):
    mol, ndets, hs, hd, split, env = CASES[case]
    atoms, charges, nspins = H.molecule(mol)
    p64 = H.round_f32(ON.init_ferminet_params(nspins, atoms.shape[0], ndets, hs, hd, seed=seed + 1))
    if not split or min(nspins) == 0:
        # non-split trees: single DenseGeneral_0 / _env
        p = p64["params"]
        if "SplitChannelDense_0" in p["orbital_layer"]:
            p["orbital_layer"] = {"DenseGeneral_0": p["orbital_layer"]["SplitChannelDense_0"]["DenseGeneral_0"]}
        if "_env_up" in p["envelope_layer"]:
            p["envelope_layer"] = {"_env": p["envelope_layer"]["_env_up"]}
    if env == "null":
        p64["params"]["envelope_layer"] = {}
    el = H.synthetic_walkers(atoms, charges, nspins, W, seed=seed)
    wf = M.ferminet_handle(H.to_f32(p64), nspins, atoms.shape[0], ndets, hs, hd, env, split)
    sysh = M.system_handle(atoms.float(), charges.float())

    def oracle_fn(e):
        return ON.ferminet_logpsi(p64, e, atoms, nspins, env)

    return wf, sysh, el, atoms, charges, nspins, oracle_fn


@pytest.mark.parametrize("case", list(CASES))
def test_local_energy_matches_oracle(case):
    rt = H.emu_runtime()
    W = 2 if case.startswith("ar") else 5
    wf, sysh, el, atoms, charges, nspins, fn = _setup(case, W)
    out = rt.local_energy(wf, sysh, el.float().contiguous())
    ref = H.oracle_batch(fn, el, atoms, charges, track=True)
    assert np.array_equal(out["sign"].numpy(), ref["sign"])
    H.assert_fp32_parity({k: v.numpy() for k, v in out.items()}, ref, el)
    np.testing.assert_allclose(out["e_pot"].numpy(), ref["e_pot"], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(out["e_loc"].numpy(), out["e_kin"].numpy() + out["e_pot"].numpy(), rtol=1e-6)
    np.testing.assert_allclose(out["e_kin"].numpy(),
                               -0.5 * (out["lap"].numpy() + (out["grad"].numpy() ** 2).sum(-1)), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("case", ["li_small", "lih_nosplit", "h_single_channel"])
def test_logpsi_value_path_matches_tracked_path(case):
    rt = H.emu_runtime()
    wf, sysh, el, atoms, charges, nspins, fn = _setup(case, 6, seed=3)
    e32 = el.float().contiguous()
    lp, sg = rt.logpsi(wf, sysh, e32)
    out = rt.local_energy(wf, sysh, e32)
    ref = H.oracle_batch(fn, el, track=False)
    assert np.array_equal(sg.numpy(), ref["sign"])
    np.testing.assert_allclose(lp.numpy(), ref["logpsi"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(lp.numpy(), out["logpsi"].numpy(), rtol=0, atol=2e-5)


def test_walker_tiling_gives_identical_results():
    """A workspace too small for the batch makes the library tile the walker axis; results are bit-identical."""
    rt = H.emu_runtime()
    wf, sysh, el, *_ = _setup("li_small", 7, seed=5)
    e32 = el.float().contiguous()
    full = rt.local_energy(wf, sysh, e32)
    need1 = rt.workspace_bytes(wf, 1, True)
    rt2 = H.Runtime(rt.lib, "cpu", workspace_limit_bytes=int(need1 * 2.5), _emulation=True)
    tiled = rt2.local_energy(wf, sysh, e32)
    for k in full:
        assert torch.equal(full[k], tiled[k]), k
    rt3 = H.Runtime(rt.lib, "cpu", workspace_limit_bytes=1 << 10, _emulation=True)
    rt3.workspace = lambda nbytes: torch.empty(1 << 10, dtype=torch.uint8)
    with pytest.raises(H._abi.JaqmcB200Error) as ei:
        rt3.local_energy(wf, sysh, e32)
    assert ei.value.code == H._abi.ERR_WORKSPACE_TOO_SMALL


def test_partial_sums():
    rt = H.emu_runtime()
    wf, sysh, el, *_ = _setup("li_small", 9, seed=7)
    sums = torch.zeros(3)
    out = rt.local_energy(wf, sysh, el.float().contiguous(), sums=sums)
    e = out["e_loc"].double()
    assert np.isclose(float(sums[0]), float(e.sum()), rtol=1e-5)
    assert np.isclose(float(sums[1]), float((e * e).sum()), rtol=1e-5)
    assert float(sums[2]) == 9.0


def test_hydrogen_atom_closed_form():
    """``log psi = alpha |r|``: E_L = -alpha^2/2 - (alpha + 1)/|r| (hydrogen closed form; exact ground state -0.5 at
    alpha = -1, reference tests/hydrogen/atom_test.py:47-52)."""
    from oracle import estimators as OE

    rt = H.emu_runtime()
    g = torch.Generator().manual_seed(0)
    el = torch.randn(16, 1, 3, generator=g)
    sysh = M.system_handle(torch.zeros(1, 3), torch.ones(1))
    for alpha in (-0.8, -1.0):
        wf = M.hydrogen_handle({"params": {"alpha": torch.tensor([alpha])}})
        out = rt.local_energy(wf, sysh, el.contiguous())
        ref = OE.hydrogen_local_energy(alpha, el.double())
        np.testing.assert_allclose(out["e_loc"].numpy(), ref.numpy(), rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(out["logpsi"].numpy(), alpha * el.norm(dim=-1).squeeze(-1).numpy(), rtol=1e-6)
        assert (out["sign"] == 1).all()
    assert np.allclose(out["e_loc"].numpy(), -0.5, atol=1e-5)
