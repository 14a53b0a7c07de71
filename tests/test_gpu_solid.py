"""GPU parity tests of the periodic FermiNet (complex log psi) and the Ewald potential against the float64 oracle."""

import numpy as np
import pytest
import torch

import test_emu_solid as S

pytestmark = pytest.mark.gpu


def _rt():
    from jaqmc_b200._runtime import runtime

    return runtime(torch.device("cuda", 0))


@pytest.mark.parametrize("kind,kw", [
    ("cubic_h2", {}),
    ("fcc_lih_221", {}),
    ("fcc_lih_221", dict(ndets=16, hs=(256,) * 4, hd=(32,) * 4)),   # default network widths (tcgen05 dense path)
])
def test_solid_local_energy_matches_oracle(kind, kw):
    out = S.check_solid(_rt(), kind, 6, device="cuda", **kw)
    assert np.isfinite(out["e_loc"].real).all()


@pytest.mark.parametrize("kind,distance_type,sym_type", S.PBC_OPTIONS)
def test_solid_distance_and_symmetry_options_match_oracle(kind, distance_type, sym_type):
    """geometry/pbc.py `nu` distance and fcc / bcc / hexagonal direction sets on the CUDA kernels."""
    out = S.check_solid(_rt(), kind, 6, device="cuda", distance_type=distance_type, sym_type=sym_type)
    assert np.isfinite(out["e_loc"].real).all()


def test_solid_nu_bcc_default_widths():
    """Default network widths with the 4-feature `nu` inputs (tcgen05 dense layers, K = 4 pair layer kernels)."""
    out = S.check_solid(_rt(), "fcc_lih_221", 4, device="cuda", ndets=16, hs=(256,) * 4, hd=(32,) * 4,
                        distance_type="nu", sym_type="bcc")
    assert np.isfinite(out["e_loc"].real).all()


def test_solid_parity_lih_222_full_network():
    """BASELINE config 5 at its full per-walker size: LiH rock salt 2x2x2 (32 electrons, 16 atoms, 98 components per
    group, 512-column complex orbital layers), default network widths."""
    out = S.check_solid(_rt(), "fcc_lih_222", 3, device="cuda", ndets=16, hs=(256,) * 4, hd=(32,) * 4)
    assert np.isfinite(out["e_loc"].real).all()


def test_solid_full_batch_properties():
    """4096 walkers: finite energies, invariance of psi under a simulation-lattice translation of one electron
    (reference tests/wavefunction/solid_test.py:112-137), determinism."""
    from jaqmc_b200.ewald import EwaldSum

    rt = _rt()
    W = 4096
    wf, sysh, el, logpsi, (sim, cell_atoms, cell_charges), f32 = S._setup("fcc_lih_221", W, device="cuda")
    ew = EwaldSum(sim, device="cuda")
    e32 = torch.from_numpy(el).cuda()
    out = rt.local_energy_complex(wf, sysh, e32, ew, f32(cell_atoms), f32(cell_charges))
    assert torch.isfinite(out["e_loc"].real).all()
    shifted = e32.clone()
    shifted[:, 1] += torch.as_tensor(sim[0] - sim[1], dtype=torch.float32, device="cuda")
    out_s = rt.local_energy_complex(wf, sysh, shifted.contiguous(), ew, f32(cell_atoms), f32(cell_charges))
    d = out_s["logpsi"] - out["logpsi"]
    scale = out["logpsi"].real.abs() + out["grad"].abs().pow(2).sum(1).sqrt() * e32.reshape(W, -1).norm(dim=1)
    assert (d.real.abs() / scale).median() < 2e-6 and (d.real.abs() / scale).max() < 1e-4
    assert torch.allclose(out_s["e_pot"], out["e_pot"], rtol=1e-4, atol=1e-3)
    out2 = rt.local_energy_complex(wf, sysh, e32, ew, f32(cell_atoms), f32(cell_charges))
    for k in out:
        assert torch.equal(torch.view_as_real(out[k]) if out[k].is_complex() else out[k],
                           torch.view_as_real(out2[k]) if out2[k].is_complex() else out2[k]), k
