"""Periodic FermiNet (complex log psi) kernels + host orchestration on the CPU emulation build against the oracle."""

import numpy as np
import pytest
import torch

import helpers as H
from jaqmc_b200 import _marshal as M
from oracle import estimators as OE
from oracle import networks as ON

F64 = torch.float64


solid_system = H.solid_system


def _setup(kind, W, ndets=2, hs=(16, 16), hd=(8, 8), seed=0, device="cpu", distance_type="tri", sym_type="minimal"):
    prim, sim, patoms, cell_atoms, cell_charges, nspins, klist = solid_system(kind)
    n = sum(nspins)
    p64 = H.round_f32(ON.init_solid_params(nspins, patoms.shape[0], ndets, hs, hd, seed=seed + 21,
                                           distance_type=distance_type))
    g = np.random.default_rng(seed)
    el = (cell_atoms[g.integers(0, len(cell_atoms), (W, n))] + 0.8 * g.normal(size=(W, n, 3))).astype(np.float32)
    f32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).to(device)  # noqa: E731
    wf = M.solid_handle(H.to_f32(p64, device), nspins, patoms.shape[0], sim, prim, f32(klist), ndets, hs, hd,
                        distance_type=distance_type, sym_type=sym_type)
    sysh = M.system_handle(f32(patoms), None)
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))  # noqa: E731

    def logpsi(e):
        return ON.solid_logpsi(p64, e, t(patoms), nspins, t(sim), t(prim), t(klist), distance_type, sym_type)

    return wf, sysh, el, logpsi, (sim, cell_atoms, cell_charges), f32


def oracle_complex(logpsi, el):
    out = dict(logpsi=[], grad=[], lap=[], e_kin=[])
    for w in range(el.shape[0]):
        v, g, lap = OE.forward_laplacian(logpsi, torch.as_tensor(el[w].astype(np.float64)))
        out["logpsi"].append(complex(v))
        out["grad"].append(g.numpy())
        out["lap"].append(complex(lap))
        out["e_kin"].append(complex(-0.5 * lap - 0.5 * (g * g).sum()))
    return {k: np.asarray(v) for k, v in out.items()}


def check_solid(rt, kind, W, device="cpu", **kw):
    from jaqmc_b200.ewald import EwaldSum

    wf, sysh, el, logpsi, (sim, cell_atoms, cell_charges), f32 = _setup(kind, W, device=device, **kw)
    ew = EwaldSum(sim, device=device)
    e32 = torch.from_numpy(el).to(device)
    out = rt.local_energy_complex(wf, sysh, e32, ew, f32(cell_atoms), f32(cell_charges))
    out = {k: v.cpu().numpy() for k, v in out.items()}
    ref = oracle_complex(logpsi, el)
    gn = np.sqrt((np.abs(ref["grad"]) ** 2).sum(1))
    r = np.linalg.norm(el.reshape(W, -1), axis=1)
    l_scale = np.abs(ref["logpsi"].real) + gn * r
    dl = ref["logpsi"] - out["logpsi"]
    dphase = np.angle(np.exp(1j * dl.imag))  # phases agree modulo 2 pi
    assert np.max(np.abs(dl.real) / l_scale) < 1e-5 and np.median(np.abs(dl.real) / l_scale) < 1e-6
    assert np.max(np.abs(dphase) / (1.0 + gn * r)) < 1e-5
    gs = np.abs(ref["grad"]).max(axis=1, keepdims=True) + 1.0
    assert np.max(np.abs(out["grad"] - ref["grad"]) / gs) < 1e-3 and np.median(np.abs(out["grad"] - ref["grad"]) / gs) < 1e-5
    e_scale = 0.5 * np.abs(ref["lap"]) + 0.5 * (np.abs(ref["grad"]) ** 2).sum(1) + 1.0
    ek = np.abs(out["e_kin"] - ref["e_kin"]) / e_scale
    assert ek.max() < 1e-4 and np.median(ek) < 1e-5, ek
    ref_ew = OE.EwaldSum(sim)
    pot = np.array([OE.solid_potential_energy(ref_ew, el[w], cell_atoms, cell_charges) for w in range(W)])
    np.testing.assert_allclose(out["e_pot"], pot, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(out["e_loc"].real, out["e_kin"].real + out["e_pot"], rtol=1e-6, atol=1e-6)
    # value-only (sampling) path: real part and phase
    lp, ph = rt.logpsi(wf, sysh, e32)
    np.testing.assert_allclose(lp.cpu().numpy(), out["logpsi"].real, rtol=1e-5, atol=1e-5)
    assert np.max(np.abs(np.angle(np.exp(1j * (ph.cpu().numpy() - out["logpsi"].imag))))) < 1e-4
    return out


@pytest.mark.parametrize("kind", ["cubic_h2", "fcc_lih_221"])
def test_solid_local_energy_matches_oracle(kind):
    check_solid(H.emu_runtime(), kind, 4)


# geometry/pbc.py options: the polynomial `nu` distance (4 features per pair) and the over-complete direction sets of
# get_symmetry_lat (4 directions for fcc / hexagonal, 6 for bcc)
PBC_OPTIONS = [("cubic_h2", "nu", "minimal"), ("fcc_lih_221", "tri", "fcc"), ("fcc_lih_221", "nu", "bcc"),
               ("cubic_h2", "tri", "hexagonal"), ("fcc_lih_221", "nu", "fcc")]


@pytest.mark.parametrize("kind,distance_type,sym_type", PBC_OPTIONS)
def test_solid_distance_and_symmetry_options_match_oracle(kind, distance_type, sym_type):
    check_solid(H.emu_runtime(), kind, 3, distance_type=distance_type, sym_type=sym_type)
