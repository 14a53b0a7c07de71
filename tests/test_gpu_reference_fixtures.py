"""The CUDA kernels, through the public host classes, against the golden vectors produced by the REFERENCE's own source
(``tests/golden/ref_*.npz``, see tests/test_reference_fixtures.py and scripts/make_reference_fixtures.py).
float32 kernels vs float64 reference values: tolerances as in tests/helpers.py (scaled north-star tolerances), sign
bit-exact, MH decisions identical outside the 1e-6 margin."""

import numpy as np
import pytest
import torch

import helpers as H
import test_reference_fixtures as R
from jaqmc_b200.data import MoleculeData, SolidData

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _wavefunction(meta):
    from jaqmc_b200 import wavefunction as WF

    kw = dict(meta["kwargs"])
    nspins = tuple(meta["nspins"])
    cls = {"ferminet": WF.FermiNetWavefunction, "lapnet": WF.LapNetWavefunction, "psiformer": WF.PsiformerWavefunction}[meta["kind"]]
    return cls(nspins=nspins, **kw)


@pytest.mark.parametrize("name", R.MOLECULE)
def test_kernels_match_reference_molecule(name):
    meta, p64, z = R.load(name)
    wf = _wavefunction(meta)
    params = H.to_f32(p64, DEV)
    f32 = lambda k: torch.from_numpy(z[k].astype(np.float32)).to(DEV)  # noqa: E731
    data = MoleculeData(f32("electrons"), f32("atoms"), f32("charges"))
    out = {k: v.cpu().numpy() for k, v in wf.local_energy(params, data).items()}
    ref = {k: z[k] for k in ("logpsi", "sign", "grad", "lap", "e_kin", "e_pot")}
    assert np.array_equal(out["sign"], ref["sign"])
    H.assert_fp32_parity(out, ref, z["electrons"])
    np.testing.assert_allclose(out["e_pot"], ref["e_pot"], rtol=3e-6)
    ev = wf.evaluate(params, data)
    assert np.array_equal(ev["sign_logpsi"].cpu().numpy(), ref["sign"])
    _, l_scale = H.fp32_scales(ref, z["electrons"])
    assert (np.abs(ev["logpsi"].cpu().numpy() - ref["logpsi"]) / l_scale).max() < 1e-5
    orb = wf.orbitals(params, data).cpu().numpy()           # (W, ndets, n, n): the pretraining head (N2)
    scale = np.abs(z["orbitals"]).max(axis=(-1, -2), keepdims=True)
    assert (np.abs(orb - z["orbitals"]) / scale).max() < 2e-5


@pytest.mark.parametrize("name", R.SOLID)
def test_kernels_match_reference_solid(name):
    from jaqmc_b200.ewald import EwaldSum
    from jaqmc_b200.wavefunction import SolidWavefunction

    meta, p64, z = R.load(name)
    wf = SolidWavefunction(nspins=tuple(meta["nspins"]), simulation_lattice=z["sim_lattice"],
                           primitive_lattice=z["prim_lattice"], klist=z["klist"], **meta["kwargs"])
    f32 = lambda k: torch.from_numpy(z[k].astype(np.float32)).to(DEV)  # noqa: E731
    data = SolidData(f32("electrons"), f32("cell_atoms"), f32("cell_charges"), f32("prim_atoms"))
    out = wf.local_energy(H.to_f32(p64, DEV), data, ewald=EwaldSum(z["sim_lattice"], device=DEV))
    out = {k: v.cpu().numpy() for k, v in out.items()}
    W = z["electrons"].shape[0]
    gn = np.sqrt((np.abs(z["grad"]) ** 2).sum(1))
    r = np.linalg.norm(z["electrons"].reshape(W, -1), axis=1)
    dl = z["logpsi"] - out["logpsi"]
    assert np.max(np.abs(dl.real) / (np.abs(z["logpsi"].real) + gn * r)) < 1e-5
    assert np.max(np.abs(np.angle(np.exp(1j * dl.imag))) / (1.0 + gn * r)) < 1e-5
    e_scale = 0.5 * np.abs(z["lap"]) + 0.5 * gn ** 2 + 1.0
    assert (np.abs(out["e_kin"] - z["e_kin"]) / e_scale).max() < 1e-4
    np.testing.assert_allclose(out["e_pot"], z["e_pot"], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("name", ["mcmc_lih", "mcmc_pbc"])
def test_mh_decisions_match_reference(name):
    """``MCMCSampler.step`` through ``SamplePlan`` on the reference's noise: accept decisions identical to the
    reference's ``_mh_update`` except where ``|dlog p - log u| < 1e-6`` (recomputed from the fixture's log-probabilities),
    pmove and adapted width equal when no decision differs."""
    from jaqmc_b200.sampler import BatchLogProb, MCMCSampler, SamplePlan, make_pbc_gaussian_proposal
    from jaqmc_b200.wavefunction import FermiNetWavefunction, SolidWavefunction

    meta, p64, z = R.load(name)
    f32 = lambda k: torch.from_numpy(z[k].astype(np.float32)).to(DEV)  # noqa: E731
    S = meta["steps"]
    if meta["kind"] == "mcmc":
        wf = FermiNetWavefunction(nspins=tuple(meta["nspins"]), **meta["kwargs"])
        data = MoleculeData(f32("electrons"), f32("atoms"), f32("charges"))
        sampler = MCMCSampler(steps=S, adapt_frequency=meta["adapt_frequency"])
    else:
        wf = SolidWavefunction(nspins=tuple(meta["nspins"]), simulation_lattice=z["sim_lattice"],
                               primitive_lattice=z["prim_lattice"], klist=z["klist"], **meta["kwargs"])
        data = SolidData(f32("electrons"), f32("cell_atoms"), f32("cell_charges"), f32("prim_atoms"))
        sampler = MCMCSampler(steps=S, adapt_frequency=meta["adapt_frequency"],
                              sampling_proposal=make_pbc_gaussian_proposal(z["sim_lattice"]))
    params = H.to_f32(p64, DEV)
    plan = SamplePlan(wf, sampler)
    state = plan.init(data)._replace(stddev=torch.full((1,), float(np.float32(meta["stddev0"])), device=DEV))
    noise = (f32("normals").contiguous(), f32("uniforms").contiguous())
    d1, stats, st1 = sampler.step(BatchLogProb(wf, params, data), data, state, noise, record_accepts=True)
    accepted = stats["accepted"].cpu().numpy().astype(bool)
    ok = np.ones(accepted.shape[1], dtype=bool)
    n_diff = 0
    for s in range(S):
        diff = (accepted[s] != z["accepted"][s]) & ok
        assert (z["margin"][s][diff] < 1e-6).all(), (s, z["margin"][s][diff])   # the north star's exclusion margin
        n_diff += int(diff.sum())
        ok &= ~diff
    assert n_diff == 0, f"{n_diff} accept decisions differ from the reference"
    assert abs(float(stats["pmove"]) - float(z["pmove"][0])) < 1e-6
    got = d1.electrons.cpu().numpy().astype(np.float64)
    if meta["kind"] == "mcmc":
        np.testing.assert_allclose(got, z["electrons_after_step1"], atol=2e-6)
    else:
        inv = np.linalg.inv(z["sim_lattice"])
        d = (got - z["electrons_after_step1"]) @ inv
        assert np.abs((d - np.round(d)) @ z["sim_lattice"]).max() < 2e-5
    # second step from the reference's positions: exercises the width adaptation (adapt_frequency = 2)
    data2 = data.merge({"electrons": f32("electrons_after_step1")})
    _, stats2, st2 = plan.step(params, data2, st1, noise)
    if abs(float(stats2["pmove"]) - float(z["pmove"][1])) < 1e-6:
        assert abs(float(st2.stddev) - float(z["stddev_after"][1])) < 1e-6
        assert st2.counter == int(z["counter_after"][1])
