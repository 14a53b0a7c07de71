"""GPU tests of the estimator surface (reference estimator/base.py:251-362, estimator/kinetic/euclidean.py:114-135,
estimator/total_energy.py:36-60, app/molecule/hamiltonian.py:9-22, app/solid/hamiltonian.py:18-56,
app/hydrogen_atom.py:28-39) and of the envelope options (wavefunction/output/envelope.py:18-140)."""

import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import helpers as H
from jaqmc_b200 import _marshal as M
from oracle import estimators as OE
from oracle import networks as ON

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _rt():
    from jaqmc_b200._runtime import runtime

    return runtime(DEV)


def test_hydrogen_atom_local_energy_on_cuda():
    """Config C1: ``log psi = alpha |r|`` -> E_L = -alpha^2/2 - (alpha + 1)/|r|; exactly -0.5 Ha at alpha = -1
    (reference tests/hydrogen/atom_test.py:47-52 converges to it)."""
    from jaqmc_b200.data import MoleculeData
    from jaqmc_b200.estimator import EstimatorPipeline, EuclideanKinetic, TotalEnergy, potential_energy
    from jaqmc_b200.wavefunction import HydrogenAtom

    g = torch.Generator().manual_seed(0)
    el = torch.randn(4096, 1, 3, generator=g).to(DEV)
    data = MoleculeData(el, torch.zeros(1, 3, device=DEV), torch.ones(1, device=DEV))
    wf = HydrogenAtom()
    params = wf.init_params(data)
    assert float(params["params"]["alpha"]) == pytest.approx(-0.8)
    out = wf.local_energy(params, data)
    ref = OE.hydrogen_local_energy(-0.8, el.cpu().double())
    np.testing.assert_allclose(out["e_loc"].cpu().numpy(), ref.numpy(), rtol=3e-6, atol=3e-6)
    params = {"params": {"alpha": torch.tensor([-1.0], device=DEV)}}
    pipe = EstimatorPipeline({"kinetic": EuclideanKinetic(f_log_psi=wf), "potential": potential_energy,
                              "total": TotalEnergy()})
    stats, walkers = pipe.evaluate(params, data)
    assert abs(float(stats["total_energy"]) + 0.5) < 1e-5 and float(stats["total_energy_var"]) < 1e-8
    np.testing.assert_allclose(walkers["total_energy"].cpu().numpy(), -0.5, atol=2e-5)
    lp = wf.logpsi(params, data)
    np.testing.assert_allclose(lp.cpu().numpy(), -el.norm(dim=-1).squeeze(-1).cpu().numpy(), rtol=1e-6)


def test_estimator_pipeline_end_to_end_matches_oracle():
    """``EstimatorPipeline({kinetic, potential, total})`` -- the estimator half of
    ``EvaluationWorkStage.compute_step`` (workflow/stage/evaluation.py:190-192) -- per-walker keys and the
    ``mean_reduce`` statistics against the oracle."""
    from jaqmc_b200.data import MoleculeData
    from jaqmc_b200.estimator import EstimatorPipeline, EuclideanKinetic, TotalEnergy, potential_energy
    from jaqmc_b200.wavefunction import FermiNetWavefunction

    atoms, charges, nspins = H.molecule("LiH")
    hs, hd, ndets = [64, 64, 64], [16, 16, 16], 4
    wf = FermiNetWavefunction(nspins=nspins, ndets=ndets, hidden_dims_single=hs, hidden_dims_double=hd)
    p64 = H.round_f32(ON.init_ferminet_params(nspins, atoms.shape[0], ndets, hs, hd, seed=3))
    W = 16
    el = H.synthetic_walkers(atoms, charges, nspins, W, seed=5)
    data = MoleculeData(el.float().to(DEV), atoms.float().to(DEV), charges.float().to(DEV))
    params = H.to_f32(p64, DEV)
    pipe = EstimatorPipeline({"kinetic": EuclideanKinetic(f_log_psi=wf), "potential": potential_energy,
                              "total": TotalEnergy()})
    stats, walkers = pipe.evaluate(params, data)
    ref = H.oracle_batch(lambda e: ON.ferminet_logpsi(p64, e, atoms, nspins), el, atoms, charges)
    assert set(walkers) == {"energy:kinetic", "energy:potential", "total_energy"}
    scale = 0.5 * np.abs(ref["lap"]) + 0.5 * (ref["grad"] ** 2).sum(1) + np.abs(ref["e_pot"])
    for key, r in (("energy:kinetic", ref["e_kin"]), ("energy:potential", ref["e_pot"]),
                   ("total_energy", ref["e_kin"] + ref["e_pot"])):
        err = np.abs(walkers[key].cpu().numpy() - r) / scale
        assert err.max() < 1e-4 and np.median(err) < 1e-5, (key, err)
    tot = torch.from_numpy(ref["e_kin"] + ref["e_pot"])
    want = OE.mean_reduce({"total_energy": tot})
    s = float(scale.mean())
    assert abs(float(stats["total_energy"]) - float(want["total_energy"])) < 1e-5 * s
    assert abs(float(stats["total_energy_var"]) - float(want["total_energy_var"])) < 1e-3 * max(1.0, float(want["total_energy_var"]))
    # TotalEnergy's error behaviour (estimator/total_energy.py:52-58)
    with pytest.raises(ValueError):
        TotalEnergy().evaluate_batch_walkers(params, data, {})
    with pytest.raises(ValueError):
        TotalEnergy().evaluate_batch_walkers(params, data, {"energy:x": torch.zeros(W, 2, device=DEV)})


def test_solid_potential_energy_estimator():
    """``SolidPotentialEnergy`` (app/solid/hamiltonian.py:18-56) through the estimator interface."""
    from jaqmc_b200.data import SolidData
    from jaqmc_b200.ewald import SolidPotentialEnergy

    prim, sim, patoms, cell_atoms, cell_charges, nspins, klist = H.solid_system("fcc_lih_221")
    f32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=DEV)  # noqa: E731
    W = 8
    el = H.solid_walkers(cell_atoms, sum(nspins), W, seed=2)
    data = SolidData(torch.from_numpy(el).to(DEV), f32(cell_atoms), f32(cell_charges), f32(patoms))
    est = SolidPotentialEnergy(sim, device=DEV)
    stats, _ = est.evaluate_batch_walkers(None, data)
    ref_ew = OE.EwaldSum(sim)
    ref = np.array([OE.solid_potential_energy(ref_ew, el[w], cell_atoms, cell_charges) for w in range(W)])
    np.testing.assert_allclose(stats["energy:potential"].cpu().numpy(), ref, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("envelope", ["isotropic", "null", "abs_isotropic"])
@pytest.mark.parametrize("mol,hs,hd", [("LiH", (64, 64), (16, 16)), ("Li", (16, 16), (8, 8))])
def test_envelope_types(envelope, mol, hs, hd):
    """``EnvelopeType.isotropic`` (exponent sigma * -r, no abs) and ``null`` (factor of ones), on the fused tcgen05
    epilogue (64-wide network) and on the separate envelope pass (16-wide network)."""
    rt = _rt()
    atoms, charges, nspins = H.molecule(mol)
    ndets = 4
    p64 = H.round_f32(ON.init_ferminet_params(nspins, atoms.shape[0], ndets, hs, hd, seed=9))
    if envelope == "isotropic":   # make some decay rates negative: isotropic and abs_isotropic then differ
        for nm in p64["params"]["envelope_layer"]:
            sg = p64["params"]["envelope_layer"][nm]["sigma"]
            sg[::2] = -0.3 * sg[::2]
    if envelope == "null":
        p64["params"]["envelope_layer"] = {}
    W = 8
    el = H.synthetic_walkers(atoms, charges, nspins, W, seed=1)
    wf = M.ferminet_handle(H.to_f32(p64, DEV), nspins, atoms.shape[0], ndets, hs, hd, envelope)
    sysh = M.system_handle(atoms.float().to(DEV), charges.float().to(DEV))
    e32 = el.float().contiguous().to(DEV)
    out = {k: v.cpu().numpy() for k, v in rt.local_energy(wf, sysh, e32).items()}
    ref = H.oracle_batch(lambda e: ON.ferminet_logpsi(p64, e, atoms, nspins, envelope), el, atoms, charges)
    assert np.array_equal(out["sign"], ref["sign"])
    H.assert_fp32_parity(out, ref, el)
    lp, sg = rt.logpsi(wf, sysh, e32)
    _, l_scale = H.fp32_scales(ref, el)
    assert (np.abs(lp.cpu().numpy() - ref["logpsi"]) / l_scale).max() < 1e-5
    if envelope == "isotropic":
        other = H.oracle_batch(lambda e: ON.ferminet_logpsi(p64, e, atoms, nspins, "abs_isotropic"), el[:2], atoms,
                               charges, track=False)
        assert np.abs(other["logpsi"] - ref["logpsi"][:2]).max() > 1e-3   # the test distinguishes the two types


_SPLIT_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import helpers as H
from jaqmc_b200 import _marshal as M
from jaqmc_b200._runtime import runtime
from oracle import networks as ON
dev = torch.device("cuda", 0)
atoms, charges, nspins = H.molecule("Li")          # n_up = 2, n_dn = 1
hs, hd, ndets = (256,) * 4, (32,) * 4, 16
p64 = H.round_f32(ON.init_ferminet_params(nspins, 1, ndets, hs, hd, seed=2))
el = H.synthetic_walkers(atoms, charges, nspins, 8, seed=0)
wf = M.ferminet_handle(H.to_f32(p64, dev), nspins, 1, ndets, hs, hd)
sysh = M.system_handle(atoms.float().to(dev), charges.float().to(dev))
out = {k: v.cpu().numpy() for k, v in runtime(dev).local_energy(wf, sysh, el.float().contiguous().to(dev)).items()}
ref = H.oracle_batch(lambda e: ON.ferminet_logpsi(p64, e, atoms, nspins), el, atoms, charges)
assert np.array_equal(out["sign"], ref["sign"])
H.assert_fp32_parity(out, ref, el)
print("SPLIT-OK")
"""


def test_envelope_not_applied_twice_when_only_one_channel_takes_the_tensor_core_path():
    """with ``JAQMC_B200_TC_MIN_WORK`` between the work of the two orbital launches of Li (2 up,
    1 down electrons: 8 walkers -> 2.2 M / 1.1 M multiply-adds) only the spin-up launch is tcgen05-eligible.  The
    library reads the switch once per process, hence the subprocess."""
    env = dict(os.environ, JAQMC_B200_TC_MIN_WORK="1.5e6")
    r = subprocess.run([sys.executable, "-c", _SPLIT_SCRIPT, H.ROOT], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SPLIT-OK" in r.stdout, r.stdout + r.stderr
