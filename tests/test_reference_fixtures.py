"""The float64 oracle against golden vectors produced by the REFERENCE's own source.

``tests/golden/ref_*.npz`` are written by ``scripts/make_reference_fixtures.py``: the reference's modules
(/root/reference/src/jaqmc: wavefunction classes, features, envelopes, LogDet, Jastrow, ``potential_energy``,
``EwaldSum``, solid ``PotentialEnergy``, ``MCMCSampler``) executed in float64 -- in the build container over the
torch-backed stand-ins of jax / flax (``scripts/refshim``; ``--backend jax`` regenerates them from the real stack).
Parameters are stored under the reference's own Flax tree paths, so these tests also pin the tree layout the C-ABI
marshalling expects.  Agreement is asked to 1e-9: both sides are float64 evaluations of the same
function, so anything beyond rounding is a transcription error in the oracle."""

import glob
import json
import os

import numpy as np
import pytest
import torch

import helpers as H
from oracle import estimators as OE
from oracle import networks as ON

GOLDEN = os.path.join(H.ROOT, "tests", "golden")
F64 = torch.float64
RTOL = 1e-9


def load(name):
    path = os.path.join(GOLDEN, f"ref_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    flat = {k[len("param:"):]: torch.from_numpy(z[k].astype(np.float64)) for k in z.files if k.startswith("param:")}
    tree = {}
    for path_, v in flat.items():
        node = tree
        keys = path_.split("/")
        for k in keys[:-1]:
            node = node.setdefault(k, {})
        node[keys[-1]] = v
    return meta, tree, z


def oracle_logpsi_fn(meta, params, z):
    """``e -> (sign, logpsi)`` of the oracle for a molecular fixture."""
    kw = meta["kwargs"]
    atoms = torch.from_numpy(z["atoms"].astype(np.float64))
    nspins = tuple(meta["nspins"])
    env = kw.get("envelope", "abs_isotropic")
    if meta["kind"] == "ferminet":
        return lambda e: ON.ferminet_logpsi(params, e, atoms, nspins, env, kw.get("use_last_layer", False))
    if meta["kind"] == "lapnet":
        return lambda e: ON.lapnet_logpsi(params, e, atoms, nspins, kw["num_heads"], env)
    if meta["kind"] == "psiformer":
        return lambda e: ON.psiformer_logpsi(params, e, atoms, nspins, kw.get("layer_norm_mode", "pre"), env)
    raise KeyError(meta["kind"])


def oracle_orbitals(meta, params, z, e):
    kw = meta["kwargs"]
    atoms = torch.from_numpy(z["atoms"].astype(np.float64))
    nspins = tuple(meta["nspins"])
    env = kw.get("envelope", "abs_isotropic")
    if meta["kind"] == "ferminet":
        return ON.ferminet_orbitals(params, e, atoms, nspins, env, kw.get("use_last_layer", False))
    if meta["kind"] == "lapnet":
        return ON.lapnet_orbitals(params, e, atoms, nspins, kw["num_heads"], env)[0]
    return ON.psiformer_orbitals(params, e, atoms, nspins, kw.get("layer_norm_mode", "pre"), env)[0]


MOLECULE = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_*.npz"))
                  if os.path.basename(p)[4:].split("_")[0] in ("ferminet", "lapnet", "psiformer"))


def test_fixtures_present():
    """The committed set: every network family, every envelope type, both LayerNorm options, the sampler, Ewald."""
    need = {"ferminet_li", "ferminet_n2", "ferminet_h_single_channel", "ferminet_lih_isotropic", "ferminet_lih_diagonal",
            "ferminet_lih_null", "ferminet_lih_nosplit", "ferminet_lih_last_layer", "lapnet_n2", "lapnet_lih_layernorm",
            "lapnet_li_nojastrow", "psiformer_n2", "psiformer_lih_post", "psiformer_he_null"}
    assert need <= set(MOLECULE), need - set(MOLECULE)
    for nm in ("solid_cubic_h2", "solid_fcc_lih_221", "ewald", "mcmc_lih", "mcmc_pbc"):
        assert os.path.exists(os.path.join(GOLDEN, f"ref_{nm}.npz")), nm


@pytest.mark.parametrize("name", MOLECULE)
def test_oracle_matches_reference_molecule(name):
    meta, params, z = load(name)
    fn = oracle_logpsi_fn(meta, params, z)
    el = torch.from_numpy(z["electrons"].astype(np.float64))
    atoms = torch.from_numpy(z["atoms"].astype(np.float64))
    charges = torch.from_numpy(z["charges"].astype(np.float64))
    ref = H.oracle_batch(fn, el, atoms, charges)
    assert np.array_equal(ref["sign"], z["sign"])
    np.testing.assert_allclose(ref["logpsi"], z["logpsi"], rtol=RTOL, atol=1e-11)
    gscale = np.abs(z["grad"]).max()
    np.testing.assert_allclose(ref["grad"], z["grad"], rtol=0, atol=1e-9 * gscale)
    lscale = np.abs(z["lap"]) + (z["grad"] ** 2).sum(1)
    assert (np.abs(ref["lap"] - z["lap"]) / lscale).max() < 1e-9      # forward recurrences vs autograd Hessian trace
    assert (np.abs(ref["e_kin"] - z["e_kin"]) / lscale).max() < 1e-9
    np.testing.assert_allclose(ref["e_pot"], z["e_pot"], rtol=1e-12)
    from oracle import lap as L

    orb = oracle_orbitals(meta, params, z, el[0])
    np.testing.assert_allclose(L.value(orb).numpy(), z["orbitals"][0], rtol=1e-9, atol=1e-12)


SOLID = ["solid_cubic_h2", "solid_fcc_lih_221", "solid_cubic_h2_nu", "solid_fcc_lih_221_tri_fcc", "solid_fcc_lih_221_nu_bcc",
         "solid_cubic_h2_tri_hexagonal"]


@pytest.mark.parametrize("name", SOLID)
def test_oracle_matches_reference_solid(name):
    meta, params, z = load(name)
    t = lambda k: torch.from_numpy(z[k].astype(np.float64))  # noqa: E731
    nspins = tuple(meta["nspins"])
    opts = (meta["kwargs"].get("distance_type", "tri"), meta["kwargs"].get("sym_type", "minimal"))

    def logpsi(e):
        return ON.solid_logpsi(params, e, t("prim_atoms"), nspins, t("sim_lattice"), t("prim_lattice"), t("klist"), *opts)

    ew = OE.EwaldSum(z["sim_lattice"])
    assert abs(ew.alpha - float(z["ewald_alpha"])) < 1e-12 * ew.alpha
    assert len(ew.gweight) == int(z["ewald_n_g"])      # same half-space G selection as ewald_gmax=200 + tol 1e-12
    for w in range(z["electrons"].shape[0]):
        e = t("electrons")[w]
        v, g, lap = OE.forward_laplacian(logpsi, e)
        dl = complex(v) - complex(z["logpsi"][w])
        assert abs(dl.real) < 1e-9 * (1 + abs(z["logpsi"][w].real))
        assert abs(np.angle(np.exp(1j * dl.imag))) < 1e-9      # phases agree modulo 2 pi
        gs = np.abs(z["grad"][w]).max()
        assert np.abs(g.numpy() - z["grad"][w]).max() < 1e-8 * gs
        ls = abs(z["lap"][w]) + (np.abs(z["grad"][w]) ** 2).sum()
        assert abs(complex(lap) - complex(z["lap"][w])) < 1e-8 * ls
        ek = complex(-0.5 * lap - 0.5 * (g * g).sum())
        assert abs(ek - complex(z["e_kin"][w])) < 1e-8 * ls
        pot = OE.solid_potential_energy(ew, z["electrons"][w], z["cell_atoms"], z["cell_charges"])
        assert abs(pot - float(z["e_pot"][w])) < 1e-9 * abs(float(z["e_pot"][w]))


def test_oracle_ewald_matches_reference():
    _, _, z = load("ewald")
    e = OE.EwaldSum(z["nacl_lattice"]).energy(z["nacl_coords"], z["nacl_charges"])
    assert abs(e - float(z["nacl_energy"])) < 1e-10 * abs(float(z["nacl_energy"]))
    assert abs(float(z["nacl_madelung"]) + 1.74756) < 1e-4     # reference tests/estimator/ewald_test.py:72-152
    e2 = OE.EwaldSum(z["tri_lattice"]).energy(z["tri_coords"], z["tri_charges"])     # triclinic cell, net charge
    assert abs(e2 - float(z["tri_energy"])) < 1e-10 * abs(float(z["tri_energy"]))


def _mcmc_blp(meta, params, z):
    nspins = tuple(meta["nspins"])
    t = lambda k: torch.from_numpy(z[k].astype(np.float64))  # noqa: E731
    if meta["kind"] == "mcmc":
        atoms = t("atoms")
        one = lambda e: ON.ferminet_logpsi(params, e, atoms, nspins)[1]  # noqa: E731
        wrap = None
    else:
        one = lambda e: ON.solid_logpsi(params, e, t("prim_atoms"), nspins, t("sim_lattice"), t("prim_lattice"),  # noqa: E731
                                        t("klist")).real
        lat = t("sim_lattice")
        inv = torch.linalg.inv(lat)
        wrap = lambda x: ((x @ inv) % 1.0) @ lat  # noqa: E731
    return (lambda x: torch.stack([2.0 * one(x[w]) for w in range(x.shape[0])])), wrap


@pytest.mark.parametrize("name", ["mcmc_lih", "mcmc_pbc"])
def test_oracle_mcmc_matches_reference(name):
    """``MCMCSampler._mh_update`` decisions step by step, then two whole ``step`` calls (the second one adapts the
    proposal width: adapt_frequency = 2)."""
    meta, params, z = load(name)
    blp, wrap = _mcmc_blp(meta, params, z)
    normals = torch.from_numpy(z["normals"])
    uniforms = torch.from_numpy(z["uniforms"])
    x = torch.from_numpy(z["electrons"].astype(np.float64))
    sd = torch.tensor(float(np.float32(meta["stddev0"])), dtype=F64)
    lp = blp(x)
    for s in range(meta["steps"]):
        x, lp, cond, ratio = OE.mh_update(blp, x, lp, normals[s], uniforms[s], sd, wrap)
        assert np.array_equal(cond.numpy(), z["accepted"][s]), s
        np.testing.assert_allclose(lp.numpy(), z["logprob"][s], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(x.numpy(), z["electrons_after_step1"], rtol=0, atol=1e-12)
    assert z["accepted"].any() and not z["accepted"].all()
    state = (sd, torch.zeros(2, dtype=F64), 0)
    x = torch.from_numpy(z["electrons"].astype(np.float64))
    for it in range(2):
        x, pmove, state, _ = OE.mcmc_step(blp, x, normals, uniforms, state, steps=meta["steps"],
                                          adapt_frequency=meta["adapt_frequency"], wrap=wrap)
        assert abs(pmove - float(z["pmove"][it])) < 1e-12
        assert abs(float(state[0]) - float(z["stddev_after"][it])) < 1e-12
        np.testing.assert_allclose(state[1].numpy(), z["pmoves_after"][it], atol=1e-12)
        assert state[2] == int(z["counter_after"][it])
        np.testing.assert_allclose(x.numpy(), z[f"electrons_after_step{it + 1}"], rtol=0, atol=1e-12)
    assert float(z["stddev_after"][1]) != float(z["stddev_after"][0])   # the adaptation branch was exercised


@pytest.mark.parametrize("name", ["observables_lih", "observables_li"])
def test_oracle_observables_match_reference(name):
    """psi-ratio consumers: ``SpinSquared`` (estimator/spin.py) and the non-local ECP integral
    (estimator/ecp/nonlocal_integral.py) of the reference against the oracle's restatement."""
    from jaqmc_b200.ecp import Quadrature, get_quadrature
    from oracle import observables as OO

    meta, params, z = load(name)
    nspins = tuple(meta["nspins"])
    atoms = torch.from_numpy(z["atoms"].astype(np.float64))
    phase = lambda e: ON.ferminet_logpsi(params, e, atoms, nspins)  # noqa: E731
    quad = get_quadrature(meta["quadrature"])
    np.testing.assert_allclose(quad.pts.numpy(), z["quad_pts"], atol=1e-14)      # the host mirror's tables
    np.testing.assert_allclose(quad.coefs.numpy(), z["quad_coefs"], atol=1e-15)
    n = z["electrons"].shape[1]
    for w in range(z["electrons"].shape[0]):
        e = torch.from_numpy(z["electrons"][w].astype(np.float64))
        s2 = OO.spin_squared(phase, e, *nspins)
        assert abs(float(s2) - float(z["s2"][w])) < 1e-9 * (1 + abs(float(z["s2"][w])))
        rot = Quadrature.rotation_matrices(2 * np.pi * torch.from_numpy(z["u1"][w]), 1.0 - 2.0 * torch.from_numpy(z["u2"][w]))
        pts = torch.einsum("ijk,lk->ilj", rot, quad.pts)
        atom_pos = atoms[None].expand(n, -1, -1)
        got = OO.nonlocal_integral(phase, e, atom_pos, pts, quad.coefs, meta["num_channels"] - 1)
        np.testing.assert_allclose(got.numpy(), z["nonlocal_integrals"][w], rtol=1e-8, atol=1e-10)
