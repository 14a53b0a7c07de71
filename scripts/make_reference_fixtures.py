#!/usr/bin/env python
"""Golden vectors produced by the REFERENCE's own source (bytedance/jaqmc at /root/reference), for pinning the oracle.

    python scripts/make_reference_fixtures.py [--backend shim|jax] [--out tests/golden] [--only NAME ...]

``--backend jax``  : the unmodified stack (needs jax / flax / pyserde and the reference's other dependencies installed;
                     x64 on CPU; gradients / Laplacians by the reference's ``forward_laplacian``).  Not runnable in the
                     build container (no jax).
``--backend shim`` : the reference's modules imported from /root/reference/src over the torch-float64 stand-ins of
                     ``scripts/refshim`` (see its README for what that does and does not establish); gradients /
                     Laplacians by ``torch.autograd`` of the reference's ``logpsi``.  This is how the committed
                     ``tests/golden/ref_*.npz`` were made.

Every fixture holds: ``meta`` (JSON: kind, constructor arguments, system), the parameter tree flattened to
``param:<path>`` arrays (values exactly float32-representable, so the float32 kernels see the same numbers), the
walkers, and the reference's outputs in float64.  ``tests/test_reference_fixtures.py`` checks the oracle against them
(CPU), ``tests/test_gpu_reference_fixtures.py`` the CUDA kernels.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--backend", default="shim", choices=["shim", "jax"])
ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
ap.add_argument("--only", nargs="*", default=None)
args = ap.parse_args()

if args.backend == "shim":
    sys.path.insert(0, os.path.join(ROOT, "scripts", "refshim"))
    import bootstrap

    bootstrap.install()
    import torch

    ref = bootstrap.ref

    def to_backend(a):
        return torch.as_tensor(np.asarray(a, dtype=np.float64))

    def to_numpy(t):
        # a copy: the reference's ``counter += 1`` rebinds an immutable jax array but mutates a torch tensor in place
        return t.detach().cpu().numpy().copy() if isinstance(t, torch.Tensor) else np.array(t)

    def value_grad_lap(f, x):
        """(f(x), grad, laplacian) of a scalar (possibly complex) function of the electron array."""
        x = x.clone().requires_grad_(True)
        y = f(x)
        parts = [y.real, y.imag] if y.is_complex() else [y]
        gs, laps = [], []
        for part in parts:
            (g,) = torch.autograd.grad(part, x, create_graph=True)
            gf = g.reshape(-1)
            lap = 0.0
            for i in range(gf.numel()):
                (gg,) = torch.autograd.grad(gf[i], x, retain_graph=True)
                lap = lap + gg.reshape(-1)[i]
            gs.append(gf.detach())
            laps.append(lap.detach())
        if len(parts) == 2:
            return y.detach(), torch.complex(gs[0], gs[1]), torch.complex(laps[0], laps[1])
        return y.detach(), gs[0], laps[0]
else:   # the real stack
    import importlib

    import jax

    jax.config.update("jax_enable_x64", True)
    jax.config.update("jax_platforms", "cpu")
    import jax.numpy as jnp

    def ref(name):
        return importlib.import_module(name)

    def to_backend(a):
        return jnp.asarray(np.asarray(a, dtype=np.float64))

    def to_numpy(t):
        return np.asarray(t)

    def value_grad_lap(f, x):
        from jaqmc.laplacian import forward_laplacian   # the reference's own interpreter

        out = forward_laplacian(f)(x)
        return out.x, out.dense_jacobian.reshape(-1), out.laplacian

from jaqmc_b200 import systems  # noqa: E402  (atoms / charges / synthetic walkers of the named systems)


# ---------------------------------------------------------------------------------------------------------------
def flatten(tree, prefix=""):
    out = {}
    for k in sorted(tree):
        v = tree[k]
        if isinstance(v, dict):
            out.update(flatten(v, f"{prefix}{k}/"))
        else:
            out[f"{prefix}{k}"] = v
    return out


def unflatten(flat):
    tree = {}
    for path, v in flat.items():
        node = tree
        keys = path.split("/")
        for k in keys[:-1]:
            node = node.setdefault(k, {})
        node[keys[-1]] = v
    return tree


def jittered_params(wf, data, seed):
    """The reference's ``init_params`` tree, every leaf moved off its initial value (zero biases, unit envelopes) so
    that each term matters, rounded to float32-representable values."""
    import jax as J

    params = wf.init_params(data, J.random.PRNGKey(seed))
    g = np.random.default_rng(seed)
    flat = flatten(params)
    out = {}
    for path, v in flat.items():
        a = to_numpy(v).astype(np.float64)
        a = a + 0.15 * g.standard_normal(a.shape)
        out[path] = a.astype(np.float32).astype(np.float64)
    return out


def tree_to_backend(flat):
    return unflatten({k: to_backend(v) for k, v in flat.items()})


def save(name, meta, flat_params, arrays):
    os.makedirs(args.out, exist_ok=True)
    payload = {"meta": np.array(json.dumps(meta))}
    for k, v in flat_params.items():
        payload["param:" + k] = np.asarray(v, dtype=np.float32)
    for k, v in arrays.items():
        payload[k] = np.asarray(v)
    path = os.path.join(args.out, f"ref_{name}.npz")
    np.savez_compressed(path, **payload)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


# ---------------------------------------------------------------------------------------------------------------
MOLECULE_CASES = {
    # name: (kind, molecule, constructor kwargs, walkers)
    "ferminet_li": ("ferminet", "Li", dict(ndets=4, hidden_dims_single=[32, 32, 32], hidden_dims_double=[8, 8, 8]), 4),
    "ferminet_n2": ("ferminet", "N2", dict(ndets=4, hidden_dims_single=[32, 32], hidden_dims_double=[8, 8]), 3),
    "ferminet_h_single_channel": ("ferminet", "H", dict(ndets=2, hidden_dims_single=[8, 8], hidden_dims_double=[4, 4]), 3),
    "ferminet_lih_isotropic": ("ferminet", "LiH", dict(ndets=3, hidden_dims_single=[16, 16], hidden_dims_double=[8, 8],
                                                        envelope="isotropic"), 3),
    "ferminet_lih_diagonal": ("ferminet", "LiH", dict(ndets=3, hidden_dims_single=[16, 16], hidden_dims_double=[8, 8],
                                                       envelope="diagonal"), 3),
    "ferminet_lih_null": ("ferminet", "LiH", dict(ndets=3, hidden_dims_single=[16, 16], hidden_dims_double=[8, 8],
                                                   envelope="null"), 3),
    "ferminet_lih_nosplit": ("ferminet", "LiH", dict(ndets=3, hidden_dims_single=[16, 16], hidden_dims_double=[8, 8],
                                                      orbitals_spin_split=False), 3),
    "ferminet_lih_last_layer": ("ferminet", "LiH", dict(ndets=3, hidden_dims_single=[16, 16], hidden_dims_double=[8, 8],
                                                         use_last_layer=True), 3),
    "lapnet_n2": ("lapnet", "N2", dict(ndets=4, num_layers=2, num_heads=2, heads_dim=16), 3),
    "lapnet_lih_layernorm": ("lapnet", "LiH", dict(ndets=3, num_layers=2, num_heads=2, heads_dim=8, use_layernorm=True), 3),
    "lapnet_li_nojastrow": ("lapnet", "Li", dict(ndets=2, num_layers=3, num_heads=2, heads_dim=8, jastrow="none"), 3),
    "psiformer_n2": ("psiformer", "N2", dict(ndets=4, num_layers=2, num_heads=2, heads_dim=16, mlp_hidden_dims=[32]), 3),
    "psiformer_lih_post": ("psiformer", "LiH", dict(ndets=3, num_layers=2, num_heads=2, heads_dim=8, mlp_hidden_dims=[16],
                                                    layer_norm_mode="post"), 3),
    "psiformer_he_null": ("psiformer", "He", dict(ndets=2, num_layers=2, num_heads=2, heads_dim=8, mlp_hidden_dims=[16, 24],
                                                  layer_norm_mode="null", bias_orbitals=True), 3),
}

WF_CLASS = {
    "ferminet": ("jaqmc.app.molecule.wavefunction.ferminet", "FermiNetWavefunction"),
    "lapnet": ("jaqmc.app.molecule.wavefunction.lapnet", "LapNetWavefunction"),
    "psiformer": ("jaqmc.app.molecule.wavefunction.psiformer", "PsiformerWavefunction"),
}


def molecule_case(name, kind, mol, kw, W, seed=0):
    atoms, charges, nspins = systems.molecule(mol)
    el = systems.synthetic_walkers(atoms, charges, nspins, W, seed=seed + 11).numpy()
    mod, cls = WF_CLASS[kind]
    wf = getattr(ref(mod), cls)(nspins=tuple(nspins), **kw)
    MoleculeData = ref("jaqmc.app.molecule.data").MoleculeData
    potential_energy = ref("jaqmc.app.molecule.hamiltonian").potential_energy
    kin = ref("jaqmc.estimator.kinetic._common")._apply_kinetic_formula
    A, Z = to_backend(atoms.numpy()), to_backend(charges.numpy())
    data0 = MoleculeData(electrons=to_backend(el[0]), atoms=A, charges=Z)
    flat = jittered_params(wf, data0, seed + 3)
    params = tree_to_backend(flat)
    out = dict(logpsi=[], sign=[], grad=[], lap=[], e_kin=[], e_pot=[], orbitals=[])
    for w in range(W):
        x = to_backend(el[w])
        data = MoleculeData(electrons=x, atoms=A, charges=Z)
        sign, lp = wf.phase_logpsi(params, data)
        v, g, lap = value_grad_lap(lambda e: wf.logpsi(params, MoleculeData(electrons=e, atoms=A, charges=Z)), x)
        assert abs(float(to_numpy(v)) - float(to_numpy(lp))) < 1e-12
        out["logpsi"].append(to_numpy(lp))
        out["sign"].append(to_numpy(sign))
        out["grad"].append(to_numpy(g))
        out["lap"].append(to_numpy(lap))
        out["e_kin"].append(to_numpy(kin(lap, (g * g).sum())))
        out["e_pot"].append(to_numpy(potential_energy(None, data, None, None, None)[0]["energy:potential"]))
        out["orbitals"].append(to_numpy(wf.orbitals(params, data)))
    meta = dict(kind=kind, molecule=mol, nspins=list(nspins), kwargs=kw, backend=args.backend,
                reference="bytedance/jaqmc 0.1.0 app/molecule/wavefunction/%s.py" % kind)
    arrays = {k: np.stack(v) for k, v in out.items()}
    arrays.update(electrons=el, atoms=atoms.numpy(), charges=charges.numpy())
    save(name, meta, flat, arrays)


# ---------------------------------------------------------------------------------------------------------------
def solid_case(name, kind, kw, W, seed=0):
    prim, sim, patoms, cell_atoms, cell_charges, nspins, klist = systems.solid_system(kind)
    n = sum(nspins)
    el = systems.solid_walkers(cell_atoms, n, W, seed=seed + 5).astype(np.float64)
    SolidWavefunction = ref("jaqmc.app.solid.wavefunction").SolidWavefunction
    SolidData = ref("jaqmc.app.solid.data").SolidData
    wf = SolidWavefunction(nspins=tuple(nspins), simulation_lattice=to_backend(sim), primitive_lattice=to_backend(prim),
                           klist=to_backend(klist), **kw)
    mk = lambda e: SolidData(electrons=e, atoms=to_backend(cell_atoms), charges=to_backend(cell_charges),  # noqa: E731
                             primitive_atoms=to_backend(patoms))
    flat = jittered_params(wf, mk(to_backend(el[0])), seed + 9)
    params = tree_to_backend(flat)
    PotentialEnergy = ref("jaqmc.app.solid.hamiltonian").PotentialEnergy
    pe = PotentialEnergy(supercell_lattice=to_backend(sim))
    pe.init(mk(to_backend(el[0])), None)
    kin = ref("jaqmc.estimator.kinetic._common")._apply_kinetic_formula
    out = dict(logpsi=[], grad=[], lap=[], e_kin=[], e_pot=[])
    for w in range(W):
        x = to_backend(el[w])
        v, g, lap = value_grad_lap(lambda e: wf.logpsi(params, mk(e)), x)
        out["logpsi"].append(to_numpy(v))
        out["grad"].append(to_numpy(g))
        out["lap"].append(to_numpy(lap))
        out["e_kin"].append(to_numpy(kin(lap, (g * g).sum())))
        out["e_pot"].append(to_numpy(pe.evaluate_single_walker(None, mk(x), None, None, None)[0]["energy:potential"]))
    ew = pe.ewald
    meta = dict(kind="solid", system=kind, nspins=list(nspins), kwargs=kw, backend=args.backend,
                reference="bytedance/jaqmc 0.1.0 app/solid/wavefunction.py, app/solid/hamiltonian.py, estimator/ewald.py")
    arrays = {k: np.stack(v) for k, v in out.items()}
    arrays.update(electrons=el, prim_lattice=prim, sim_lattice=sim, prim_atoms=patoms, cell_atoms=cell_atoms,
                  cell_charges=cell_charges, klist=klist, ewald_alpha=to_numpy(ew.alpha),
                  ewald_n_g=np.array(int(ew.gpoints.shape[0])), ewald_gpoints=to_numpy(ew.gpoints),
                  ewald_gweight=to_numpy(ew.gweight))
    save(name, meta, flat, arrays)


def madelung_case():
    """EwaldSum on point-charge lattices (reference tests/estimator/ewald_test.py:72-152: NaCl -1.74756, CaF2 -5.03879)."""
    EwaldSum = ref("jaqmc.estimator.ewald").EwaldSum
    a = 1.0
    lat = a * np.array([[0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]])
    coords = np.array([[0.0, 0.0, 0.0], [0.5 * a, 0.5 * a, 0.5 * a]])
    charges = np.array([1.0, -1.0])
    e = EwaldSum(to_backend(lat)).energy(to_backend(coords), to_backend(charges))
    madelung = float(to_numpy(e)) * (a / 2)     # per ion pair, in units of nearest-neighbour distance
    g = np.random.default_rng(3)
    lat2 = np.array([[5.1, 0.2, 0.0], [0.3, 4.7, 0.4], [0.1, -0.2, 6.0]])
    c2 = g.uniform(-1, 7, (9, 3))
    q2 = g.uniform(-2, 2, 9)
    e2 = EwaldSum(to_backend(lat2)).energy(to_backend(c2), to_backend(q2))
    save("ewald", dict(kind="ewald", backend=args.backend, reference="bytedance/jaqmc 0.1.0 estimator/ewald.py"), {},
         dict(nacl_lattice=lat, nacl_coords=coords, nacl_charges=charges, nacl_energy=to_numpy(e),
              nacl_madelung=np.array(madelung), tri_lattice=lat2, tri_coords=c2, tri_charges=q2, tri_energy=to_numpy(e2)))
    print("NaCl Madelung constant from the reference's EwaldSum:", madelung)


# ---------------------------------------------------------------------------------------------------------------
def mcmc_case(name="mcmc_lih", pbc=False):
    """``MCMCSampler._mh_update`` / ``step`` (sampler/mcmc.py:96-197) on queued noise: per-step accept decisions,
    final walkers, pmove and the adapted state -- with the Gaussian proposal, and with the PBC proposal of
    geometry/pbc.py:187-201 on a periodic wavefunction."""
    if args.backend != "shim":
        raise SystemExit("the MCMC fixture replays queued noise through the stand-in jax.random (shim backend only); "
                         "with real JAX use jax.random and store the draws instead")
    import jax as J
    from jax import random as R

    MCMCSampler = ref("jaqmc.sampler.mcmc").MCMCSampler
    g = np.random.default_rng(21)
    S, W = 5, 24
    if not pbc:
        atoms, charges, nspins = systems.molecule("LiH")
        el = systems.synthetic_walkers(atoms, charges, nspins, W, seed=4).numpy()
        kw = dict(ndets=3, hidden_dims_single=[16, 16], hidden_dims_double=[8, 8])
        wf = ref("jaqmc.app.molecule.wavefunction.ferminet").FermiNetWavefunction(nspins=tuple(nspins), **kw)
        MoleculeData = ref("jaqmc.app.molecule.data").MoleculeData
        A, Z = to_backend(atoms.numpy()), to_backend(charges.numpy())
        mk = lambda e: MoleculeData(electrons=e, atoms=A, charges=Z)  # noqa: E731
        sampler = MCMCSampler(steps=S, adapt_frequency=2)
        meta = dict(kind="mcmc", wf="ferminet", molecule="LiH", nspins=list(nspins), kwargs=kw)
        extra = dict(atoms=atoms.numpy(), charges=charges.numpy())
    else:
        prim, sim, patoms, cell_atoms, cell_charges, nspins, klist = systems.solid_system("fcc_lih_221")
        el = systems.solid_walkers(cell_atoms, sum(nspins), W, seed=4).astype(np.float64)
        kw = dict(ndets=2, hidden_dims_single=[16, 16], hidden_dims_double=[8, 8])
        wf = ref("jaqmc.app.solid.wavefunction").SolidWavefunction(
            nspins=tuple(nspins), simulation_lattice=to_backend(sim), primitive_lattice=to_backend(prim),
            klist=to_backend(klist), **kw)
        SolidData = ref("jaqmc.app.solid.data").SolidData
        mk = lambda e: SolidData(electrons=e, atoms=to_backend(cell_atoms), charges=to_backend(cell_charges),  # noqa: E731
                                 primitive_atoms=to_backend(patoms))
        proposal = ref("jaqmc.geometry.pbc").make_pbc_gaussian_proposal(to_backend(sim))
        sampler = MCMCSampler(steps=S, adapt_frequency=2, sampling_proposal=proposal)
        meta = dict(kind="mcmc_pbc", wf="solid", system="fcc_lih_221", nspins=list(nspins), kwargs=kw)
        extra = dict(prim_lattice=prim, sim_lattice=sim, prim_atoms=patoms, cell_atoms=cell_atoms,
                     cell_charges=cell_charges, klist=klist)
    flat = jittered_params(wf, mk(to_backend(el[0])), 13)
    params = tree_to_backend(flat)
    n = el.shape[1]
    normals = g.standard_normal((S, W, n, 3)).astype(np.float32).astype(np.float64)
    uniforms = g.uniform(1e-6, 1.0, (S, W)).astype(np.float32).astype(np.float64)

    def batch_log_prob(x):   # sampler/base.py:173-178: 2 * vmap(logpsi)
        return 2 * J.vmap(lambda e: wf.logpsi(params, mk(e)))(x)

    stddev0 = 0.15 if not pbc else 0.5
    state = sampler.init(None, None)
    state = type(state)(stddev=to_backend(np.float32(stddev0)), pmoves=state.pmoves, counter=state.counter)
    # (a) step by step through the reference's _mh_update: decisions and log-probabilities
    x = to_backend(el)
    lp = batch_log_prob(x).real
    acc, lps, margins = [], [], []
    num = to_backend(0.0)
    for s in range(S):
        # distance of every decision from its threshold, from the reference's own proposal and log-probability
        R.queue_normal(to_backend(normals[s]))
        lp2 = batch_log_prob(sampler.sampling_proposal(None, x, state.stddev)).real
        margins.append(to_numpy(abs((lp2 - lp) - to_backend(np.log(uniforms[s])))))
        R.queue_normal(to_backend(normals[s]))
        R.queue_uniform(to_backend(uniforms[s]))
        x_new, _, lp_new, num = sampler._mh_update(batch_log_prob, x, None, lp, num, stddev=state.stddev)
        acc.append(to_numpy((lp_new != lp) | (x_new != x).reshape(W, -1).any(-1)))
        lps.append(to_numpy(lp_new))
        x, lp = x_new, lp_new
    # (b) the whole step (fori_loop + width adaptation), twice so that the adaptation branch (adapt_frequency=2) runs
    states, pmoves, finals = [], [], []
    xs = to_backend(el)
    for it in range(2):
        for s in range(S):
            R.queue_normal(to_backend(normals[s]))
            R.queue_uniform(to_backend(uniforms[s]))
        xs, stats, state = sampler.step(batch_log_prob, xs, state, None)
        pmoves.append(to_numpy(stats["pmove"]))
        states.append([to_numpy(state.stddev), to_numpy(state.pmoves), to_numpy(state.counter)])
        finals.append(to_numpy(xs))
    assert np.array_equal(finals[0], to_numpy(x))
    meta.update(backend=args.backend, steps=S, adapt_frequency=2, stddev0=stddev0,
                reference="bytedance/jaqmc 0.1.0 sampler/mcmc.py" + (", geometry/pbc.py" if pbc else ""))
    arrays = dict(electrons=el, normals=normals, uniforms=uniforms, accepted=np.stack(acc), logprob=np.stack(lps),
                  margin=np.stack(margins),
                  electrons_after_step1=finals[0], electrons_after_step2=finals[1], pmove=np.array(pmoves),
                  stddev_after=np.array([s[0] for s in states]), pmoves_after=np.stack([s[1] for s in states]),
                  counter_after=np.array([s[2] for s in states]), **extra)
    save(name, meta, flat, arrays)


# ---------------------------------------------------------------------------------------------------------------
def observables_case(name, mol, kw, W):
    """psi-ratio consumers on a FermiNet: ``SpinSquared`` (estimator/spin.py) and the non-local ECP integral
    (estimator/ecp/nonlocal_integral.py, icosahedron_12 quadrature, 2 non-local channels) with queued rotation draws."""
    if args.backend != "shim":
        raise SystemExit("observables fixtures replay queued uniforms through the stand-in jax.random (shim backend)")
    from jax import random as R

    atoms, charges, nspins = systems.molecule(mol)
    el = systems.synthetic_walkers(atoms, charges, nspins, W, seed=23).numpy()
    wf = ref("jaqmc.app.molecule.wavefunction.ferminet").FermiNetWavefunction(nspins=tuple(nspins), **kw)
    MoleculeData = ref("jaqmc.app.molecule.data").MoleculeData
    A, Z = to_backend(atoms.numpy()), to_backend(charges.numpy())
    mk = lambda e: MoleculeData(electrons=e, atoms=A, charges=Z)  # noqa: E731
    flat = jittered_params(wf, mk(to_backend(el[0])), 17)
    params = tree_to_backend(flat)
    SpinSquared = ref("jaqmc.estimator.spin").SpinSquared
    spin = SpinSquared(n_up=nspins[0], n_down=nspins[1], phase_logpsi=wf.phase_logpsi, data_field="electrons")
    spin.init(mk(to_backend(el[0])), None)
    quad = ref("jaqmc.estimator.ecp.quadrature").get_quadrature("icosahedron_12")
    evaluate = ref("jaqmc.estimator.ecp.nonlocal_integral").make_nonlocal_integral(3, quad)
    n = el.shape[1]
    g = np.random.default_rng(31)
    u1 = g.uniform(size=(W, n)).astype(np.float32).astype(np.float64)
    u2 = g.uniform(size=(W, n)).astype(np.float32).astype(np.float64)
    s2, integrals = [], []
    for w in range(W):
        x = to_backend(el[w])
        s2.append(to_numpy(spin.evaluate_single_walker(params, mk(x), None, None, None)[0]["spin:s2"]))
        R.queue_uniform(to_backend(u1[w]))
        R.queue_uniform(to_backend(u2[w]))
        atom_pos = to_backend(np.broadcast_to(atoms.numpy(), (n, atoms.shape[0], 3)).copy())
        integrals.append(to_numpy(evaluate(lambda e: wf.phase_logpsi(params, mk(e)), x, atom_pos, None)))
    meta = dict(kind="observables", wf="ferminet", molecule=mol, nspins=list(nspins), kwargs=kw, backend=args.backend,
                quadrature="icosahedron_12", num_channels=3,
                reference="bytedance/jaqmc 0.1.0 estimator/spin.py, estimator/ecp/nonlocal_integral.py, estimator/ecp/quadrature.py")
    save(name, meta, flat, dict(electrons=el, atoms=atoms.numpy(), charges=charges.numpy(), s2=np.stack(s2),
                                nonlocal_integrals=np.stack(integrals), u1=u1, u2=u2, quad_pts=to_numpy(quad.pts),
                                quad_coefs=to_numpy(quad.coefs)))


# ---------------------------------------------------------------------------------------------------------------
def main():
    want = lambda nm: args.only is None or nm in args.only  # noqa: E731
    for name, (kind, mol, kw, W) in MOLECULE_CASES.items():
        if want(name):
            molecule_case(name, kind, mol, kw, W)
    if want("solid_cubic_h2"):
        solid_case("solid_cubic_h2", "cubic_h2", dict(ndets=2, hidden_dims_single=[16, 16], hidden_dims_double=[8, 8]), 2)
    if want("solid_fcc_lih_221"):
        solid_case("solid_fcc_lih_221", "fcc_lih_221", dict(ndets=2, hidden_dims_single=[16, 16], hidden_dims_double=[8, 8]), 2)
    # geometry/pbc.py options: polynomial `nu` distance (4 features per pair) and over-complete direction sets
    small = dict(ndets=2, hidden_dims_single=[16, 16], hidden_dims_double=[8, 8])
    for nm, kind, opts in (("solid_cubic_h2_nu", "cubic_h2", dict(distance_type="nu")),
                           ("solid_fcc_lih_221_tri_fcc", "fcc_lih_221", dict(sym_type="fcc")),
                           ("solid_fcc_lih_221_nu_bcc", "fcc_lih_221", dict(distance_type="nu", sym_type="bcc")),
                           ("solid_cubic_h2_tri_hexagonal", "cubic_h2", dict(sym_type="hexagonal"))):
        if want(nm):
            solid_case(nm, kind, dict(small, **opts), 2)
    if want("ewald"):
        madelung_case()
    if want("mcmc_lih"):
        mcmc_case("mcmc_lih", pbc=False)
    if want("mcmc_pbc"):
        mcmc_case("mcmc_pbc", pbc=True)
    if want("observables_lih"):
        observables_case("observables_lih", "LiH", dict(ndets=3, hidden_dims_single=[16, 16], hidden_dims_double=[8, 8]), 3)
    if want("observables_li"):
        observables_case("observables_li", "Li", dict(ndets=2, hidden_dims_single=[16, 16], hidden_dims_double=[8, 8]), 3)


if __name__ == "__main__":
    main()
