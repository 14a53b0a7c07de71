"""Pins the float64 oracle against every known-answer check the reference's own tests hold for the
hot path.  CPU only.

Mirrors, by construction rather than by code:
  * tests/estimator/kinetic_forward_laplacian_test.py:42-85   (analytic Gaussian)
  * tests/estimator/kinetic_forward_laplacian_test.py:254-524 (real nets, FL vs second route)
  * tests/laplacian/primitives/slogdet_test.py:28-173         (slogdet rule vs Hessian)
  * tests/estimator/ewald_test.py:72-152                      (Madelung constants)
  * tests/wavefunction/molecule_wavefunction_test.py:52-82    (antisymmetry)
  * tests/wavefunction/solid_test.py:112-137                  (PBC translation invariance)
  * tests/hydrogen/atom_test.py                               (hydrogen closed form)
"""

import math

import numpy as np
import pytest
import torch

from oracle import estimators as E
from oracle import lap as L
from oracle import networks as N
from oracle import supercell as SC

F64 = torch.float64


def two_electron_walker():
    """The reference's fixed walker (kinetic_forward_laplacian_test.py:31-39)."""
    el = torch.tensor([[0.7, -0.2, 0.3], [-0.4, 0.5, -0.6]], dtype=F64)
    atoms = torch.zeros(1, 3, dtype=F64)
    return el, atoms


@pytest.mark.parametrize(
    "n_particles,n_dims,coeff",
    [(1, 1, 0.5), (3, 3, 0.5), (2, 3, 0.3), (5, 3, 0.3), (3, 2, 0.3), (10, 1, 0.3), (20, 3, 0.2)],
)
def test_gaussian_kinetic_closed_form(n_particles, n_dims, coeff):
    g = torch.Generator().manual_seed(42 + n_particles * 10 + n_dims)
    x = torch.randn(n_particles, n_dims, generator=g, dtype=F64)

    def logpsi(p):
        return L.sum_(L.sum_(L.square(p), dim=-1), dim=0) * (-coeff)

    exact = coeff * n_particles * n_dims - 2 * coeff**2 * float((x * x).sum())
    for route in ("forward", "brute"):
        ke = float(E.kinetic_energy(logpsi, x, route=route))
        assert np.isclose(ke, exact, rtol=1e-10, atol=1e-12), (route, ke, exact)


@pytest.mark.parametrize("batch,n", [((), 3), ((4,), 5), ((2, 3), 2)])
def test_slogdet_rule_vs_hessian(batch, n):
    g = torch.Generator().manual_seed(7)
    k = 6
    w = torch.randn(k, *batch, n, n, generator=g, dtype=F64)
    b = torch.randn(*batch, n, n, generator=g, dtype=F64) + 2 * torch.eye(n, dtype=F64)
    x0 = torch.randn(k, generator=g, dtype=F64)

    def mats(x):  # nonlinear map R^k -> batch of matrices, so A_L != 0
        xs = L.reshape(x, k, *([1] * (len(batch) + 2)))
        return L.sum_(L.tanh(xs * w) + xs * xs * 0.1 * w, dim=0) + b

    def total(x):
        _, ld = L.logdet(mats(x))
        while L.value(ld).dim() > 0:
            ld = L.sum_(ld, dim=0)
        return ld

    _, gf, lf = E.forward_laplacian(total, x0)
    _, gb, lb = L.brute_force(total, x0)
    assert torch.allclose(gf, gb, rtol=1e-9, atol=1e-11)
    assert np.isclose(float(lf), float(lb), rtol=1e-9, atol=1e-10)


def test_slogdet_rule_complex():
    g = torch.Generator().manual_seed(11)
    n, k = 3, 4
    wr = torch.randn(k, n, n, generator=g, dtype=F64)
    wi = torch.randn(k, n, n, generator=g, dtype=F64)
    b = torch.randn(n, n, generator=g, dtype=F64) + 2 * torch.eye(n, dtype=F64)
    x0 = torch.randn(k, generator=g, dtype=F64)

    def total(x):
        xs = L.reshape(x, k, 1, 1)
        re = L.sum_(L.tanh(xs * wr), dim=0) + b
        im = L.sum_(L.sin(xs * wi), dim=0)
        a = L.to_complex(re) + L.to_complex(im) * 1j
        return L.logdet(a)[1]

    v, gf, lf = E.forward_laplacian(total, x0)
    vb, gb, lb = L.brute_force(total, x0)
    assert torch.allclose(v, vb)
    assert torch.allclose(gf, gb, rtol=1e-9, atol=1e-11)
    assert torch.allclose(lf, lb, rtol=1e-9, atol=1e-10)


def _check_two_routes(fn, el):
    v, gf, lf = E.forward_laplacian(fn, el)
    vb, gb, lb = L.brute_force(fn, el)
    assert torch.allclose(torch.as_tensor(v), torch.as_tensor(vb), rtol=1e-12, atol=1e-12)
    assert torch.allclose(gf, gb, rtol=1e-8, atol=1e-10)
    assert torch.allclose(torch.as_tensor(lf), torch.as_tensor(lb), rtol=1e-8, atol=1e-9)
    kf = E.kinetic_energy(fn, el, "forward")
    kb = E.kinetic_energy(fn, el, "brute")
    assert torch.allclose(kf, kb, rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("nspins", [(1, 1), (2, 0)])
def test_ferminet_forward_laplacian_matches_hessian(nspins):
    el, atoms = two_electron_walker()
    p = N.init_ferminet_params(nspins, 1, ndets=2, hidden_single=(8, 8), hidden_double=(4, 4), seed=123)
    _check_two_routes(lambda e: N.ferminet_logpsi(p, e, atoms, nspins)[1], el)


def test_ferminet_three_electron_two_atoms():
    g = torch.Generator().manual_seed(3)
    atoms = torch.tensor([[0.0, 0.0, -0.7], [0.0, 0.0, 0.7]], dtype=F64)
    el = torch.randn(3, 3, generator=g, dtype=F64)
    p = N.init_ferminet_params((2, 1), 2, ndets=3, hidden_single=(8, 8, 8), hidden_double=(4, 4, 4), seed=5)
    _check_two_routes(lambda e: N.ferminet_logpsi(p, e, atoms, (2, 1))[1], el)


def test_lapnet_forward_laplacian_matches_hessian():
    el, atoms = two_electron_walker()
    p = N.init_lapnet_params((1, 1), 1, ndets=2, num_layers=2, heads=2, heads_dim=4, seed=123)
    _check_two_routes(lambda e: N.lapnet_logpsi(p, e, atoms, (1, 1), heads=2)[1], el)


def test_psiformer_forward_laplacian_matches_hessian():
    el, atoms = two_electron_walker()
    p = N.init_psiformer_params((1, 1), 1, ndets=2, num_layers=2, heads=2, heads_dim=4, mlp_hidden=(8,), seed=123)
    _check_two_routes(lambda e: N.psiformer_logpsi(p, e, atoms, (1, 1))[1], el)


def _solid_setup():
    a = 4.0
    prim = (torch.ones(3, 3, dtype=F64) - torch.eye(3, dtype=F64)) * a / 2
    S = np.eye(3)
    sim = torch.tensor(S, dtype=F64) @ prim
    prim_atoms = torch.tensor([[0.0, 0.0, 0.0], [a / 2, a / 2, a / 2]], dtype=F64)
    rec = SC.get_reciprocal_vectors(prim.numpy())
    kpts = SC.get_supercell_kpts(S, rec)
    klist = torch.tensor(np.concatenate([kpts, kpts], 0), dtype=F64)  # one orbital per spin
    return prim, sim, prim_atoms, klist


def test_solid_forward_laplacian_matches_hessian_and_translation_invariance():
    prim, sim, prim_atoms, klist = _solid_setup()
    g = torch.Generator().manual_seed(9)
    el = torch.randn(2, 3, generator=g, dtype=F64)
    p = N.init_solid_params((1, 1), 2, ndets=2, hidden_single=(8, 8), hidden_double=(4, 4), seed=2)

    def fn(e):
        return N.solid_logpsi(p, e, prim_atoms, (1, 1), sim, prim, klist)

    _check_two_routes(fn, el)
    # translation by a simulation-cell lattice vector leaves |psi| unchanged (solid_test.py:112-137)
    shifted = el.clone()
    shifted[0] += sim[1]
    assert np.isclose(float(fn(el).real), float(fn(shifted).real), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("net", ["ferminet", "lapnet", "psiformer"])
def test_antisymmetry(net):
    g = torch.Generator().manual_seed(1)
    atoms = torch.tensor([[0.0, 0.0, -0.7], [0.0, 0.0, 0.7]], dtype=F64)
    el = torch.randn(4, 3, generator=g, dtype=F64)
    nspins = (2, 2)
    if net == "ferminet":
        p = N.init_ferminet_params(nspins, 2, ndets=2, hidden_single=(8, 8), hidden_double=(4, 4), seed=1)
        fn = lambda e: N.ferminet_logpsi(p, e, atoms, nspins)  # noqa: E731
    elif net == "lapnet":
        p = N.init_lapnet_params(nspins, 2, ndets=2, num_layers=2, heads=2, heads_dim=4, seed=1)
        fn = lambda e: N.lapnet_logpsi(p, e, atoms, nspins, heads=2)  # noqa: E731
    else:
        p = N.init_psiformer_params(nspins, 2, ndets=2, num_layers=2, heads=2, heads_dim=4, mlp_hidden=(8,), seed=1)
        fn = lambda e: N.psiformer_logpsi(p, e, atoms, nspins)  # noqa: E731
    s1, l1 = fn(el)
    sw = el.clone()
    sw[[0, 1]] = sw[[1, 0]]
    s2, l2 = fn(sw)
    assert float(s1) == -float(s2)
    assert np.isclose(float(l1), float(l2), rtol=1e-10)
    # opposite-spin exchange is not a symmetry
    sw2 = el.clone()
    sw2[[0, 2]] = sw2[[2, 0]]
    assert not np.isclose(float(fn(sw2)[1]), float(l1), rtol=1e-6)


def test_lapnet_attention_is_softmax_formula():
    g = torch.Generator().manual_seed(2)
    n, h, d = 5, 2, 4
    q, k, v = (torch.randn(n, h, d, generator=g, dtype=F64) for _ in range(3))
    out = N.attention_core(q, k, v)
    logits = torch.einsum("ihd,jhd->hij", q, k) / math.sqrt(d)
    ref = torch.einsum("hij,jhd->ihd", torch.softmax(logits, dim=-1), v)
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)


def _ewald_energy(lattice, atom_coords, atom_charges, S, electron_coords):
    lattice = np.asarray(lattice, dtype=np.float64)
    S = np.asarray(S, dtype=np.float64)
    sup = S @ lattice
    scale = round(float(np.linalg.det(S)))
    tr = SC.get_supercell_copies(lattice, S)
    atoms = (np.asarray(atom_coords)[None] + tr[:, None]).reshape(-1, 3)
    charges = np.tile(np.asarray(atom_charges, dtype=np.float64), scale)
    ew = E.EwaldSum(sup)
    return E.solid_potential_energy(ew, electron_coords, atoms, charges)


def test_madelung_nacl_primitive():
    Lc = 2.0
    lattice = (np.ones((3, 3)) - np.eye(3)) * Lc / 2
    e = _ewald_energy(lattice, [[0.0, 0.0, 0.0]], [1.0], np.eye(3), [[Lc / 2, Lc / 2, Lc / 2]])
    assert abs(e + 1.74756) < 1e-4


def test_madelung_nacl_conventional():
    Lc = 2.0
    lattice = (np.ones((3, 3)) - np.eye(3)) * Lc / 2
    S = np.ones((3, 3)) - 2 * np.eye(3)
    el = [[Lc / 2, Lc / 2, Lc / 2], [Lc / 2, 0, 0], [0, Lc / 2, 0], [0, 0, Lc / 2]]
    e = _ewald_energy(lattice, [[0.0, 0.0, 0.0]], [1.0], S, el)
    assert abs(e / 4 + 1.74756) < 1e-4


def test_madelung_caf2():
    Lc = 4 / np.sqrt(3)
    lattice = (np.ones((3, 3)) - np.eye(3)) * Lc / 2
    el = [[Lc / 4, Lc / 4, Lc / 4], [Lc / 4, -Lc / 4, Lc / 4]]
    e = _ewald_energy(lattice, [[0.0, 0.0, 0.0]], [2.0], np.eye(3), el)
    assert abs(e + 5.03879) < 1e-4


def test_hydrogen_closed_form_and_exact_ground_state():
    g = torch.Generator().manual_seed(0)
    r = torch.randn(6, 1, 3, generator=g, dtype=F64)
    for alpha in (-1.0, -0.8):
        for w in range(r.shape[0]):
            ke = E.kinetic_energy(lambda e: N.hydrogen_logpsi(alpha, e), r[w])
            pe = -1.0 / r[w].norm()
            closed = E.hydrogen_local_energy(alpha, r[w : w + 1])
            assert np.isclose(float(ke + pe), float(closed[0]), rtol=1e-10)
            if alpha == -1.0:  # exact ground state: E_L = -0.5 everywhere (tests/hydrogen/atom_test.py)
                assert np.isclose(float(ke + pe), -0.5, atol=1e-10)


def test_coulomb_potential_hand_value():
    # H2-like: two protons at +-0.7 on z, electrons on the axis
    atoms = torch.tensor([[0.0, 0.0, -0.7], [0.0, 0.0, 0.7]], dtype=F64)
    el = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 1.7]], dtype=F64)
    ch = torch.tensor([1.0, 1.0], dtype=F64)
    v = float(E.potential_energy(el, atoms, ch))
    expect = -(1 / 0.7 + 1 / 0.7 + 1 / 2.4 + 1 / 1.0) + 1 / 1.7 + 1 / 1.4
    assert np.isclose(v, expect, rtol=1e-12)


def test_mh_accept_rule_and_width_adaptation():
    g = torch.Generator().manual_seed(4)
    W, n = 64, 2
    x = torch.randn(W, n, 3, generator=g, dtype=F64)
    steps = 3
    normals = torch.randn(steps, W, n, 3, generator=g, dtype=F64)
    uniforms = torch.rand(steps, W, generator=g, dtype=F64)

    def blp(xx):  # 2*log|psi| for a Gaussian
        return -(xx * xx).sum((-1, -2))

    state = (torch.tensor(0.3, dtype=F64), torch.zeros(100, dtype=F64), 0)
    x1, pmove, st, acc = E.mcmc_step(blp, x, normals, uniforms, state, steps=steps)
    # replay by hand
    xr, lp = x.clone(), blp(x)
    nacc = 0
    for s in range(steps):
        x2 = xr + normals[s] * 0.3
        lp2 = blp(x2)
        c = (lp2 - lp) > torch.log(uniforms[s])
        assert torch.equal(c, acc[s])
        xr = torch.where(c[:, None, None], x2, xr)
        lp = torch.where(c, lp2, lp)
        nacc += int(c.sum())
    assert torch.equal(xr, x1)
    assert np.isclose(pmove, nacc / (steps * W))
    assert st[2] == 1 and float(st[1][1]) == pmove and float(st[0]) == 0.3
    # adaptation fires when counter % adapt_frequency == 0
    st99 = (torch.tensor(0.3, dtype=F64), torch.full((100,), 0.9, dtype=F64), 99)
    _, _, st100, _ = E.mcmc_step(blp, x, normals, uniforms, st99, steps=steps)
    assert np.isclose(float(st100[0]), 0.33)
