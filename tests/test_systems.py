"""Synthetic benchmark systems (jaqmc_b200/systems.py): geometry sanity of the BASELINE configurations."""

import numpy as np

from jaqmc_b200 import systems as S


def test_molecules_are_neutral_and_sized():
    for name, n, A in (("Li", 3, 1), ("N2", 14, 2), ("C6H6", 42, 12), ("H", 1, 1)):
        atoms, charges, nspins = S.molecule(name)
        assert atoms.shape == (A, 3) and sum(nspins) == n == int(charges.sum())
    atoms, _, _ = S.molecule("C6H6")
    cc = np.linalg.norm((atoms[0] - atoms[1]).numpy())
    ch = np.linalg.norm((atoms[0] - atoms[6]).numpy())
    assert abs(cc - 1.397 * S.ANGSTROM) < 1e-9 and abs(ch - 1.084 * S.ANGSTROM) < 1e-9


def test_lih_supercell_222():
    prim, sim, patoms, cell_atoms, cell_charges, nspins, klist = S.solid_system("fcc_lih_222")
    assert cell_atoms.shape == (16, 3) and nspins == (16, 16) and klist.shape == (32, 3)
    assert abs(np.linalg.det(sim) - 8 * np.linalg.det(prim)) < 1e-9 and cell_charges.sum() == 32
    # every k-point folds onto the supercell's Gamma point: exp(i k . T) = 1 for simulation lattice vectors T
    assert np.allclose(np.exp(1j * klist @ sim.T), 1.0, atol=1e-9)
    el = S.solid_walkers(cell_atoms, 32, 5, seed=1)
    assert el.shape == (5, 32, 3) and el.dtype == np.float32


def test_synthetic_walkers_are_float32_exact_and_seeded():
    atoms, charges, nspins = S.molecule("N2")
    a = S.synthetic_walkers(atoms, charges, nspins, 4, seed=3)
    b = S.synthetic_walkers(atoms, charges, nspins, 4, seed=3)
    assert (a == b).all() and (a.float().double() == a).all()
