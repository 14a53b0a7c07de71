"""Synthetic systems of the benchmark configurations (BASELINE.json ``configs``): geometries from the reference's
config builders and seeded walker batches.  Product-side data only -- nothing here touches ``oracle/``.

* molecules: ``app/molecule/config/{atom,diatomic}.py`` (N2 needs ``bond_length=2.068`` bohr, the default 1.4 is H2's);
  benzene has no preset in the reference (custom ``system.atoms`` list): standard D6h geometry, C-C 1.397 A, C-H 1.084 A
* LiH rock salt: ``app/solid/config/rock_salt.py:26-77`` with ``lattice_constant = 4.0`` A and a diagonal supercell
* walkers: electrons ~ N(atom position, 1) as ``initialize_electrons_gaussian`` (``app/molecule/data.py:36-44``)
"""

from __future__ import annotations

import math

import numpy as np
import torch

F64 = torch.float64
ANGSTROM = 1.8897261246


def molecule(name):
    """(atoms (A,3) float64 [bohr], charges (A,), nspins)."""
    if name == "Li":
        return torch.zeros(1, 3, dtype=F64), torch.tensor([3.0], dtype=F64), (2, 1)
    if name == "Li3up":  # spin-polarised lithium: a single spin channel with several electrons (test geometry)
        return torch.zeros(1, 3, dtype=F64), torch.tensor([3.0], dtype=F64), (3, 0)
    if name == "H":
        return torch.zeros(1, 3, dtype=F64), torch.tensor([1.0], dtype=F64), (1, 0)
    if name == "He":
        return torch.zeros(1, 3, dtype=F64), torch.tensor([2.0], dtype=F64), (1, 1)
    if name == "LiH":
        return (torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 3.015]], dtype=F64), torch.tensor([3.0, 1.0], dtype=F64), (2, 2))
    if name == "Ar":  # 18 electrons: exercises the n > 16 code paths (generic LogDet kernel)
        return torch.zeros(1, 3, dtype=F64), torch.tensor([18.0], dtype=F64), (9, 9)
    if name in ("Zn", "As", "Cd"):  # 30 / 33 / 48 electrons on one nucleus: tile-shape edge cases of the attention kernels
        z = {"Zn": 30, "As": 33, "Cd": 48}[name]
        return torch.zeros(1, 3, dtype=F64), torch.tensor([float(z)], dtype=F64), ((z + 1) // 2, z // 2)
    if name == "N2":
        return (torch.tensor([[0.0, 0.0, -1.034], [0.0, 0.0, 1.034]], dtype=F64), torch.tensor([7.0, 7.0], dtype=F64), (7, 7))
    if name == "C6H6":  # 42 electrons, 12 atoms
        rc, rh = 1.397 * ANGSTROM, (1.397 + 1.084) * ANGSTROM
        pos, z = [], []
        for k in range(6):
            a = math.pi / 3 * k
            pos.append([rc * math.cos(a), rc * math.sin(a), 0.0])
            z.append(6.0)
        for k in range(6):
            a = math.pi / 3 * k
            pos.append([rh * math.cos(a), rh * math.sin(a), 0.0])
            z.append(1.0)
        return torch.tensor(pos, dtype=F64), torch.tensor(z, dtype=F64), (21, 21)
    raise KeyError(name)


def synthetic_walkers(atoms, charges, nspins, W, seed=0):
    """Electrons ~ N(atom, 1) assigned to atoms in proportion to nuclear charge (app/molecule/data.py:36-44 style)."""
    g = torch.Generator().manual_seed(seed)
    n = sum(nspins)
    owners = []
    z = charges.clone()
    for _ in range(n):
        i = int(torch.argmax(z))
        owners.append(i)
        z[i] -= 1.0
    centers = atoms[torch.tensor(owners)]
    el = centers[None] + torch.randn(W, n, 3, generator=g, dtype=F64)
    return el.to(torch.float32).to(F64)  # exactly float32-representable


def solid_system(kind):
    """(prim_lattice, sim_lattice, prim_atoms, cell_atoms, cell_charges, nspins, klist) in bohr.  ``klist`` holds one
    k-point per orbital (up orbitals, then down), the k-points that fold onto the supercell's Gamma point."""
    if kind == "cubic_h2":       # simple cubic cell, 2 atoms, supercell = primitive cell, Gamma point
        prim = 3.2 * np.eye(3)
        S = np.eye(3, dtype=int)
        patoms = np.array([[0.0, 0.0, 0.0], [1.4, 0.3, 0.2]])
        z = np.array([1.0, 1.0])
        per_cell = (1, 1)
    elif kind in ("fcc_lih_221", "fcc_lih_222"):  # FCC rock salt, Li at the origin, H at (a/2, a/2, a/2)
        a = 4.4 if kind == "fcc_lih_221" else 4.0 * ANGSTROM
        prim = a / 2 * np.array([[0.0, 1.0, 1.0], [1.0, 0.0, 1.0], [1.0, 1.0, 0.0]])
        S = np.diag([2, 2, 1]) if kind == "fcc_lih_221" else np.diag([2, 2, 2])
        patoms = np.array([[0.0, 0.0, 0.0], [a / 2, a / 2, a / 2]])
        z = np.array([3.0, 1.0])
        per_cell = (1, 1) if kind == "fcc_lih_221" else (2, 2)   # the small test cell keeps one orbital per k-point
    else:
        raise KeyError(kind)
    sim = S @ prim
    ncell = int(round(abs(np.linalg.det(S))))
    shifts = np.array([[i, j, k] for i in range(S[0, 0]) for j in range(S[1, 1]) for k in range(S[2, 2])]) @ prim
    cell_atoms = (patoms[None] + shifts[:, None]).reshape(-1, 3)
    cell_charges = np.tile(z, len(shifts))
    b = 2 * np.pi * np.linalg.inv(prim).T
    frac = np.array([[i / S[0, 0], j / S[1, 1], k / S[2, 2]] for i in range(S[0, 0]) for j in range(S[1, 1])
                     for k in range(S[2, 2])])
    ks = frac @ b
    if kind == "cubic_h2":
        nspins, klist = (1, 1), np.zeros((2, 3))
    elif kind == "fcc_lih_221":
        nspins = (2, 2)
        klist = np.concatenate([ks[[0, 2]], ks[[1, 3]]])
    else:
        nspins = (per_cell[0] * ncell, per_cell[1] * ncell)
        klist = np.concatenate([np.repeat(ks, per_cell[0], axis=0), np.repeat(ks, per_cell[1], axis=0)])
    return prim, sim, patoms, cell_atoms, cell_charges, nspins, klist


def solid_walkers(cell_atoms, n, W, seed=0, sigma=0.8):
    """Electrons ~ N(random atom of the simulation cell, sigma) (app/solid/data.py:70 wraps them into the cell; the
    features are periodic so the wrap does not change any result)."""
    g = np.random.default_rng(seed)
    return (cell_atoms[g.integers(0, len(cell_atoms), (W, n))] + sigma * g.normal(size=(W, n, 3))).astype(np.float32)
