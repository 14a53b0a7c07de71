"""Supercell helpers (oracle; test infrastructure only).

Restates ``utils/supercell.py:26-85`` (folding k-points) and ``:88-124`` (translation vectors) of the
reference in NumPy float64; used to build the synthetic LiH rock-salt 2x2x2 workload and the Madelung
known-answer checks (``tests/estimator/ewald_test.py:10-60``).
"""

from __future__ import annotations

import numpy as np


def get_reciprocal_vectors(lattice):
    return 2 * np.pi * np.linalg.inv(np.asarray(lattice, dtype=np.float64)).T


def _int_box(S):
    u = np.array([0, 1])
    mesh = np.meshgrid(u, u, u, indexing="ij")
    return np.stack([m.ravel() for m in mesh], axis=-1)


def get_supercell_copies(latvec, S):
    """Translation vectors tiling the supercell ``S @ latvec`` with primitive cells."""
    latvec = np.asarray(latvec, dtype=np.float64)
    S = np.asarray(S, dtype=np.float64)
    box = _int_box(S) @ S
    lo, hi = box.min(0), box.max(0)
    ranges = [np.arange(np.floor(a), np.ceil(b)) for a, b in zip(lo, hi)]
    mesh = np.meshgrid(*ranges, indexing="ij")
    pts = np.stack([m.ravel() for m in mesh], axis=-1)
    frac = pts @ np.linalg.inv(S)
    ok = np.all((frac >= -1e-5) & (frac < 1 - 1e-5), axis=1)
    return pts[ok] @ latvec


def get_supercell_kpts(S, recvec):
    """Primitive-cell k-points folding onto the supercell Gamma point."""
    S = np.asarray(S, dtype=np.float64)
    box = _int_box(S) @ S.T
    lo = np.floor(box.min(0)).astype(int)
    hi = np.ceil(box.max(0)).astype(int)
    ranges = [np.arange(a, b + 1) for a, b in zip(lo, hi)]
    mesh = np.meshgrid(*ranges, indexing="ij")
    n = np.stack([m.ravel() for m in mesh], axis=-1)
    kf = n @ np.linalg.inv(S)
    ok = np.all((kf >= -1e-5) & (kf < 1 - 1e-5), axis=1)
    return np.mod(kf[ok], 1.0) @ np.asarray(recvec, dtype=np.float64)
