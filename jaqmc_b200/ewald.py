"""Ewald summation for periodic systems: host-side mirror of ``estimator/ewald.py`` (``EwaldSum``) and of the solid
``PotentialEnergy`` estimator (``app/solid/hamiltonian.py:18-56``).

The one-off setup (Ewald parameter, lattice images, selection of the reciprocal vectors, constants) follows
``estimator/ewald.py:50-110`` and runs on the host in float64 NumPy; the per-walker energy (``:112-173``) is the CUDA
kernel ``k_ewald`` reached through ``jaqmc_b200_ewald``.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi
from ._runtime import runtime


def _minimum_image_kind(lattice: np.ndarray, tol: float = 1e-10) -> int:
    """0 diagonal, 1 orthogonal, 2 general -- the three branches of ``build_distance_fn`` (geometry/pbc.py:128-184)."""
    off_diag = lattice - np.diag(np.diagonal(lattice))
    if np.all(np.abs(off_diag) < tol):
        return 0
    gram_upper = np.triu(lattice @ lattice.T)
    return 1 if np.allclose(gram_upper, 0.0, atol=tol) else 2


class EwaldSum:
    """``EwaldSum(supercell_lattice, ewald_gmax=200, nlatvec=1)``; ``energy(coords, charges)`` of the reference becomes
    :meth:`energy` over a walker batch of electrons plus the fixed ions."""

    def __init__(self, supercell_lattice, ewald_gmax: int = 200, nlatvec: int = 1, device=None, g_weight_tol=1e-12):
        lat = np.asarray(torch.as_tensor(supercell_lattice).detach().cpu().numpy(), dtype=np.float64)
        if lat.shape != (3, 3):
            raise ValueError(f"supercell_lattice: expected (3, 3), got {lat.shape}")
        self.latvec = lat
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        volume = float(np.linalg.det(lat))
        recvec = np.linalg.inv(lat).T
        # separation parameter from the smallest perpendicular height of the cell (ewald.py:80-83)
        self.alpha = 5.0 / float(np.min(1.0 / np.linalg.norm(recvec, axis=1)))
        # real-space images (ewald.py:50-56)
        r = np.arange(-nlatvec, nlatvec + 1)
        idx = np.stack(np.meshgrid(r, r, r, indexing="ij"), axis=-1).reshape(-1, 3)
        images = idx @ lat
        center = int(np.argmin(np.linalg.norm(images, axis=1)))
        # reciprocal vectors: half space {x>0} u {x=0,y>0} u {x=y=0,z>0}, kept where the weight exceeds the tolerance
        # (ewald.py:85-105, select_big_3d :176-200).  The Gaussian factor bounds the useful index range, so the
        # (2*gmax+1)^3 grid is never materialised.
        b = 2.0 * np.pi * recvec
        gmin2 = float(np.min(np.sum(b * b, axis=1)))
        g2max = gmin2
        while 4.0 * np.pi * np.exp(-g2max / (4.0 * self.alpha**2)) / (volume * gmin2) > g_weight_tol * 1e-3:
            g2max *= 1.1
        bound = int(np.ceil(np.sqrt(g2max) * np.max(np.linalg.norm(lat, axis=1)) / (2.0 * np.pi))) + 1
        bound = min(int(ewald_gmax), bound)
        full, pos = np.arange(-bound, bound + 1), np.arange(1, bound + 1)
        zero = np.array([0])
        gp, gw = [], []
        for ix, iy, iz in ((pos, full, full), (zero, pos, full), (zero, zero, pos)):
            m = np.stack(np.meshgrid(ix, iy, iz, indexing="ij"), axis=-1).reshape(-1, 3)
            g = m @ b
            g2 = np.sum(g * g, axis=1)
            wgt = 4.0 * np.pi * np.exp(-g2 / (4.0 * self.alpha**2)) / (volume * g2)
            keep = wgt > g_weight_tol
            gp.append(g[keep])
            gw.append(wgt[keep])
        self.gpoints = np.concatenate(gp, axis=0)
        self.gweight = np.concatenate(gw, axis=0)
        self.ijconst = -np.pi / (volume * self.alpha**2)
        self.self_const_factor = -self.alpha / np.sqrt(np.pi)
        self.cellvolume = volume
        self.mic_kind = _minimum_image_kind(lat)
        # general-cell minimum image: the reference's meshgrid (default 'xy' indexing) order decides argmin ties
        mesh = np.meshgrid(*[np.array([0, 1, 2])] * 3)
        shifts = (np.stack([m.ravel() for m in mesh], axis=0).T - 1) @ lat

        dev = self.device
        f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to(dev)  # noqa: E731
        self._t = dict(lattice=f32(lat), inv_lattice=f32(np.linalg.inv(lat)), mic_shifts=f32(shifts), images=f32(images),
                       gpoints=f32(self.gpoints), gweight=f32(self.gweight))
        s = _abi.Ewald()
        for k, t in self._t.items():
            setattr(s, k, t.data_ptr())
        s.n_images, s.center_image, s.n_g = images.shape[0], center, self.gpoints.shape[0]
        s.mic_kind = self.mic_kind
        s.alpha, s.self_const_factor, s.ijconst = self.alpha, self.self_const_factor, self.ijconst
        self._struct = s

    def energy(self, electrons: torch.Tensor, atoms: torch.Tensor, charges: torch.Tensor, _rt=None) -> torch.Tensor:
        """``electrons`` (W, n, 3) [charge -1], ``atoms`` (A, 3), ``charges`` (A,) -> potential energy (W,)."""
        rt = _rt if _rt is not None else runtime(electrons.device)
        el = electrons.reshape(electrons.shape[0], -1, 3).contiguous()
        for t, nm in ((el, "electrons"), (atoms, "atoms"), (charges, "charges")):
            rt._check_tensor(t, nm)
        W, n = el.shape[0], el.shape[1]
        e_pot = torch.empty(W, dtype=torch.float32, device=el.device)
        rc = rt.lib.jaqmc_b200_ewald(C.byref(self._struct), C.c_void_p(el.data_ptr()), W, n,
                                     C.c_void_p(atoms.data_ptr()), C.c_void_p(charges.data_ptr()), atoms.shape[0],
                                     C.c_void_p(e_pot.data_ptr()), rt._stream())
        _abi.check(rt.lib, rc)
        return e_pot


class SolidPotentialEnergy:
    """Solid-state ``PotentialEnergy`` estimator (app/solid/hamiltonian.py:18-56): electrons and ions in one Ewald sum."""

    def __init__(self, supercell_lattice, device=None):
        self.ewald = EwaldSum(supercell_lattice, device=device)

    def evaluate_batch_walkers(self, params, data, prev_walker_stats=None, state=None, rngs=None):
        return {"energy:potential": self.ewald.energy(data.electrons, data.atoms, data.charges)}, state
