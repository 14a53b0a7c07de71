"""TEST INFRASTRUCTURE -- float64 restatement of the psi-ratio consumers: ``SpinSquared``
(``estimator/spin.py:75-146``) and the non-local ECP integral (``estimator/ecp/nonlocal_integral.py:23-165``) for one
walker, on top of any ``phase_logpsi(electrons) -> (sign, log|psi|)`` callable."""

from __future__ import annotations

import math

import torch

F64 = torch.float64


def spin_squared(phase_logpsi, electrons: torch.Tensor, n_up: int, n_down: int) -> torch.Tensor:
    """``S_z (S_z + 1) + sum_{i in minority} (1 - sum_{j in majority} psi(swap_ij) / psi)``; ties: minority = up."""
    n = n_up + n_down
    if n_up > n_down:
        majority, minority = list(range(n_up)), list(range(n_up, n))
    else:
        majority, minority = list(range(n_up, n)), list(range(n_up))
    sz = abs(n_up - n_down) * 0.5
    s0, lp0 = phase_logpsi(electrons)
    total = torch.zeros((), dtype=F64)
    for i in minority:
        ratio_sum = torch.zeros((), dtype=F64)
        for j in majority:
            perm = list(range(n))
            perm[i], perm[j] = j, i
            s, lp = phase_logpsi(electrons[perm])
            ratio_sum = ratio_sum + s * s0 * torch.exp(lp - lp0)
        total = total + (1.0 - ratio_sum)
    return sz * (sz + 1.0) + total


def legendre(x: torch.Tensor, num_l: int):
    return [torch.ones_like(x), x, (3 * x ** 2 - 1) / 2][:num_l]


def nonlocal_integral(phase_logpsi, electrons, atom_positions, quad_points, coefs, n_nonlocal):
    """Per-channel integrals ``(n_elec, n_atoms, n_nonlocal)`` (open boundaries): for electron e and atom I the electron
    is moved to ``atom + |r_eI| * p`` for every rotated quadrature point p of that electron (``quad_points``
    (n_elec, P, 3)), ``integral_l = (2l + 1) sum_p coef_p P_l(cos theta_p) psi'/psi``."""
    n, A = electrons.shape[0], atom_positions.shape[1]
    s0, lp0 = phase_logpsi(electrons)
    out = torch.zeros(n, A, n_nonlocal, dtype=F64)
    for e in range(n):
        for a in range(A):
            rv = electrons[e] - atom_positions[e, a]
            r = rv.norm()
            rdir = rv / r
            ratios = []
            for p in range(quad_points.shape[1]):
                moved = electrons.clone()
                moved[e] = atom_positions[e, a] + r * quad_points[e, p]
                s, lp = phase_logpsi(moved)
                ratios.append(s * s0 * torch.exp(lp - lp0))
            ratios = torch.stack(ratios)
            cos_t = quad_points[e] @ rdir
            for l, pl in enumerate(legendre(cos_t, n_nonlocal)):
                out[e, a, l] = (pl * ratios * coefs).sum() * 4 * math.pi * (2 * l + 1) / (4 * math.pi)
    return out
