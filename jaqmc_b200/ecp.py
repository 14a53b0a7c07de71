"""Non-local ECP integral: host-side mirror of ``estimator/ecp/nonlocal_integral.py`` (``make_nonlocal_integral``,
``legendre_polynomials``) and ``estimator/ecp/quadrature.py`` (``Quadrature``, ``Octahedron``, ``Icosahedron``,
``get_quadrature``) over the psi-ratio entry point: the wavefunction ratios at ALL displaced configurations of all
walkers (electron x atom x quadrature point) are one batched value-only forward pass on the CUDA kernels; the angular
weights and the quadrature sum are a few elementwise torch operations on the result.
"""

from __future__ import annotations

import math

import torch

from ._runtime import runtime

DEFAULT_QUADRATURE_ID = "icosahedron_12"


def _expand_sign(values):
    """All sign permutations of a point, zeros not flipped (quadrature.py:21-41)."""
    if len(values) == 1:
        return [[values[0]]] + ([[-values[0]]] if values[0] != 0 else [])
    rest = _expand_sign(values[1:])
    out = [[values[0], *r] for r in rest]
    if values[0] != 0:
        out.extend([[-values[0], *r] for r in rest])
    return out


class Quadrature:
    """Points ``pts`` (P, 3) and weights ``coefs`` (P,) on the unit sphere (quadrature.py:44-131)."""

    def __init__(self, n_points: int) -> None:
        self.n_points = n_points

    @staticmethod
    def rotation_matrices(phi: torch.Tensor, cos_theta: torch.Tensor) -> torch.Tensor:
        """The reference's rotation from its two uniform draws ``phi = 2 pi u1``, ``cos_theta = 1 - 2 u2``
        (quadrature.py:59-100)."""
        sin_theta = torch.sqrt(1.0 - cos_theta ** 2)
        sp, cp = torch.sin(phi), torch.cos(phi)
        m11 = sp ** 2 + cos_theta * cp ** 2
        m12 = sp * cp * (cos_theta - 1)
        m13 = sin_theta * cp
        m22 = cp ** 2 + cos_theta * sp ** 2
        m23 = sin_theta * sp
        return torch.stack([m11, m12, m13, m12, m22, m23, -m13, -m23, cos_theta], dim=-1).reshape(*phi.shape, 3, 3)

    def sample_rotated_points(self, shape, rngs, device) -> torch.Tensor:
        """``(*shape, P, 3)`` randomly rotated points; ``rngs`` is a ``torch.Generator`` or a ``(u1, u2)`` pair of
        uniforms of shape ``shape``."""
        if isinstance(rngs, (tuple, list)):
            u1, u2 = rngs
        else:
            u1 = torch.rand(*shape, generator=rngs, device=device)
            u2 = torch.rand(*shape, generator=rngs, device=device)
        rot = self.rotation_matrices(2 * math.pi * u1, 1.0 - 2.0 * u2)
        return torch.einsum("...jk,lk->...lj", rot, self.pts.to(rot))

    def integrate(self, values: torch.Tensor) -> torch.Tensor:
        return (values * self.coefs.to(values)).sum(dim=-1) * 4 * math.pi


class Octahedron(Quadrature):
    _coef_table = {
        6: [1.0 / 6.0] * 6,
        18: [1.0 / 30.0] * 6 + [1.0 / 15.0] * 12,
        26: [1.0 / 21.0] * 6 + [4.0 / 105.0] * 12 + [27.0 / 840.0] * 8,
        50: [4.0 / 315.0] * 6 + [64.0 / 2835.0] * 12 + [27.0 / 1280.0] * 8 + [14641.0 / 725760.0] * 24,
    }

    def __init__(self, n_points: int) -> None:
        super().__init__(n_points)
        if n_points not in self._coef_table:
            raise ValueError(f"Octahedron quadrature supports 6, 18, 26, or 50 points, got {n_points}")
        p, q, r, s = 1 / math.sqrt(2), 1 / math.sqrt(3), 1 / math.sqrt(11), 3 / math.sqrt(11)
        pts = []
        for v in ([1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [p, p, 0.0], [p, 0.0, p], [0.0, p, p], [q, q, q],
                  [r, r, s], [r, s, r], [s, r, r]):
            pts.extend(_expand_sign(v))
        self.pts = torch.tensor(pts[:n_points], dtype=torch.float64)
        self.coefs = torch.tensor(self._coef_table[n_points], dtype=torch.float64)


class Icosahedron(Quadrature):
    _coef_table = {12: [1.0 / 12.0] * 12, 32: [5.0 / 168.0] * 12 + [27.0 / 840.0] * 20}

    def __init__(self, n_points: int) -> None:
        super().__init__(n_points)
        if n_points not in self._coef_table:
            raise ValueError(f"Icosahedron quadrature supports 12 or 32 points, got {n_points}")
        pi = math.pi
        polars = [(0.0, 0.0), (pi, 0.0)]
        at2 = math.atan(2.0)
        polars += [(at2, 2 * k * pi / 5) for k in range(5)]
        polars += [(pi - at2, (2 * k + 1) * pi / 5) for k in range(5)]
        down = math.sqrt(15 + 6 * math.sqrt(5.0))
        th1, th2 = math.acos((2 + math.sqrt(5.0)) / down), math.acos(1.0 / down)
        polars += [(th1, (2 * k + 1) * pi / 5) for k in range(5)]
        polars += [(th2, (2 * k + 1) * pi / 5) for k in range(5)]
        polars += [(pi - th1, 2 * k * pi / 5) for k in range(5)]
        polars += [(pi - th2, 2 * k * pi / 5) for k in range(5)]
        self.pts = torch.tensor([[math.sin(t) * math.cos(f), math.sin(t) * math.sin(f), math.cos(t)]
                                 for t, f in polars[:n_points]], dtype=torch.float64)
        self.coefs = torch.tensor(self._coef_table[n_points], dtype=torch.float64)


def get_quadrature(quadrature_id: str | None = None) -> Quadrature:
    quadrature_id = quadrature_id or DEFAULT_QUADRATURE_ID
    parts = quadrature_id.split("_")
    if len(parts) != 2:
        raise ValueError(f"Invalid quadrature_id format: {quadrature_id}. Expected format: 'type_npoints' (e.g., 'icosahedron_12')")
    kinds = {"icosahedron": Icosahedron, "octahedron": Octahedron}
    if parts[0] not in kinds:
        raise ValueError(f"Unknown quadrature type: {parts[0]}")
    return kinds[parts[0]](int(parts[1]))


def legendre_polynomials(x: torch.Tensor, num_l: int) -> torch.Tensor:
    """``P_0 .. P_{num_l - 1}`` stacked on a new leading axis (nonlocal_integral.py:187-214)."""
    if num_l > 3:
        raise ValueError(f"Legendre polynomials up to l=2 are supported, but num_l={num_l} was requested.")
    polys = [torch.ones_like(x), x, (3 * x ** 2 - 1) / 2]
    return torch.stack(polys[:num_l]) if num_l > 0 else x.new_zeros((0, *x.shape))


def make_nonlocal_integral(num_channels: int, quadrature: Quadrature, lattice=None, twist=None):
    """Batched counterpart of the reference factory (nonlocal_integral.py:23-165).  The returned callable is

        evaluate(wf, params, data, atom_positions, rngs) -> (W, n_elec, n_atoms, n_nonlocal)

    with ``atom_positions`` (W, n_elec, n_atoms, 3) the nearest atom image of every electron-atom pair and ``rngs`` a
    ``torch.Generator`` or the ``(u1, u2)`` uniforms (W, n_elec) of the per-electron random rotations."""
    n_nonlocal = num_channels - 1
    lat = None if lattice is None else torch.as_tensor(lattice, dtype=torch.float64)
    tw = None if twist is None else torch.as_tensor(twist, dtype=torch.float64)

    def evaluate(wf, params, data, atom_positions: torch.Tensor, rngs) -> torch.Tensor:
        x = data.electrons.contiguous()
        W, n = x.shape[0], x.shape[1]
        A = atom_positions.shape[2]
        dev = x.device
        if A == 0 or n_nonlocal == 0:
            return x.new_zeros((W, n, A, n_nonlocal))
        pts = quadrature.sample_rotated_points((W, n), rngs, dev).to(x.dtype)          # (W, n, P, 3)
        P = pts.shape[2]
        r_vec = x[:, :, None, :] - atom_positions                                        # (W, n, A, 3)
        r = r_vec.norm(dim=-1)
        r_dir = r_vec / r[..., None]
        disp = atom_positions[:, :, :, None, :] + r[..., None, None] * pts[:, :, None, :, :]   # (W, n, A, P, 3)
        bloch = None
        if lat is not None:
            l32 = lat.to(x)
            inv = torch.linalg.inv(l32)
            frac = disp @ inv
            wrapped = (frac - torch.floor(frac)) @ l32                                   # geometry/pbc.py:97-111
            if tw is not None:
                shift = (disp - wrapped) @ inv
                kdot = 2 * math.pi * (shift @ torch.remainder(tw.to(x), 1.0))
                bloch = torch.polar(torch.ones_like(kdot), kdot)
            disp = wrapped
        # moves: electron e displaced (one electron per move), ordered (e, atom, point)
        idx = torch.full((n * A * P, 2), -1, dtype=torch.int32, device=dev)
        idx[:, 0] = torch.arange(n, device=dev, dtype=torch.int32).repeat_interleave(A * P)
        pos = torch.zeros(W, n * A * P, 2, 3, dtype=x.dtype, device=dev)
        pos[:, :, 0, :] = disp.reshape(W, n * A * P, 3)
        handle, sysh = wf._sampling_handles(params, data)
        log_ratio, sign_ratio = runtime(dev).psi_ratios(handle, sysh, x, idx, pos.contiguous())
        complex_wf = type(wf).__name__ == "SolidWavefunction"
        mag = torch.exp(log_ratio).reshape(W, n, A, P)
        sgn = sign_ratio.reshape(W, n, A, P)
        ratios = torch.polar(mag, sgn) if complex_wf else mag * sgn                      # phase difference / sign product
        if bloch is not None:
            ratios = ratios * bloch
        cos_theta = torch.einsum("wepk,weak->weap", pts, r_dir)
        pl = legendre_polynomials(cos_theta, n_nonlocal)                                 # (L, W, n, A, P)
        out = []
        for l in range(n_nonlocal):
            out.append(quadrature.integrate(pl[l] * ratios) * (2 * l + 1))
        return torch.stack(out, dim=-1) / (4 * math.pi)

    return evaluate
