"""``SpinSquared`` estimator: host-side mirror of ``estimator/spin.py:24-146`` over the psi-ratio entry point
(``jaqmc_b200_psi_ratios``: every swapped configuration of every walker in one batched value-only forward pass).

    S^2_local = S_z (S_z + 1) + n_min - sum_{i in minority} sum_{j in majority} psi(r_{i <-> j}) / psi(r)
"""

from __future__ import annotations

from dataclasses import dataclass

import torch

from ._runtime import runtime


@dataclass
class SpinSquared:
    """Same fields as the reference (``n_up``, ``n_down``, ``phase_logpsi`` -- here the wavefunction object --,
    ``data_field``); ``evaluate_batch_walkers`` returns ``{"spin:s2": (W,)}``."""

    n_up: int = 0
    n_down: int = 0
    phase_logpsi: object = None
    data_field: str = "electrons"

    def init(self, data=None, rngs=None) -> None:
        n_up, n_down = int(self.n_up), int(self.n_down)
        n = n_up + n_down
        # ties: minority = up (estimator/spin.py:62-68)
        if n_up > n_down:
            self._majority_idx, self._minority_idx = list(range(n_up)), list(range(n_up, n))
        else:
            self._majority_idx, self._minority_idx = list(range(n_up, n)), list(range(n_up))
        self._sz = abs(n_up - n_down) * 0.5
        return None

    def evaluate_batch_walkers(self, params, data, prev_walker_stats=None, state=None, rngs=None):
        if not hasattr(self, "_sz"):
            self.init()
        wf = self.phase_logpsi
        x = getattr(data, self.data_field).contiguous()
        W, dev = x.shape[0], x.device
        mi, ma = self._minority_idx, self._majority_idx
        s2 = torch.full((W,), self._sz * (self._sz + 1.0), dtype=torch.float32, device=dev)
        if not mi or not ma:
            return {"spin:s2": s2 + float(len(mi))}, state
        # move q = (i, j): electron i goes to x_j and electron j to x_i
        pairs = [(i, j) for i in mi for j in ma]
        idx = torch.tensor(pairs, dtype=torch.int32, device=dev)
        pos = torch.stack([x[:, idx[:, 1].long()], x[:, idx[:, 0].long()]], dim=2).contiguous()   # (W, Q, 2, 3)
        handle, sysh = wf._sampling_handles(params, data)
        log_ratio, sign_ratio = runtime(dev).psi_ratios(handle, sysh, x, idx, pos)
        ratio = sign_ratio * torch.exp(log_ratio)                      # psi(swap_ij) / psi
        per_minority = 1.0 - ratio.reshape(W, len(mi), len(ma)).sum(dim=2)
        return {"spin:s2": s2 + per_minority.sum(dim=1)}, state
