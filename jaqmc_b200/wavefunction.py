"""Wavefunction classes: host-side mirror of the reference's ``Wavefunction`` / ``MoleculeWavefunction``
protocol (``wavefunction/base.py:56-112``, ``app/molecule/wavefunction/base.py:17-72``) over the CUDA kernels.

Differences from the reference, by design of the boundary: the reference evaluates one walker
and is batched by an external ``jax.vmap``; here every method takes the walker batch (``electrons`` (W, n, 3))
and returns per-walker arrays, because the batch is what one kernel launch processes.  Parameter trees keep the
reference's Flax layout, so checkpoints map one-to-one.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

from . import _marshal
from ._runtime import runtime
from .data import MoleculeData


def _batched(electrons: torch.Tensor):
    if electrons.dim() == 2:
        return electrons.unsqueeze(0).contiguous(), True
    if electrons.dim() != 3 or electrons.shape[-1] != 3:
        raise ValueError(f"electrons: expected (W, n, 3) or (n, 3), got {tuple(electrons.shape)}")
    return electrons.contiguous(), False


class Wavefunction:
    """Base class: ``init_params``, ``evaluate`` (reference wavefunction/base.py:71-99), plus the batched
    ``local_energy`` used by the estimators."""

    nspins: tuple

    # -- to be provided by subclasses ----------------------------------------------------------
    def _handle(self, params, n_atoms: int):
        raise NotImplementedError

    def init_params(self, data: MoleculeData, rngs):
        raise NotImplementedError

    # -- protocol --------------------------------------------------------------------------------
    def _check(self, data: MoleculeData):
        n = sum(self.nspins)
        if data.electrons.shape[-2] != n:
            raise ValueError(f"electrons has {data.electrons.shape[-2]} electrons, nspins={self.nspins}")

    def evaluate(self, params, data: MoleculeData) -> dict:
        """Returns ``{"logpsi", "sign_logpsi"}`` per walker (reference output/logdet.py:24-37)."""
        self._check(data)
        el, squeeze = _batched(data.electrons)
        rt = runtime(el.device)
        wf = self._handle(params, data.atoms.shape[0])
        sysh = _marshal.system_handle(data.atoms, None)
        logpsi, sign = rt.logpsi(wf, sysh, el)
        if squeeze:
            logpsi, sign = logpsi[0], sign[0]
        return {"logpsi": logpsi, "sign_logpsi": sign}

    def logpsi(self, params, data: MoleculeData) -> torch.Tensor:
        return self.evaluate(params, data)["logpsi"]

    def phase_logpsi(self, params, data: MoleculeData):
        out = self.evaluate(params, data)
        return out["sign_logpsi"], out["logpsi"]

    def __call__(self, params, data: MoleculeData):
        return self.evaluate(params, data)

    def local_energy(self, params, data: MoleculeData, sums: torch.Tensor | None = None) -> dict:
        """Value, gradient, Laplacian, kinetic / potential / local energy per walker in one pass."""
        self._check(data)
        el, _ = _batched(data.electrons)
        rt = runtime(el.device)
        wf = self._handle(params, data.atoms.shape[0])
        sysh = _marshal.system_handle(data.atoms, data.charges)
        return rt.local_energy(wf, sysh, el, sums=sums)


@dataclass
class FermiNetWavefunction(Wavefunction):
    """FermiNet ansatz (reference app/molecule/wavefunction/ferminet.py:21-74); same fields and defaults."""

    nspins: tuple = (1, 1)
    ndets: int = 16
    hidden_dims_single: list = field(default_factory=lambda: [256] * 4)
    hidden_dims_double: list = field(default_factory=lambda: [32] * 4)
    use_last_layer: bool = False
    envelope: str = "abs_isotropic"
    orbitals_spin_split: bool = True
    full_det: bool = True

    def __post_init__(self):
        if not self.full_det:
            raise ValueError("FermiNet requires full_det=True.")
        if self.use_last_layer:
            raise NotImplementedError("use_last_layer=True is not supported by the CUDA pipeline")
        if len(self.hidden_dims_single) != len(self.hidden_dims_double):
            raise ValueError("hidden_dims_single and hidden_dims_double must have the same length")

    def init_params(self, data: MoleculeData, rngs) -> dict:
        """Flax-layout tree: LeCun-normal kernels, zero biases, ``pi = sigma = 1``
        (flax ``nn.Dense`` defaults; output/envelope.py:134-135).  ``rngs`` is an int seed or a ``torch.Generator``."""
        dev = data.electrons.device
        g = rngs if isinstance(rngs, torch.Generator) else torch.Generator(device="cpu").manual_seed(int(rngs))
        n_up, n_dn = self.nspins
        n = n_up + n_dn
        A = data.atoms.shape[0]
        nch = 2 if (n_up > 0 and n_dn > 0) else 1

        def lecun(*shape):
            fan_in = shape[0]
            t = torch.randn(*shape, generator=g, dtype=torch.float32) / math.sqrt(fan_in)
            return t.to(dev)

        bb = {}
        d1, d2 = 4 * A, 4
        idx = 0
        L = len(self.hidden_dims_single)
        for layer in range(L):
            h1 = self.hidden_dims_single[layer]
            bb[f"Dense_{idx}"] = {"kernel": lecun(d1 * (1 + nch) + d2 * nch, h1), "bias": torch.zeros(h1, device=dev)}
            idx += 1
            if layer < L - 1:
                h2 = self.hidden_dims_double[layer]
                bb[f"Dense_{idx}"] = {"kernel": lecun(d2, h2), "bias": torch.zeros(h2, device=dev)}
                idx += 1
                d2 = h2
            d1 = h1
        split = self.orbitals_spin_split and nch == 2
        if split:
            orb = {"SplitChannelDense_0": {f"DenseGeneral_{s}": {"kernel": lecun(d1, self.ndets, n)} for s in range(2)}}
        else:
            orb = {"DenseGeneral_0": {"kernel": lecun(d1, self.ndets, n)}}
        ones = lambda: torch.ones(n, A, self.ndets, device=dev)  # noqa: E731
        names = ["_env_up", "_env_down"] if split else ["_env"]
        env = {nm: {"pi": ones(), "sigma": ones()} for nm in names} if self.envelope != "null" else {}
        return {"params": {"backbone_layer": bb, "orbital_layer": orb, "envelope_layer": env}}

    def _handle(self, params, n_atoms: int):
        # rebuilt per call (a few dozen pointer reads): parameter leaves may have been replaced by the optimizer
        return _marshal.ferminet_handle(params, self.nspins, n_atoms, self.ndets, self.hidden_dims_single,
                                        self.hidden_dims_double, self.envelope, self.orbitals_spin_split)
