"""Energy estimators: host-side mirror of ``estimator/base.py`` (``PerWalkerEstimator``, ``mean_reduce``,
``EstimatorPipeline``), ``estimator/kinetic/euclidean.py`` (``EuclideanKinetic``, forward-Laplacian mode),
``app/molecule/hamiltonian.py`` (``potential_energy``), ``estimator/total_energy.py`` (``TotalEnergy``) and
``estimator/loss_grad.py`` (``LossAndGrad``).

Batched like the wavefunction classes: ``evaluate_batch_walkers`` is the unit of work (the reference reaches it by
``chunked_vmap`` of ``evaluate_single_walker``, estimator/base.py:251-270).
"""

from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _marshal
from ._runtime import runtime
from .data import MoleculeData


def mean_reduce(walker_stats: dict, include_variance: bool = True) -> dict:
    """``pmean(nanmean(x))`` and variance ``E[x^2] - E[x]^2`` (estimator/base.py:27-53).  The cross-device mean is a
    NCCL all-reduce of scalars when ``torch.distributed`` is initialised."""
    dist = torch.distributed.is_available() and torch.distributed.is_initialized()
    out = {}
    for k, v in walker_stats.items():
        m = torch.nanmean(v, dim=0)
        m2 = torch.nanmean(v * v, dim=0) if include_variance else None
        if dist:
            ws = torch.distributed.get_world_size()
            buf = torch.stack([m, m2]) if include_variance else m.reshape(1)
            torch.distributed.all_reduce(buf)
            buf = buf / ws
            m = buf[0]
            m2 = buf[1] if include_variance else None
        out[k] = m
        if include_variance:
            out[f"{k}_var"] = m2 - m * m
    return out


def potential_energy(params, data: MoleculeData, prev_walker_stats=None, state=None, rngs=None):
    """Coulomb potential per walker (app/molecule/hamiltonian.py:9-22) -> ``({"energy:potential": (W,)}, state)``."""
    el = data.electrons.contiguous()
    rt = runtime(el.device)
    sysh = _marshal.system_handle(data.atoms, data.charges)
    return {"energy:potential": rt.coulomb(sysh, el)}, state


@dataclass
class EuclideanKinetic:
    """``E_kin = -1/2 (lap log|psi| + |grad log|psi||^2)`` by the forward Laplacian
    (estimator/kinetic/euclidean.py:114-135).  ``f_log_psi`` is the wavefunction object."""

    f_log_psi: object = None
    mode: str = "forward_laplacian"
    data_field: str = "electrons"
    sparse: bool = True
    vmap_chunk_size: int | None = None

    MODES = ("forward_laplacian", "scan", "fori_loop")

    def __post_init__(self):
        # LaplacianMode (estimator/kinetic/_common.py): ``scan`` / ``fori_loop`` are the reference's older ways of
        # obtaining the SAME Laplacian (a loop of Hessian-vector products, euclidean.py:84-112) and ``sparse`` /
        # ``vmap_chunk_size`` only steer XLA's memory use.  Every mode maps to the one fused forward-Laplacian kernel
        # pipeline: the result is the same quantity (the reference's own tests hold the modes to 2e-5 of each other,
        # tests/estimator/kinetic_forward_laplacian_test.py:254-524), so the option surface is accepted, not emulated.
        if str(self.mode) not in self.MODES:
            raise ValueError(f"Unknown Laplacian mode: {self.mode!r}. Must be one of: {list(self.MODES)}")

    def evaluate_batch_walkers(self, params, data: MoleculeData, prev_walker_stats=None, state=None, rngs=None):
        out = self.f_log_psi.local_energy(params, data)
        stats = {"energy:kinetic": out["e_kin"]}
        if prev_walker_stats is not None:
            prev_walker_stats.setdefault("_fused", out)
        return stats, state


@dataclass
class TotalEnergy:
    """Sum of every ``energy:*`` key per walker (estimator/total_energy.py:36-60)."""

    def evaluate_batch_walkers(self, params, data, prev_walker_stats, state=None, rngs=None):
        keys = [k for k in prev_walker_stats if k.startswith("energy:")]
        if not keys:
            raise ValueError("TotalEnergy needs at least one 'energy:*' key from earlier estimators")
        total = None
        for k in keys:
            v = prev_walker_stats[k]
            if v.dim() != 1:
                raise ValueError(f"Energy term {k!r} must be a scalar per walker, got shape {tuple(v.shape[1:])}")
            total = v if total is None else total + v
        return {"total_energy": total}, state


def clip_observable(x: torch.Tensor, method: str, scale: float = 100.0) -> torch.Tensor:
    """Outlier clipping of the local energies (reference utils/clip.py): the window statistics are taken over ALL walkers
    -- the reference all-gathers them across devices first."""
    if method == "none":
        return x
    allx = x
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        parts = [torch.empty_like(x) for _ in range(torch.distributed.get_world_size())]
        torch.distributed.all_gather(parts, x.contiguous())
        allx = torch.cat(parts)
    if method == "iqr":
        q1, q3 = torch.nanquantile(allx, 0.25), torch.nanquantile(allx, 0.75)
        return torch.minimum(torch.maximum(x, q1 - scale * (q3 - q1)), q3 + scale * (q3 - q1))
    if method == "mad":
        med = torch.nanquantile(allx, 0.5)
        dev = torch.nanquantile((allx - med).abs(), 0.5)
        return torch.minimum(torch.maximum(x, med - scale * dev), med + scale * dev)
    raise ValueError(f"Unknown clip method {method!r}.")


@dataclass
class LossAndGrad:
    """VMC loss and parameter gradients (reference estimator/loss_grad.py:25-128; same fields and defaults):
    ``grads = 2 (<d log psi * E_clip> - <E_clip> <d log psi>)``.

    The reference vmaps ``jax.value_and_grad(f_log_psi)`` (a W x P score tensor), multiplies by the clipped local energies
    and averages.  Both averages are one reverse pass here: ``wf.logpsi_vjp`` with the per-walker cotangent
    ``2 (E_clip,w - <E_clip>) / W`` (W = global walker count), followed by the all-reduce of the gradient over devices
    that ``mean_reduce`` performs in the reference.  ``f_log_psi`` is the wavefunction object."""

    f_log_psi: object = None
    loss_key: str = "total_energy"
    clip_method: str = "mad"
    clip_scale: float = 5.0

    def evaluate(self, params, data: MoleculeData, walker_stats: dict) -> dict:
        loss = walker_stats[self.loss_key]
        if loss.dim() != 1:
            raise ValueError(f"Expected scalar loss value (i.e. ndim=0), got shape {tuple(loss.shape[1:])}.")
        dist = torch.distributed.is_available() and torch.distributed.is_initialized()
        clipped = clip_observable(loss, self.clip_method, self.clip_scale)
        ok = torch.isfinite(clipped)
        stats = torch.stack([torch.where(ok, clipped, torch.zeros_like(clipped)).sum(), ok.sum().to(clipped.dtype),
                             torch.nansum(loss), torch.isfinite(loss).sum().to(loss.dtype)])
        if dist:
            torch.distributed.all_reduce(stats)
        mean_c, count = stats[0] / stats[1], stats[1]
        cot = torch.where(ok, 2.0 * (clipped - mean_c) / count, torch.zeros_like(clipped))
        grads, logpsi = self.f_log_psi.logpsi_vjp(params, data, cot)
        if dist:
            for leaf in _leaves(grads):
                torch.distributed.all_reduce(leaf)
        return {"loss": stats[2] / stats[3], "grads": grads, "clipped_loss": mean_c, "logpsi": logpsi}


def _leaves(tree):
    if isinstance(tree, dict):
        for k in sorted(tree):
            yield from _leaves(tree[k])
    else:
        yield tree


class EstimatorPipeline:
    """Runs estimators in insertion order, threading the accumulated per-walker stats
    (estimator/base.py:331-362), then reduces them with :func:`mean_reduce`."""

    def __init__(self, estimators: dict):
        self.estimators = dict(estimators)

    def evaluate(self, params, data: MoleculeData, state=None, rngs=None):
        walker_stats: dict = {}
        for name, est in self.estimators.items():
            fn = est.evaluate_batch_walkers if hasattr(est, "evaluate_batch_walkers") else est
            stats, _ = fn(params, data, walker_stats, state, rngs)
            walker_stats.update(stats)
        walker_stats.pop("_fused", None)
        return mean_reduce(walker_stats), walker_stats
