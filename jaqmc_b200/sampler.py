"""Metropolis-Hastings sampler: host-side mirror of ``sampler/mcmc.py`` (``MCMCSampler``, ``MCMCState``) and of
the ``SamplePlan`` closure ``x -> 2 log|psi(x)|`` (``sampler/base.py:136-138,173-178``).

The reference's sampler never sees the wavefunction, only a ``batch_log_prob`` callable.  Here that callable is
a :class:`BatchLogProb` object: calling it returns ``2 log|psi|`` like the reference's closure, and it also carries
the wavefunction descriptor so that ``MCMCSampler.step`` can run all sub-steps inside the library
(``jaqmc_b200_mh_step``: forward pass + fused accept/select/next-proposal kernel per sub-step).  Any other callable
is driven through the stand-alone propose / accept kernels.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import NamedTuple

import torch

from . import _marshal
from ._runtime import runtime
from .data import MoleculeData


class MCMCState(NamedTuple):
    """``stddev`` (1,), ``pmoves`` (adapt_frequency,), ``counter`` -- reference sampler/mcmc.py:39-50."""

    stddev: torch.Tensor
    pmoves: torch.Tensor
    counter: int


class BatchLogProb:
    """``x -> 2 log|psi(params, data.merge(electrons=x))|`` for a walker batch (sampler/base.py:173-178)."""

    def __init__(self, wf, params, data: MoleculeData):
        self.wf, self.params, self.data = wf, params, data

    def __call__(self, electrons: torch.Tensor) -> torch.Tensor:
        return 2.0 * self.wf.logpsi(self.params, self.data.merge({"electrons": electrons}))


def _noise(rngs, steps, shape, device):
    """``rngs``: a ``torch.Generator`` on ``device`` or an explicit ``(normals, uniforms)`` pair."""
    if isinstance(rngs, (tuple, list)):
        normals, uniforms = rngs
        return normals.contiguous(), uniforms.contiguous()
    normals = torch.randn(steps, *shape, generator=rngs, device=device, dtype=torch.float32)
    uniforms = torch.rand(steps, shape[0], generator=rngs, device=device, dtype=torch.float32)
    return normals, uniforms.clamp_min_(torch.finfo(torch.float32).tiny)


@dataclass
class MCMCSampler:
    """Same fields and defaults as the reference (sampler/mcmc.py:72-75)."""

    steps: int = 10
    initial_width: float = 0.1
    adapt_frequency: int = 100
    pmove_range: tuple = (0.5, 0.55)

    def init(self, data, rngs=None) -> MCMCState:
        dev = data.electrons.device if hasattr(data, "electrons") else data.device
        return MCMCState(
            stddev=torch.full((1,), float(self.initial_width), dtype=torch.float32, device=dev),
            pmoves=torch.zeros(self.adapt_frequency, dtype=torch.float32, device=dev),
            counter=0,
        )

    def step(self, batch_log_prob, data: MoleculeData, state: MCMCState, rngs, record_accepts: bool = False,
             graph_cache: dict | None = None):
        """``(data, {"pmove": ...}, new_state)`` after ``steps`` all-electron MH updates (sampler/mcmc.py:139-197).
        ``graph_cache`` (a dict owned by the caller, see ``SamplePlan(graph=True)``) replays the sub-steps as one CUDA
        graph on static buffers; results are identical to the launch-by-launch path."""
        if self.steps == 0:
            return data, {"pmove": torch.zeros((), device=data.electrons.device)}, state
        x = data.electrons.contiguous().clone()
        W = x.shape[0]
        dev = x.device
        normals, uniforms = _noise(rngs, self.steps, tuple(x.shape), dev)
        rt = runtime(dev)
        if graph_cache is not None and isinstance(batch_log_prob, BatchLogProb) and not record_accepts and dev.type == "cuda":
            blp = batch_log_prob
            key = (id(blp.wf), tuple(x.shape), self.steps)
            ent = graph_cache.get(key)
            if ent is None:
                wf = blp.wf._handle(blp.params, blp.data.atoms.shape[0])
                sysh = _marshal.system_handle(blp.data.atoms, None)
                bufs = dict(x=x.clone(), normals=normals.clone(), uniforms=uniforms.clone(), stddev=state.stddev.clone())
                replay = rt.capture_mh_step(wf, sysh, bufs["x"], bufs["normals"], bufs["uniforms"], bufs["stddev"])
                ent = graph_cache[key] = (replay, bufs, blp.params)
            replay, bufs, _ = ent
            bufs["x"].copy_(x)
            bufs["normals"].copy_(normals)
            bufs["uniforms"].copy_(uniforms)
            bufs["stddev"].copy_(state.stddev)
            n_acc = replay().clone()
            x = bufs["x"].clone()
            accepted = None
        elif isinstance(batch_log_prob, BatchLogProb):
            blp = batch_log_prob
            wf = blp.wf._handle(blp.params, blp.data.atoms.shape[0])
            sysh = _marshal.system_handle(blp.data.atoms, None)
            logpsi = torch.empty(W, dtype=torch.float32, device=dev)
            n_acc, accepted = rt.mh_step(wf, sysh, x, logpsi, normals, uniforms, state.stddev, logpsi_valid=False,
                                         record_accepts=record_accepts)
        else:
            import ctypes as C

            from . import _abi
            lp = batch_log_prob(x).contiguous()
            if lp.dim() != 1:
                raise ValueError(f"log_amplitude should return a scalar, got shape {tuple(lp.shape[1:])}.")
            n_acc = torch.zeros(1, dtype=torch.float32, device=dev)
            accepted = torch.empty(self.steps, W, dtype=torch.uint8, device=dev) if record_accepts else None
            x2 = torch.empty_like(x)
            p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
            for s in range(self.steps):
                _abi.check(rt.lib, rt.lib.jaqmc_b200_mh_propose(p(x), p(normals[s]), p(state.stddev), p(x2), x.numel(),
                                                                  rt._stream()))
                lp2 = batch_log_prob(x2).contiguous()
                acc_p = C.c_void_p(accepted[s].data_ptr()) if record_accepts else C.c_void_p(0)
                _abi.check(rt.lib, rt.lib.jaqmc_b200_mh_accept(p(x), p(x2), p(lp), p(lp2), p(uniforms[s]), W,
                                                                 x.shape[1] * 3, p(n_acc), acc_p, rt._stream()))
        pmove = (n_acc / float(self.steps * W)).reshape(())
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(pmove)  # pmean over devices (sampler/mcmc.py:180)
            pmove = pmove / torch.distributed.get_world_size()
        # adaptive width (sampler/mcmc.py:182-195), kept on the device: no host synchronisation
        counter = state.counter + 1
        t = counter % self.adapt_frequency
        pmoves = state.pmoves.clone()
        pmoves[t] = pmove
        stddev = state.stddev
        if t == 0:
            m = pmoves.mean()
            stddev = torch.where(m > self.pmove_range[1], stddev * 1.1,
                                 torch.where(m < self.pmove_range[0], stddev / 1.1, stddev))
        new_state = MCMCState(stddev=stddev, pmoves=pmoves, counter=counter)
        stats = {"pmove": pmove}
        if record_accepts:
            stats["accepted"] = accepted
        return data.merge({"electrons": x}), stats, new_state


class SamplePlan:
    """``SamplePlan`` for the single batched field ``electrons`` (sampler/base.py:123-181)."""

    def __init__(self, wf, sampler: MCMCSampler, graph: bool = False):
        self.wf, self.sampler = wf, sampler
        self._graphs = {} if graph else None   # CUDA-graph replay of the MH sub-steps (parameters must stay in place)

    def init(self, data: MoleculeData, rngs=None) -> MCMCState:
        return self.sampler.init(data, rngs)

    def step(self, params, data: MoleculeData, state: MCMCState, rngs):
        return self.sampler.step(BatchLogProb(self.wf, params, data), data, state, rngs, graph_cache=self._graphs)
