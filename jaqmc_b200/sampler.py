"""Metropolis-Hastings sampler: host-side mirror of ``sampler/mcmc.py`` (``MCMCSampler``, ``MCMCState``) and of
the ``SamplePlan`` closure ``x -> 2 log|psi(x)|`` (``sampler/base.py:136-138,173-178``).

The reference's sampler never sees the wavefunction, only a ``batch_log_prob`` callable.  Here that callable is
a :class:`BatchLogProb` object: calling it returns ``2 log|psi|`` like the reference's closure, and it also carries
the wavefunction descriptor so that ``MCMCSampler.step`` can run all sub-steps inside the library
(``jaqmc_b200_mh_step`` / ``jaqmc_b200_mh_step_pbc``: forward pass + fused accept/select/next-proposal kernel per
sub-step).  Any other callable or proposal is driven through the stand-alone propose / accept kernels.
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
    """``x -> 2 log|psi(params, data.merge(electrons=x))|`` for a walker batch (sampler/base.py:173-178).
    A complex log psi (periodic network) contributes its real part, as in ``_mh_update`` (sampler/mcmc.py:122,167)."""

    def __init__(self, wf, params, data):
        self.wf, self.params, self.data = wf, params, data

    def __call__(self, electrons: torch.Tensor) -> torch.Tensor:
        lp = self.wf.logpsi(self.params, self.data.merge({"electrons": electrons}))
        return 2.0 * (lp.real if lp.is_complex() else lp)

    def handles(self):
        """``(wavefunction descriptor, system descriptor)`` of the fused in-library MH step."""
        return self.wf._sampling_handles(self.params, self.data)


class GaussianProposal:
    """``x + normal * stddev`` (sampler/mcmc.py:53-54 ``gaussian_proposal``)."""

    lattice = None

    def __call__(self, normals: torch.Tensor, x: torch.Tensor, stddev: torch.Tensor) -> torch.Tensor:
        return x + normals * stddev


gaussian_proposal = GaussianProposal()


class PbcGaussianProposal(GaussianProposal):
    """Gaussian move wrapped into the periodic cell (geometry/pbc.py:187-201 ``make_pbc_gaussian_proposal``:
    ``wrap_positions(x + normal * stddev, lattice)``, ``wrap_positions`` = fractional coordinates modulo 1, :97-111)."""

    def __init__(self, lattice):
        self.lattice = torch.as_tensor(lattice, dtype=torch.float32).reshape(3, 3).cpu()

    def __call__(self, normals, x, stddev):
        lat = self.lattice.to(x.device)
        frac = (x + normals * stddev) @ torch.linalg.inv(lat)
        return (frac - torch.floor(frac)) @ lat


def make_pbc_gaussian_proposal(lattice) -> PbcGaussianProposal:
    return PbcGaussianProposal(lattice)


def _leaf_pointers(tree):
    """Device addresses of every tensor leaf of a parameter tree, in sorted-key order."""
    if isinstance(tree, torch.Tensor):
        return (tree.data_ptr(), tuple(tree.shape), tree.is_contiguous())
    if isinstance(tree, dict):
        return tuple((k, _leaf_pointers(tree[k])) for k in sorted(tree))
    if isinstance(tree, (list, tuple)):
        return tuple(_leaf_pointers(t) for t in tree)
    return tree


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
    sampling_proposal: object = gaussian_proposal   # runtime dep of the reference (sampler/mcmc.py:76)

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
        prop = self.sampling_proposal
        fused = isinstance(batch_log_prob, BatchLogProb) and isinstance(prop, GaussianProposal)
        if fused and graph_cache is not None and not record_accepts and dev.type == "cuda":
            blp = batch_log_prob
            wf, sysh = blp.handles()
            # the captured graph bakes in the device addresses of every parameter leaf and of the system arrays: a
            # replaced leaf (optimizer step, ``merge``) or a non-contiguous one (copied by the marshaller) recaptures
            key = (id(blp.wf), tuple(x.shape), self.steps, _leaf_pointers(blp.params),
                   tuple(t.data_ptr() for t in sysh._keep if isinstance(t, torch.Tensor)),
                   None if prop.lattice is None else tuple(prop.lattice.reshape(-1).tolist()))
            if any(isinstance(t, tuple) and t[2] is False for t in _flatten_ptrs(key[3])):
                raise ValueError("SamplePlan(graph=True) needs contiguous parameter leaves (the graph reads them in place)")
            ent = graph_cache.get(key)
            if ent is None:
                graph_cache.clear()   # one live graph per plan: drop graphs that captured stale parameter addresses
                bufs = dict(x=x.clone(), normals=normals.clone(), uniforms=uniforms.clone(), stddev=state.stddev.clone())
                replay = rt.capture_mh_step(wf, sysh, bufs["x"], bufs["normals"], bufs["uniforms"], bufs["stddev"],
                                            wrap_lattice=prop.lattice)
                ent = graph_cache[key] = (replay, bufs, blp.params)
            replay, bufs, _ = ent
            bufs["x"].copy_(x)
            bufs["normals"].copy_(normals)
            bufs["uniforms"].copy_(uniforms)
            bufs["stddev"].copy_(state.stddev)
            n_acc = replay().clone()
            x = bufs["x"].clone()
            accepted = None
        elif fused:
            wf, sysh = batch_log_prob.handles()
            logpsi = torch.empty(W, dtype=torch.float32, device=dev)
            n_acc, accepted = rt.mh_step(wf, sysh, x, logpsi, normals, uniforms, state.stddev, logpsi_valid=False,
                                         record_accepts=record_accepts, wrap_lattice=prop.lattice)
        else:
            # any other log-probability callable / proposal: drive the loop from the host with the stand-alone kernels
            lp = batch_log_prob(x)
            lp = (lp.real if lp.is_complex() else lp).contiguous()
            if lp.dim() != 1:
                raise ValueError(f"log_amplitude should return a scalar, got shape {tuple(lp.shape[1:])}.")
            n_acc = torch.zeros(1, dtype=torch.float32, device=dev)
            accepted = torch.empty(self.steps, W, dtype=torch.uint8, device=dev) if record_accepts else None
            for s in range(self.steps):
                if type(prop) is GaussianProposal:
                    x2 = rt.mh_propose(x, normals[s], state.stddev)
                else:
                    x2 = prop(normals[s], x, state.stddev).contiguous()
                lp2 = batch_log_prob(x2)
                lp2 = (lp2.real if lp2.is_complex() else lp2).contiguous()
                rt.mh_accept(x, x2, lp, lp2, uniforms[s], n_acc, accepted[s] if record_accepts else None)
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


def _flatten_ptrs(t):
    if isinstance(t, tuple) and len(t) == 3 and isinstance(t[0], int) and isinstance(t[1], tuple):
        yield t
    elif isinstance(t, tuple):
        for u in t:
            yield from _flatten_ptrs(u)


class SamplePlan:
    """``SamplePlan`` for the single batched field ``electrons`` (sampler/base.py:123-181)."""

    def __init__(self, wf, sampler: MCMCSampler, graph: bool = False):
        self.wf, self.sampler = wf, sampler
        self._graphs = {} if graph else None   # CUDA-graph replay of the MH sub-steps (parameters must stay in place)

    def init(self, data: MoleculeData, rngs=None) -> MCMCState:
        return self.sampler.init(data, rngs)

    def step(self, params, data: MoleculeData, state: MCMCState, rngs):
        return self.sampler.step(BatchLogProb(self.wf, params, data), data, state, rngs, graph_cache=self._graphs)
