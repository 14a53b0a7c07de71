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


def _tree_zeros_like(tree):
    if isinstance(tree, dict):
        return {k: _tree_zeros_like(v) for k, v in tree.items()}
    return torch.zeros_like(tree, memory_format=torch.contiguous_format)


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

    def _sampling_handles(self, params, data):
        """Descriptors of the in-library MH step (``sampler.BatchLogProb.handles``)."""
        return self._handle(params, data.atoms.shape[0]), _marshal.system_handle(data.atoms, None)

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

    def orbitals(self, params, data: MoleculeData) -> torch.Tensor:
        """Orbital matrices ``(W, ndets, n, n)`` [or ``(ndets, n, n)`` for one walker] after the envelope -- the
        pretraining head (reference app/molecule/wavefunction/base.py:60-72, ferminet.py:126-138)."""
        self._check(data)
        el, squeeze = _batched(data.electrons)
        rt = runtime(el.device)
        wf, sysh = self._sampling_handles(params, data)
        orb = rt.orbitals(wf, sysh, el, self.ndets)
        return orb[0] if squeeze else orb

    def local_energy(self, params, data: MoleculeData, sums: torch.Tensor | None = None) -> dict:
        """Value, gradient, Laplacian, kinetic / potential / local energy per walker in one pass."""
        self._check(data)
        el, _ = _batched(data.electrons)
        rt = runtime(el.device)
        wf = self._handle(params, data.atoms.shape[0])
        sysh = _marshal.system_handle(data.atoms, data.charges)
        return rt.local_energy(wf, sysh, el, sums=sums)


def capture_local_energy(wf: "Wavefunction", params, data: MoleculeData, sums: torch.Tensor | None = None):
    """``(replay, out)`` -- CUDA-graph replay of ``wf.local_energy(params, data)`` for fixed shapes; ``data.electrons``
    and the parameter leaves are read in place at every replay."""
    wf._check(data)
    el, _ = _batched(data.electrons)
    if el.data_ptr() != data.electrons.data_ptr():
        raise ValueError("capture_local_energy needs a contiguous (W, n, 3) electrons tensor")
    rt = runtime(el.device)
    handle = wf._handle(params, data.atoms.shape[0])
    sysh = _marshal.system_handle(data.atoms, data.charges)
    return rt.capture_local_energy(handle, sysh, el, sums=sums)


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
            if layer < L - 1 or self.use_last_layer:
                h2 = self.hidden_dims_double[layer]
                bb[f"Dense_{idx}"] = {"kernel": lecun(d2, h2), "bias": torch.zeros(h2, device=dev)}
                idx += 1
                d2 = h2
            d1 = h1
        if self.use_last_layer:
            d1 = d1 * (1 + nch) + d2 * nch
        split = self.orbitals_spin_split and nch == 2
        if split:
            orb = {"SplitChannelDense_0": {f"DenseGeneral_{s}": {"kernel": lecun(d1, self.ndets, n)} for s in range(2)}}
        else:
            orb = {"DenseGeneral_0": {"kernel": lecun(d1, self.ndets, n)}}
        names = ["_env_up", "_env_down"] if split else ["_env"]
        env = _envelope_params(dev, names, n, A, self.ndets, self.envelope)
        return {"params": {"backbone_layer": bb, "orbital_layer": orb, "envelope_layer": env}}

    def logpsi_vjp(self, params, data: MoleculeData, cotangent: torch.Tensor):
        """``(grads, logpsi)``: ``grads`` has the tree structure of ``params`` and holds
        ``sum_w cotangent[w] * d log|psi|(x_w) / d theta`` -- the VJP of the batched ``logpsi`` with respect to the
        parameters (what ``LossAndGrad`` needs, reference estimator/loss_grad.py:70-128, without the per-walker score
        tensor)."""
        self._check(data)
        el, _ = _batched(data.electrons)
        rt = runtime(el.device)
        A = data.atoms.shape[0]
        wf = self._handle(params, A)
        grads = _tree_zeros_like(params)
        gh = self._handle(grads, A)
        sysh = _marshal.system_handle(data.atoms, None)
        logpsi, _ = rt.ferminet_logpsi_vjp(wf, gh, sysh, el, cotangent.contiguous())
        return grads, logpsi

    def _handle(self, params, n_atoms: int):
        # rebuilt per call (a few dozen pointer reads): parameter leaves may have been replaced by the optimizer
        return _marshal.ferminet_handle(params, self.nspins, n_atoms, self.ndets, self.hidden_dims_single,
                                        self.hidden_dims_double, self.envelope, self.orbitals_spin_split,
                                        self.use_last_layer)


def _envelope_params(dev, names, n, n_atoms, ndets, envelope):
    """``pi = sigma = 1`` (output/envelope.py:131-135; diagonal: sigma (n_orb, A, 3, D), :150-156)."""
    if envelope == "null":
        return {}
    sshape = (n, n_atoms, 3, ndets) if envelope == "diagonal" else (n, n_atoms, ndets)
    return {nm: {"pi": torch.ones(n, n_atoms, ndets, device=dev), "sigma": torch.ones(*sshape, device=dev)}
            for nm in names}


def _lecun(g, dev, *shape, fan_in=None):
    fan_in = shape[0] if fan_in is None else fan_in
    return (torch.randn(*shape, generator=g, dtype=torch.float32) / math.sqrt(fan_in)).to(dev)


def _head_params(g, dev, nspins, n_atoms, ndets, hidden, split, envelope, bias_orbitals, jastrow, alpha_init):
    """orbital_layer / envelope_layer / jastrow_layer sub-trees with the reference's initialisers
    (output/orbital.py:59-78, output/envelope.py:131-135, jastrow.py:61-63)."""
    n = sum(nspins)

    def dg():
        d = {"kernel": _lecun(g, dev, hidden, ndets, n)}
        if bias_orbitals:
            d["bias"] = torch.zeros(ndets, n, device=dev)
        return d

    orb = {"SplitChannelDense_0": {"DenseGeneral_0": dg(), "DenseGeneral_1": dg()}} if split else {"DenseGeneral_0": dg()}
    names = ["_env_up", "_env_down"] if split else ["_env"]
    env = _envelope_params(dev, names, n, n_atoms, ndets, envelope)
    out = {"orbital_layer": orb, "envelope_layer": env}
    if jastrow:
        out["jastrow_layer"] = {"alpha_par": torch.full((1,), float(alpha_init), device=dev),
                                "alpha_anti": torch.full((1,), float(alpha_init), device=dev)}
    return out


@dataclass
class LapNetWavefunction(Wavefunction):
    """LapNet ansatz (reference app/molecule/wavefunction/lapnet.py:39-135); same fields and defaults."""

    nspins: tuple = (1, 1)
    ndets: int = 16
    num_layers: int = 4
    num_heads: int = 4
    heads_dim: int = 64
    use_layernorm: bool = False
    jastrow: str = "simple_ee"
    use_input_bias: bool = True
    use_backbone_bias: bool = True
    num_local_updates: int = 2
    envelope: str = "abs_isotropic"
    use_orbital_bias: bool = False
    rescale: bool = True
    jastrow_alpha_init: float = 1.0
    full_det: bool = True

    def __post_init__(self):
        if not self.full_det:
            raise ValueError("LapNet requires full_det=True.")
        if self.num_layers <= 0:
            raise ValueError("LapNet requires at least one layer.")
        if self.jastrow not in ("none", "simple_ee"):
            raise ValueError(f"Invalid jastrow: {self.jastrow!r}. Must be one of: ['none', 'simple_ee']")

    def init_params(self, data: MoleculeData, rngs) -> dict:
        """Flax-layout tree (backbone/lapnet/_backbone.py:40-64,169-189): LeCun-normal kernels, N(0,1) biases."""
        dev = data.electrons.device
        g = rngs if isinstance(rngs, torch.Generator) else torch.Generator(device="cpu").manual_seed(int(rngs))
        A = data.atoms.shape[0]
        hid = self.num_heads * self.heads_dim

        def dense(fi, fo, use_bias):
            d = {"kernel": _lecun(g, dev, fi, fo)}
            if use_bias:
                d["bias"] = torch.randn(fo, generator=g, dtype=torch.float32).to(dev)
            return d

        bb = {"input_projection": dense(4 * A + 1, hid, self.use_input_bias)}
        for l in range(self.num_layers):
            lp = {"qk_projection": dense(hid, 2 * hid, self.use_backbone_bias),
                  "value_projection": dense(hid, hid, self.use_backbone_bias),
                  "output_projection": dense(hid, hid, self.use_backbone_bias),
                  "value_update": dense(hid, hid, self.use_backbone_bias)}
            if l < self.num_layers - 1:
                for j in range(self.num_local_updates):
                    lp[f"qk_update_layers_{j}"] = dense(hid, hid, self.use_backbone_bias)
            if self.use_layernorm:
                for nm in ("qk_layernorm", "value_layernorm", "post_attention_layernorm"):
                    lp[nm] = {"scale": torch.ones(hid, device=dev), "bias": torch.zeros(hid, device=dev)}
            bb[f"layers_{l}"] = lp
        split = self.nspins[0] > 0 and self.nspins[1] > 0
        tree = {"backbone_layer": bb}
        tree.update(_head_params(g, dev, self.nspins, A, self.ndets, hid, split, self.envelope, self.use_orbital_bias,
                                 self.jastrow == "simple_ee", self.jastrow_alpha_init))
        return {"params": tree}

    def _handle(self, params, n_atoms: int):
        return _marshal.lapnet_handle(params, self.nspins, n_atoms, self.ndets, self.num_layers, self.num_heads,
                                      self.heads_dim, self.num_local_updates, self.envelope, self.rescale,
                                      self.jastrow == "simple_ee", self.use_layernorm)


@dataclass
class PsiformerWavefunction(Wavefunction):
    """Psiformer ansatz (reference app/molecule/wavefunction/psiformer.py:39-167); same fields and defaults."""

    nspins: tuple = (1, 1)
    ndets: int = 16
    num_layers: int = 4
    num_heads: int = 4
    heads_dim: int = 64
    mlp_hidden_dims: list = field(default_factory=lambda: [256])
    layer_norm_mode: str = "pre"
    jastrow: str = "simple_ee"
    with_bias: bool = True
    input_bias: bool = True
    envelope: str = "abs_isotropic"
    orbitals_spin_split: bool = True
    bias_orbitals: bool = False
    rescale: bool = True
    jastrow_alpha_init: float = 1.0
    full_det: bool = True

    def __post_init__(self):
        if not self.full_det:
            raise ValueError("Psiformer requires full_det=True.")
        if self.jastrow not in ("none", "simple_ee"):
            raise ValueError(f"Invalid jastrow: {self.jastrow!r}. Must be one of: ['none', 'simple_ee']")
        if self.layer_norm_mode not in ("pre", "post", "null"):
            raise ValueError(f"Invalid layer_norm_mode: {self.layer_norm_mode!r}")

    def init_params(self, data: MoleculeData, rngs) -> dict:
        """Flax-layout tree (backbone/psiformer.py:60-99,175-185): LeCun-normal kernels, zero biases, LayerNorm
        scale 1 / bias 0."""
        dev = data.electrons.device
        g = rngs if isinstance(rngs, torch.Generator) else torch.Generator(device="cpu").manual_seed(int(rngs))
        A = data.atoms.shape[0]
        H, dh = self.num_heads, self.heads_dim
        hid = H * dh

        def dense(fi, fo, use_bias=True):
            d = {"kernel": _lecun(g, dev, fi, fo)}
            if use_bias:
                d["bias"] = torch.zeros(fo, device=dev)
            return d

        def proj():
            d = {"kernel": _lecun(g, dev, hid, H, dh)}
            if self.with_bias:
                d["bias"] = torch.zeros(H, dh, device=dev)
            return d

        bb = {"Dense_0": dense(4 * A + 1, hid, self.input_bias)}
        for l in range(self.num_layers):
            out = {"kernel": _lecun(g, dev, H, dh, hid, fan_in=hid)}
            if self.with_bias:
                out["bias"] = torch.zeros(hid, device=dev)
            lp = {"MultiHeadDotProductAttention_0": {"query": proj(), "key": proj(), "value": proj(), "out": out}}
            if self.layer_norm_mode != "null":
                for nm in ("LayerNorm_0", "LayerNorm_1"):
                    lp[nm] = {"scale": torch.ones(hid, device=dev), "bias": torch.zeros(hid, device=dev)}
            fan = hid
            for j, h in enumerate(list(self.mlp_hidden_dims) + [hid]):
                lp[f"Dense_{j}"] = dense(fan, h)
                fan = h
            bb[f"PsiformerLayer_{l}"] = lp
        split = self.orbitals_spin_split and self.nspins[0] > 0 and self.nspins[1] > 0
        tree = {"backbone_layer": bb}
        tree.update(_head_params(g, dev, self.nspins, A, self.ndets, hid, split, self.envelope, self.bias_orbitals,
                                 self.jastrow == "simple_ee", self.jastrow_alpha_init))
        return {"params": tree}

    def _handle(self, params, n_atoms: int):
        return _marshal.psiformer_handle(params, self.nspins, n_atoms, self.ndets, self.num_layers, self.num_heads,
                                         self.heads_dim, tuple(self.mlp_hidden_dims), self.layer_norm_mode,
                                         self.envelope, self.orbitals_spin_split, self.rescale,
                                         self.jastrow == "simple_ee")


@dataclass
class SolidWavefunction:
    """Periodic FermiNet with complex orbitals (reference app/solid/wavefunction.py:40-147); same fields and defaults.
    ``klist`` (n, 3) is the k-point of every orbital (set by the workflow after SCF in the reference)."""

    nspins: tuple = (1, 1)
    simulation_lattice: object = None
    primitive_lattice: object = None
    klist: object = None
    ndets: int = 16
    hidden_dims_single: list = field(default_factory=lambda: [256] * 4)
    hidden_dims_double: list = field(default_factory=lambda: [32] * 4)
    distance_type: str = "tri"
    sym_type: str = "minimal"
    envelope_type: str = "abs_isotropic"
    orbitals_spin_split: bool = True
    full_det: bool = True

    def __post_init__(self):
        if self.distance_type not in ("tri", "nu") or self.sym_type not in ("minimal", "fcc", "bcc", "hexagonal"):
            raise ValueError(f"unknown distance_type / sym_type: {self.distance_type!r} / {self.sym_type!r}")
        if len(self.hidden_dims_single) != len(self.hidden_dims_double):
            raise ValueError("hidden_dims_single and hidden_dims_double must have the same length")
        if self.simulation_lattice is None or self.primitive_lattice is None or self.klist is None:
            raise ValueError("SolidWavefunction needs simulation_lattice, primitive_lattice and klist")

    def init_params(self, data, rngs) -> dict:
        """Flax-layout tree: backbone_layer/Dense_*, real_orbital_layer, imag_orbital_layer, envelope_layer."""
        dev = data.electrons.device
        g = rngs if isinstance(rngs, torch.Generator) else torch.Generator(device="cpu").manual_seed(int(rngs))
        n_up, n_dn = self.nspins
        n = n_up + n_dn
        A = data.primitive_atoms.shape[0]
        nch = 2 if (n_up > 0 and n_dn > 0) else 1
        bb = {}
        fw = 4 if self.distance_type == "nu" else 7
        d1, d2 = fw * A, fw
        idx = 0
        L = len(self.hidden_dims_single)
        for layer in range(L):
            h1 = self.hidden_dims_single[layer]
            bb[f"Dense_{idx}"] = {"kernel": _lecun(g, dev, d1 * (1 + nch) + d2 * nch, h1), "bias": torch.zeros(h1, device=dev)}
            idx += 1
            if layer < L - 1:
                h2 = self.hidden_dims_double[layer]
                bb[f"Dense_{idx}"] = {"kernel": _lecun(g, dev, d2, h2), "bias": torch.zeros(h2, device=dev)}
                idx += 1
                d2 = h2
            d1 = h1
        split = self.orbitals_spin_split and nch == 2
        head_r = _head_params(g, dev, self.nspins, A, self.ndets, d1, split, self.envelope_type, False, False, 1.0)
        head_i = _head_params(g, dev, self.nspins, A, self.ndets, d1, split, "null", False, False, 1.0)
        return {"params": {"backbone_layer": bb, "real_orbital_layer": head_r["orbital_layer"],
                           "imag_orbital_layer": head_i["orbital_layer"], "envelope_layer": head_r["envelope_layer"]}}

    def _handle(self, params, n_prim_atoms: int, device):
        kl = torch.as_tensor(self.klist, dtype=torch.float32).to(device).contiguous()
        return _marshal.solid_handle(params, self.nspins, n_prim_atoms, self.simulation_lattice, self.primitive_lattice,
                                     kl, self.ndets, self.hidden_dims_single, self.hidden_dims_double,
                                     self.envelope_type, self.orbitals_spin_split, self.distance_type, self.sym_type)

    def _sampling_handles(self, params, data):
        return (self._handle(params, data.primitive_atoms.shape[0], data.electrons.device),
                _marshal.system_handle(data.primitive_atoms, None))

    def evaluate(self, params, data) -> dict:
        """``{"logpsi"}``: complex log psi per walker (reference LogDet complex output)."""
        el, squeeze = _batched(data.electrons)
        rt = runtime(el.device)
        wf = self._handle(params, data.primitive_atoms.shape[0], el.device)
        sysh = _marshal.system_handle(data.primitive_atoms, None)
        re, im = rt.logpsi(wf, sysh, el)
        lp = torch.complex(re, im)
        return {"logpsi": lp[0] if squeeze else lp}

    def logpsi(self, params, data):
        return self.evaluate(params, data)["logpsi"]

    def orbitals(self, params, data) -> torch.Tensor:
        """Complex orbital matrices ``(W, ndets, n, n)`` including envelope and Bloch phase
        (reference app/solid/wavefunction.py ``get_orbitals`` / ``orbitals``)."""
        el, squeeze = _batched(data.electrons)
        wf, sysh = self._sampling_handles(params, data)
        orb = runtime(el.device).orbitals(wf, sysh, el, self.ndets, complex_valued=True)
        return orb[0] if squeeze else orb

    def phase_logpsi(self, params, data):
        lp = self.logpsi(params, data)
        return torch.exp(1j * lp.imag), lp.real

    def __call__(self, params, data):
        return self.evaluate(params, data)

    def local_energy(self, params, data, ewald=None, sums=None) -> dict:
        """Complex value / gradient / Laplacian / kinetic energy per walker, plus the Ewald potential when an
        :class:`jaqmc_b200.ewald.EwaldSum` of the simulation cell is given."""
        el, _ = _batched(data.electrons)
        rt = runtime(el.device)
        wf = self._handle(params, data.primitive_atoms.shape[0], el.device)
        sysh = _marshal.system_handle(data.primitive_atoms, None)
        return rt.local_energy_complex(wf, sysh, el, ewald, data.atoms if ewald is not None else None,
                                       data.charges if ewald is not None else None, sums=sums)


@dataclass
class HydrogenAtom(Wavefunction):
    """One-parameter demo wavefunction ``log psi = alpha |r|`` of ``jaqmc hydrogen-atom train``
    (reference app/hydrogen_atom.py:28-35); ``initial_alpha = -0.8``.  The potential ``-1/|r|`` (``:36-39``) is the
    Coulomb kernel with a unit charge at the origin."""

    initial_alpha: float = -0.8
    nspins: tuple = (1, 0)

    def init_params(self, data, rngs=None) -> dict:
        return {"params": {"alpha": torch.full((1,), float(self.initial_alpha), device=data.electrons.device)}}

    def _handle(self, params, n_atoms: int):
        return _marshal.hydrogen_handle(params, sum(self.nspins))
