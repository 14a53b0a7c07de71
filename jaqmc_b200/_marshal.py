"""Flax-layout parameter trees -> C-ABI descriptors (``include/jaqmc_b200.h``).

The trees are the reference's own, e.g. for FermiNet
``params/backbone_layer/Dense_{i}/{kernel,bias}``, ``params/orbital_layer/SplitChannelDense_0/DenseGeneral_{s}/kernel``,
``params/envelope_layer/{_env_up,_env_down,_env}/{pi,sigma}``; leaves are float32 tensors on the compute device.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _abi


class Handle:
    """A ctypes struct plus everything that must outlive the call (sub-structs, tensors)."""

    def __init__(self, struct, keep):
        self.struct = struct
        self._keep = keep


def _leaf(t, name, shape=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor leaf, got {type(t).__name__}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t if t.is_contiguous() else t.contiguous()


def system_handle(atoms: torch.Tensor, charges: torch.Tensor | None) -> Handle:
    atoms = _leaf(atoms, "atoms")
    if atoms.dim() != 2 or atoms.shape[1] != 3:
        raise ValueError(f"atoms: expected (n_atoms, 3), got {tuple(atoms.shape)}")
    keep = [atoms]
    s = _abi.System()
    s.atoms = atoms.data_ptr()
    s.n_atoms = atoms.shape[0]
    if charges is not None:
        charges = _leaf(charges, "charges", (atoms.shape[0],))
        keep.append(charges)
        s.charges = charges.data_ptr()
    else:
        s.charges = None
    return Handle(s, keep)


def ferminet_handle(params, nspins, n_atoms, ndets, hidden_dims_single, hidden_dims_double, envelope="abs_isotropic",
                    orbitals_spin_split=True) -> Handle:
    """Descriptor for ``FermiNetWavefunction`` (reference app/molecule/wavefunction/ferminet.py:43-74)."""
    n_up, n_dn = int(nspins[0]), int(nspins[1])
    n = n_up + n_dn
    L = len(hidden_dims_single)
    if L < 1 or L > _abi.MAX_LAYERS or len(hidden_dims_double) != L:
        raise ValueError(f"hidden_dims_single/double must have the same length in [1, {_abi.MAX_LAYERS}]")
    if envelope not in _abi.ENVELOPE:
        raise ValueError(f"Unknown envelope: {envelope!r}")
    nch = 2 if (n_up > 0 and n_dn > 0) else 1
    split = bool(orbitals_spin_split) and nch == 2
    p = params["params"] if "params" in params else params
    cfg = _abi.FerminetConfig()
    cfg.n_up, cfg.n_dn, cfg.n_atoms, cfg.ndets, cfg.n_layers = n_up, n_dn, int(n_atoms), int(ndets), L
    for i in range(L):
        cfg.hidden_single[i] = int(hidden_dims_single[i])
        cfg.hidden_double[i] = int(hidden_dims_double[i])
    cfg.envelope_type = _abi.ENVELOPE[envelope]
    cfg.orbitals_spin_split = int(split)
    ps = _abi.FerminetParams()
    keep = [cfg, ps]
    bb = p["backbone_layer"]
    d1, d2 = 4 * n_atoms, 4
    idx = 0
    for layer in range(L):
        fan_in = d1 * (1 + nch) + d2 * nch
        h1 = int(hidden_dims_single[layer])
        k = _leaf(bb[f"Dense_{idx}"]["kernel"], f"backbone_layer/Dense_{idx}/kernel", (fan_in, h1))
        b = _leaf(bb[f"Dense_{idx}"]["bias"], f"backbone_layer/Dense_{idx}/bias", (h1,))
        keep += [k, b]
        ps.single_kernel[layer], ps.single_bias[layer] = k.data_ptr(), b.data_ptr()
        idx += 1
        if layer < L - 1:
            h2 = int(hidden_dims_double[layer])
            k = _leaf(bb[f"Dense_{idx}"]["kernel"], f"backbone_layer/Dense_{idx}/kernel", (d2, h2))
            b = _leaf(bb[f"Dense_{idx}"]["bias"], f"backbone_layer/Dense_{idx}/bias", (h2,))
            keep += [k, b]
            ps.double_kernel[layer], ps.double_bias[layer] = k.data_ptr(), b.data_ptr()
            idx += 1
            d2 = h2
        d1 = h1
    ol = p["orbital_layer"]
    if split:
        for s in range(2):
            k = _leaf(ol["SplitChannelDense_0"][f"DenseGeneral_{s}"]["kernel"],
                      f"orbital_layer/SplitChannelDense_0/DenseGeneral_{s}/kernel", (d1, ndets, n))
            keep.append(k)
            ps.orbital_kernel[s] = k.data_ptr()
    else:
        k = _leaf(ol["DenseGeneral_0"]["kernel"], "orbital_layer/DenseGeneral_0/kernel", (d1, ndets, n))
        keep.append(k)
        ps.orbital_kernel[0] = k.data_ptr()
    if envelope != "null":
        el = p["envelope_layer"]
        names = ["_env_up", "_env_down"] if split else ["_env"]
        for s, nm in enumerate(names):
            pi = _leaf(el[nm]["pi"], f"envelope_layer/{nm}/pi", (n, n_atoms, ndets))
            sg = _leaf(el[nm]["sigma"], f"envelope_layer/{nm}/sigma", (n, n_atoms, ndets))
            keep += [pi, sg]
            ps.env_pi[s], ps.env_sigma[s] = pi.data_ptr(), sg.data_ptr()
    wf = _abi.Wavefunction()
    wf.kind = _abi.WF_FERMINET
    wf.config = C.cast(C.pointer(cfg), C.c_void_p)
    wf.params = C.cast(C.pointer(ps), C.c_void_p)
    return Handle(wf, keep)
