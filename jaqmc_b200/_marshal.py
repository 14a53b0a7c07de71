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


def _sigma_shape(envelope, n, n_atoms, ndets):
    """``sigma`` of the envelope: (n_orb, A, D), or (n_orb, A, 3, D) for the diagonal one (output/envelope.py:150-156)."""
    return (n, n_atoms, 3, ndets) if envelope == "diagonal" else (n, n_atoms, ndets)


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
                    orbitals_spin_split=True, use_last_layer=False) -> Handle:
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
    cfg.use_last_layer = int(bool(use_last_layer))
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
        if layer < L - 1 or use_last_layer:
            h2 = int(hidden_dims_double[layer])
            k = _leaf(bb[f"Dense_{idx}"]["kernel"], f"backbone_layer/Dense_{idx}/kernel", (d2, h2))
            b = _leaf(bb[f"Dense_{idx}"]["bias"], f"backbone_layer/Dense_{idx}/bias", (h2,))
            keep += [k, b]
            ps.double_kernel[layer], ps.double_bias[layer] = k.data_ptr(), b.data_ptr()
            idx += 1
            d2 = h2
        d1 = h1
    if use_last_layer:   # the orbitals see the aggregated features (backbone/ferminet.py:45-47)
        d1 = d1 * (1 + nch) + d2 * nch
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
            sg = _leaf(el[nm]["sigma"], f"envelope_layer/{nm}/sigma", _sigma_shape(envelope, n, n_atoms, ndets))
            keep += [pi, sg]
            ps.env_pi[s], ps.env_sigma[s] = pi.data_ptr(), sg.data_ptr()
    wf = _abi.Wavefunction()
    wf.kind = _abi.WF_FERMINET
    wf.config = C.cast(C.pointer(cfg), C.c_void_p)
    wf.params = C.cast(C.pointer(ps), C.c_void_p)
    h = Handle(wf, keep)
    h.config_struct, h.params_struct = cfg, ps   # typed views (parameter-gradient entry point)
    return h


class _Binder:
    """Collects leaves (keeping them alive) while filling pointer fields."""

    def __init__(self):
        self.keep = []

    def leaf(self, tree, path, shape=None, optional=False):
        node = tree
        for key in path.split("/"):
            if not isinstance(node, dict) or key not in node:
                if optional:
                    return None
                raise KeyError(f"parameter tree has no leaf {path!r}")
            node = node[key]
        t = _leaf(node, path, shape)
        self.keep.append(t)
        return t.data_ptr()


def _fill_head(b: _Binder, head: "_abi.HeadParams", p, n, n_atoms, ndets, hidden, split, envelope, jastrow):
    ol = p["orbital_layer"]
    if split:
        for s in range(2):
            base = f"SplitChannelDense_0/DenseGeneral_{s}"
            head.orbital_kernel[s] = b.leaf(ol, f"{base}/kernel", (hidden, ndets, n))
            head.orbital_bias[s] = b.leaf(ol, f"{base}/bias", (ndets, n), optional=True)
    else:
        head.orbital_kernel[0] = b.leaf(ol, "DenseGeneral_0/kernel", (hidden, ndets, n))
        head.orbital_bias[0] = b.leaf(ol, "DenseGeneral_0/bias", (ndets, n), optional=True)
    if envelope != "null":
        el = p["envelope_layer"]
        for s, nm in enumerate(["_env_up", "_env_down"] if split else ["_env"]):
            head.env_pi[s] = b.leaf(el, f"{nm}/pi", (n, n_atoms, ndets))
            head.env_sigma[s] = b.leaf(el, f"{nm}/sigma", _sigma_shape(envelope, n, n_atoms, ndets))
    if jastrow:
        head.jastrow_alpha_par = b.leaf(p["jastrow_layer"], "alpha_par", (1,))
        head.jastrow_alpha_anti = b.leaf(p["jastrow_layer"], "alpha_anti", (1,))


def _wf_handle(kind, cfg, ps, keep):
    wf = _abi.Wavefunction()
    wf.kind = kind
    wf.config = C.cast(C.pointer(cfg), C.c_void_p)
    wf.params = C.cast(C.pointer(ps), C.c_void_p)
    return Handle(wf, [cfg, ps, *keep])


def lapnet_handle(params, nspins, n_atoms, ndets=16, num_layers=4, num_heads=4, heads_dim=64, num_local_updates=2,
                  envelope="abs_isotropic", rescale=True, jastrow=True, use_layernorm=False) -> Handle:
    """Descriptor for ``LapNetWavefunction`` (reference app/molecule/wavefunction/lapnet.py:63-115)."""
    n_up, n_dn = int(nspins[0]), int(nspins[1])
    n = n_up + n_dn
    if not 1 <= num_layers <= _abi.MAX_LAYERS:
        raise ValueError(f"num_layers must be in [1, {_abi.MAX_LAYERS}]")
    if not 0 <= num_local_updates <= 4:
        raise ValueError("num_local_updates must be in [0, 4]")
    if envelope not in _abi.ENVELOPE:
        raise ValueError(f"Unknown envelope: {envelope!r}")
    p = params["params"] if "params" in params else params
    hid = num_heads * heads_dim
    cfg = _abi.LapnetConfig(n_up, n_dn, int(n_atoms), int(ndets), int(num_layers), int(num_heads), int(heads_dim),
                            int(num_local_updates), _abi.ENVELOPE[envelope], int(bool(rescale)), int(bool(use_layernorm)))
    ps = _abi.LapnetParams()
    b = _Binder()
    bb = p["backbone_layer"]
    ps.input_kernel = b.leaf(bb, "input_projection/kernel", (4 * n_atoms + 1, hid))
    ps.input_bias = b.leaf(bb, "input_projection/bias", (hid,), optional=True)
    for l in range(num_layers):
        lp = bb[f"layers_{l}"]
        if use_layernorm:
            for nm, fld in (("qk_layernorm", "qk_ln"), ("value_layernorm", "value_ln"),
                            ("post_attention_layernorm", "post_ln")):
                getattr(ps, fld + "_scale")[l] = b.leaf(lp, f"{nm}/scale", (hid,))
                getattr(ps, fld + "_bias")[l] = b.leaf(lp, f"{nm}/bias", (hid,))
        elif "qk_layernorm" in lp or "value_layernorm" in lp:
            raise ValueError("the parameter tree holds LayerNorm leaves but use_layernorm=False")
        ps.qk_kernel[l] = b.leaf(lp, "qk_projection/kernel", (hid, 2 * hid))
        ps.qk_bias[l] = b.leaf(lp, "qk_projection/bias", (2 * hid,), optional=True)
        for nm, fld in (("value_projection", "value"), ("output_projection", "output"), ("value_update", "update")):
            getattr(ps, fld + "_kernel")[l] = b.leaf(lp, f"{nm}/kernel", (hid, hid))
            getattr(ps, fld + "_bias")[l] = b.leaf(lp, f"{nm}/bias", (hid,), optional=True)
        if l < num_layers - 1:
            for j in range(num_local_updates):
                ps.qk_update_kernel[l][j] = b.leaf(lp, f"qk_update_layers_{j}/kernel", (hid, hid))
                ps.qk_update_bias[l][j] = b.leaf(lp, f"qk_update_layers_{j}/bias", (hid,), optional=True)
    split = n_up > 0 and n_dn > 0
    _fill_head(b, ps.head, p, n, n_atoms, ndets, hid, split, envelope, jastrow)
    return _wf_handle(_abi.WF_LAPNET, cfg, ps, b.keep)


def psiformer_handle(params, nspins, n_atoms, ndets=16, num_layers=4, num_heads=4, heads_dim=64, mlp_hidden_dims=(256,),
                     layer_norm_mode="pre", envelope="abs_isotropic", orbitals_spin_split=True, rescale=True,
                     jastrow=True) -> Handle:
    """Descriptor for ``PsiformerWavefunction`` (reference app/molecule/wavefunction/psiformer.py:77-135)."""
    n_up, n_dn = int(nspins[0]), int(nspins[1])
    n = n_up + n_dn
    if not 1 <= num_layers <= _abi.MAX_LAYERS:
        raise ValueError(f"num_layers must be in [1, {_abi.MAX_LAYERS}]")
    if len(mlp_hidden_dims) > _abi.MAX_MLP - 1:
        raise ValueError(f"at most {_abi.MAX_MLP - 1} hidden MLP layers are supported")
    if layer_norm_mode not in _abi.LAYERNORM:
        raise ValueError(f"Unknown layer_norm_mode: {layer_norm_mode!r}")
    if envelope not in _abi.ENVELOPE:
        raise ValueError(f"Unknown envelope: {envelope!r}")
    p = params["params"] if "params" in params else params
    hid = num_heads * heads_dim
    cfg = _abi.PsiformerConfig()
    cfg.n_up, cfg.n_dn, cfg.n_atoms, cfg.ndets = n_up, n_dn, int(n_atoms), int(ndets)
    cfg.num_layers, cfg.num_heads, cfg.heads_dim = int(num_layers), int(num_heads), int(heads_dim)
    cfg.n_mlp_hidden = len(mlp_hidden_dims)
    for j, h in enumerate(mlp_hidden_dims):
        cfg.mlp_hidden[j] = int(h)
    cfg.layer_norm_mode = _abi.LAYERNORM[layer_norm_mode]
    cfg.envelope_type = _abi.ENVELOPE[envelope]
    split = bool(orbitals_spin_split) and n_up > 0 and n_dn > 0
    cfg.orbitals_spin_split = int(split)
    cfg.rescale = int(bool(rescale))
    ps = _abi.PsiformerParams()
    b = _Binder()
    bb = p["backbone_layer"]
    ps.input_kernel = b.leaf(bb, "Dense_0/kernel", (4 * n_atoms + 1, hid))
    ps.input_bias = b.leaf(bb, "Dense_0/bias", (hid,), optional=True)
    for l in range(num_layers):
        lp = bb[f"PsiformerLayer_{l}"]
        mha = lp["MultiHeadDotProductAttention_0"]
        if layer_norm_mode != "null":
            ps.ln0_scale[l] = b.leaf(lp, "LayerNorm_0/scale", (hid,))
            ps.ln0_bias[l] = b.leaf(lp, "LayerNorm_0/bias", (hid,))
            ps.ln1_scale[l] = b.leaf(lp, "LayerNorm_1/scale", (hid,))
            ps.ln1_bias[l] = b.leaf(lp, "LayerNorm_1/bias", (hid,))
        for nm, fld in (("query", "q"), ("key", "k"), ("value", "v")):
            getattr(ps, fld + "_kernel")[l] = b.leaf(mha, f"{nm}/kernel", (hid, num_heads, heads_dim))
            getattr(ps, fld + "_bias")[l] = b.leaf(mha, f"{nm}/bias", (num_heads, heads_dim), optional=True)
        ps.out_kernel[l] = b.leaf(mha, "out/kernel", (num_heads, heads_dim, hid))
        ps.out_bias[l] = b.leaf(mha, "out/bias", (hid,), optional=True)
        fan = hid
        dims = [int(h) for h in mlp_hidden_dims] + [hid]
        for j, h in enumerate(dims):
            ps.mlp_kernel[l][j] = b.leaf(lp, f"Dense_{j}/kernel", (fan, h))
            ps.mlp_bias[l][j] = b.leaf(lp, f"Dense_{j}/bias", (h,), optional=True)
            fan = h
    _fill_head(b, ps.head, p, n, n_atoms, ndets, hid, split, envelope, jastrow)
    return _wf_handle(_abi.WF_PSIFORMER, cfg, ps, b.keep)


def solid_handle(params, nspins, n_prim_atoms, simulation_lattice, primitive_lattice, klist, ndets=16,
                 hidden_dims_single=(256,) * 4, hidden_dims_double=(32,) * 4, envelope="abs_isotropic",
                 orbitals_spin_split=True, distance_type="tri", sym_type="minimal") -> Handle:
    """Descriptor for ``SolidWavefunction`` (reference app/solid/wavefunction.py:40-147); ``klist`` (n, 3) on device.
    ``distance_type`` 'tri' | 'nu' and ``sym_type`` 'minimal' | 'fcc' | 'bcc' | 'hexagonal' as in geometry/pbc.py."""
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
    cfg = _abi.SolidConfig()
    net = cfg.net
    net.n_up, net.n_dn, net.n_atoms, net.ndets, net.n_layers = n_up, n_dn, int(n_prim_atoms), int(ndets), L
    for i in range(L):
        net.hidden_single[i] = int(hidden_dims_single[i])
        net.hidden_double[i] = int(hidden_dims_double[i])
    net.envelope_type = _abi.ENVELOPE[envelope]
    net.orbitals_spin_split = int(split)
    if str(distance_type) not in _abi.DISTANCE or str(sym_type) not in _abi.SYMMETRY:
        raise ValueError(f"Unknown distance_type / sym_type: {distance_type!r} / {sym_type!r}")
    cfg.distance_type, cfg.sym_type = _abi.DISTANCE[str(distance_type)], _abi.SYMMETRY[str(sym_type)]
    fw = 4 if str(distance_type) == "nu" else 7   # features per electron-atom / electron-electron pair
    for name, lat in (("simulation_lattice", simulation_lattice), ("primitive_lattice", primitive_lattice)):
        vals = [float(v) for v in torch.as_tensor(lat).reshape(-1).tolist()]
        if len(vals) != 9:
            raise ValueError(f"{name}: expected a (3, 3) lattice")
        for i, v in enumerate(vals):
            getattr(cfg, name)[i] = v
    ps = _abi.SolidParams()
    b = _Binder()
    bb = p["backbone_layer"]
    d1, d2 = fw * n_prim_atoms, fw
    idx = 0
    for layer in range(L):
        h1 = int(hidden_dims_single[layer])
        ps.net.single_kernel[layer] = b.leaf(bb, f"Dense_{idx}/kernel", (d1 * (1 + nch) + d2 * nch, h1))
        ps.net.single_bias[layer] = b.leaf(bb, f"Dense_{idx}/bias", (h1,))
        idx += 1
        if layer < L - 1:
            h2 = int(hidden_dims_double[layer])
            ps.net.double_kernel[layer] = b.leaf(bb, f"Dense_{idx}/kernel", (d2, h2))
            ps.net.double_bias[layer] = b.leaf(bb, f"Dense_{idx}/bias", (h2,))
            idx += 1
            d2 = h2
        d1 = h1
    for part, fld in (("real_orbital_layer", ps.real_orbital_kernel), ("imag_orbital_layer", ps.imag_orbital_kernel)):
        ol = p[part]
        if split:
            for s_ in range(2):
                fld[s_] = b.leaf(ol, f"SplitChannelDense_0/DenseGeneral_{s_}/kernel", (d1, ndets, n))
        else:
            fld[0] = b.leaf(ol, "DenseGeneral_0/kernel", (d1, ndets, n))
    if envelope != "null":
        el = p["envelope_layer"]
        for s_, nm in enumerate(["_env_up", "_env_down"] if split else ["_env"]):
            ps.net.env_pi[s_] = b.leaf(el, f"{nm}/pi", (n, n_prim_atoms, ndets))
            ps.net.env_sigma[s_] = b.leaf(el, f"{nm}/sigma", (n, n_prim_atoms, ndets))
    kl = _leaf(klist, "klist", (n, 3))
    b.keep.append(kl)
    ps.klist = kl.data_ptr()
    return _wf_handle(_abi.WF_SOLID_FERMINET, cfg, ps, b.keep)


def hydrogen_handle(params, n_electrons: int = 1) -> Handle:
    """Descriptor for the ``HydrogenAtom`` demo wavefunction (reference app/hydrogen_atom.py:28-35)."""
    p = params["params"] if "params" in params else params
    alpha = _leaf(p["alpha"].reshape(1) if p["alpha"].dim() == 0 else p["alpha"], "alpha", (1,))
    cfg = _abi.HydrogenConfig(int(n_electrons))
    ps = _abi.HydrogenParams()
    ps.alpha = alpha.data_ptr()
    return _wf_handle(_abi.WF_HYDROGEN, cfg, ps, [alpha])
