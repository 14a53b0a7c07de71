"""Single-walker wavefunction restatements (oracle; test infrastructure only).

float64 PyTorch-CPU restatement of the reference's wavefunction graphs.  Every function takes
``electrons`` either as a plain ``(n, 3)`` tensor (value route, also used under autograd for the
brute-force Hessian) or as an ``oracle.lap.Lap`` seed (forward-Laplacian route).  Parameters are
nested dicts in the reference's Flax tree layout: ``Dense`` kernels ``(in, out)``,
``DenseGeneral`` kernels ``(in, ndets, n)``, envelope ``pi``/``sigma`` ``(n_orb, n_atoms, ndets)``.

Reference files followed (``/root/reference/src/jaqmc``):

* features ............ ``wavefunction/input/atomic.py:46-79,108-147``, ``geometry/obc.py:7-88``,
                        ``geometry/pbc.py:97-111,282-324,347-381``
* FermiNet streams .... ``wavefunction/backbone/ferminet.py:29-90``, ``utils/array.py:24-45``
* LapNet .............. ``wavefunction/backbone/lapnet/_backbone.py:16-263``, ``_attention.py:20-34``
* Psiformer ........... ``wavefunction/backbone/psiformer.py:60-187`` (flax ``LayerNorm``,
                        ``MultiHeadDotProductAttention`` restated from flax 0.12.6 semantics)
* orbitals/envelope ... ``wavefunction/output/orbital.py:59-78``, ``wavefunction/output/envelope.py:98-163``
* logdet .............. ``wavefunction/output/logdet.py:53-79``
* Jastrow ............. ``wavefunction/jastrow.py:47-122``
* compositions ........ ``app/molecule/wavefunction/{ferminet,lapnet,psiformer}.py``,
                        ``app/solid/wavefunction.py:91-147``
"""

from __future__ import annotations

import math

import torch

from . import lap as L

F64 = torch.float64


# ---------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------
def split_nonempty_channels(x, nspins):
    """``utils/array.py:24-45``: split axis 0 by spin, dropping empty channels."""
    sizes = [s for s in nspins if s > 0]
    if len(sizes) <= 1:
        return [x]
    return [x[: sizes[0]], x[sizes[0] :]]


def _norm_last(v):
    return L.sqrt(L.sum_(L.square(v), dim=-1))


def pair_displacements_within(pos):
    """``geometry/obc.py:7-48``: ``disp_ij = r_i - r_j``; the diagonal is shifted by ``eye`` before
    the norm and masked afterwards so derivatives stay finite."""
    n = pos.shape[0]
    pi = L.linear_map(lambda t: t[:, None, :], pos)
    pj = L.linear_map(lambda t: t[None, :, :], pos)
    disp = pi - pj
    eye = torch.eye(n, dtype=L.value(pos).dtype)
    r = _norm_last(disp + eye[..., None]) * (1.0 - eye)
    return disp, r


def pair_displacements_between(pos_a, pos_b: torch.Tensor):
    """``geometry/obc.py:51-88``."""
    pa = L.linear_map(lambda t: t[:, None, :], pos_a)
    disp = pa - pos_b[None, :, :]
    return disp, _norm_last(disp)


def molecule_features(electrons, atoms, rescale: bool):
    """``wavefunction/input/atomic.py:46-79``."""
    ee_vec, r_ee = pair_displacements_within(electrons)
    ae_vec, r_ae = pair_displacements_between(electrons, atoms)
    n = electrons.shape[0]
    un = lambda a: L.linear_map(lambda t: t[..., None], a)  # noqa: E731
    if rescale:
        log_r_ae = un(L.log1p(r_ae))
        ae_features = L.cat([log_r_ae, ae_vec * log_r_ae / un(r_ae)], dim=2)
        # NOTE: the reference divides by r_ee whose diagonal is exactly 0 -> NaN on the diagonal of
        # ee_features; LapNet/Psiformer never consume ee_features, so it is not built here.
        ee_features = None
    else:
        ae_features = L.cat([un(r_ae), ae_vec], dim=2)
        ee_features = L.cat([un(r_ee), ee_vec], dim=2)
    ae_features = L.reshape(ae_features, n, -1)
    return dict(ae_features=ae_features, ee_features=ee_features, r_ae=r_ae, ae_vec=ae_vec, r_ee=r_ee)


# ---------------------------------------------------------------------------------------
# FermiNet backbone
# ---------------------------------------------------------------------------------------
def _residual(x, y):
    if tuple(x.shape) == tuple(y.shape):
        return (x + y) / math.sqrt(2.0)
    return y


def ferminet_aggregate(h_one, h_two, nspins):
    """``backbone/ferminet.py:65-90``: means over the FIRST electron axis of each spin block."""
    n = h_one.shape[0]
    g_one = [
        L.broadcast_to(L.mean(h, dim=0, keepdim=True), (n, h_one.shape[1]))
        for h in split_nonempty_channels(h_one, nspins)
    ]
    g_two = [L.mean(h, dim=0) for h in split_nonempty_channels(h_two, nspins)]
    return L.cat([h_one, *g_one, *g_two], dim=-1)


def fermi_layers(p, h_one, h_two, nspins, n_layers, use_last_layer=False):
    """``backbone/ferminet.py:29-63``; Dense_{2l} single, Dense_{2l+1} double.  ``use_last_layer=True`` also updates
    the double stream in the last layer and returns the aggregated features (``:45-47``)."""
    idx = 0
    for layer in range(n_layers):
        d = p[f"Dense_{idx}"]
        idx += 1
        h_in = ferminet_aggregate(h_one, h_two, nspins)
        h_one = _residual(h_one, L.tanh(L.dense(h_in, d["kernel"], d["bias"])))
        if layer < n_layers - 1 or use_last_layer:
            d = p[f"Dense_{idx}"]
            idx += 1
            h_two = _residual(h_two, L.tanh(L.dense(h_two, d["kernel"], d["bias"])))
    if use_last_layer:
        return ferminet_aggregate(h_one, h_two, nspins), h_two
    return h_one, h_two


# ---------------------------------------------------------------------------------------
# output heads
# ---------------------------------------------------------------------------------------
def orbital_projection(p, h_one, nspins):
    """``output/orbital.py:59-78``: per-spin ``DenseGeneral([ndets, n])`` -> ``(ndets, n_elec, n_orb)``."""
    active = [s for s in nspins if s > 0]
    if "SplitChannelDense_0" in p and len(active) > 1:
        sp = p["SplitChannelDense_0"]
        parts = [
            L.dense(h, sp[f"DenseGeneral_{i}"]["kernel"], sp[f"DenseGeneral_{i}"].get("bias"))
            for i, h in enumerate(split_nonempty_channels(h_one, nspins))
        ]
        orb = L.cat(parts, dim=0)
    else:
        dg = p["DenseGeneral_0"]
        orb = L.dense(h_one, dg["kernel"], dg.get("bias"))
    return L.transpose(orb, 1, 0, 2)


def _isotropic_envelope(p, r_ae, is_abs=True):
    """``output/envelope.py:123-140``: ``sum_I pi * exp(-|sigma * r|)`` -> ``(ndets, n_elec, n_orb)``
    (``is_abs=False``: exponent ``sigma * -r``)."""
    pi, sigma = p["pi"], p["sigma"]  # (n_orb, n_atoms, ndets)
    r = L.linear_map(lambda t: t[:, None, :, None], r_ae)  # (n_elec,1,n_atoms,1)
    sr = r * sigma
    exponent = -L.abs_(sr) if is_abs else -sr
    env = L.sum_(L.exp(exponent) * pi, dim=2)  # (n_elec, n_orb, ndets)
    return L.transpose(env, 2, 0, 1)


def _diagonal_envelope(p, ae_vec):
    """``output/envelope.py:143-163``: ``sum_I pi * exp(-||sigma (.) r_vec||)`` with ``sigma (n_orb, n_atoms, ndim, ndets)``."""
    pi, sigma = p["pi"], p["sigma"]
    v = L.linear_map(lambda t: t[:, None, :, :, None], ae_vec)  # (n_elec,1,n_atoms,ndim,1)
    scaled = v * sigma
    r_scaled = L.sqrt(L.sum_(L.square(scaled), dim=3))  # (n_elec, n_orb, n_atoms, ndets)
    env = L.sum_(L.exp(-r_scaled) * pi, dim=2)
    return L.transpose(env, 2, 0, 1)


def envelope(p, r_ae, nspins, envelope_type="abs_isotropic", ae_vec=None):
    """``output/envelope.py:98-120``; returns ``None`` for the ``null`` envelope (a factor of ones)."""
    if envelope_type == "null":
        return None
    if envelope_type == "diagonal":
        one = lambda q, lo, hi: _diagonal_envelope(q, ae_vec[lo:hi])  # noqa: E731
    elif envelope_type in ("isotropic", "abs_isotropic"):
        one = lambda q, lo, hi: _isotropic_envelope(q, r_ae[lo:hi], envelope_type == "abs_isotropic")  # noqa: E731
    else:
        raise ValueError(f"Unknown envelope: {envelope_type!r}")
    n_up, n = nspins[0], nspins[0] + nspins[1]
    if "_env_up" in p:
        return L.cat([one(p["_env_up"], 0, n_up), one(p["_env_down"], n_up, n)], dim=1)
    return one(p["_env"], 0, n)


def _apply_envelope(orb, p, emb, nspins, envelope_type):
    env = envelope(p.get("envelope_layer", {}), emb["r_ae"], nspins, envelope_type, emb.get("ae_vec"))
    return orb if env is None else orb * env


def logdet_sum(orbitals):
    """``output/logdet.py:53-79``: log-sum-exp over determinants.  Returns ``(sign_or_None, logpsi)``.

    The ``max`` shift is a constant (its derivative contribution cancels analytically), matching the
    reference's ``logmax`` trick up to rounding.
    """
    signs, lds = L.logdet(orbitals)
    ldx = L.value(lds)
    if ldx.is_complex():
        logmax = ldx.real.max().detach()
        s = L.sum_(L.exp(lds - logmax), dim=0)
        return None, L.log(s) + logmax
    logmax = ldx.max().detach()
    s = L.sum_(L.exp(lds - logmax) * signs, dim=0)
    sx = L.value(s)
    return torch.sign(sx), L.log(L.abs_(s)) + logmax


def simple_ee_jastrow(p, r_ee, nspins):
    """``wavefunction/jastrow.py:47-122``."""
    n_up, n_dn = nspins
    a_par, a_anti = p["alpha_par"][0], p["alpha_anti"][0]
    total = 0.0
    n = n_up + n_dn
    iu = torch.triu_indices(n, n, offset=1)
    same = ((iu[0] < n_up) == (iu[1] < n_up))
    r = L.linear_map(lambda t: t[iu[0], iu[1]], r_ee)
    c = torch.where(same, 0.25, 0.5).to(F64)
    alpha = torch.where(same, a_par, a_anti)
    if r.shape[0] == 0:
        return torch.zeros((), dtype=F64)
    total = L.sum_(-(c * alpha**2) / (r + alpha), dim=0)
    return total


# ---------------------------------------------------------------------------------------
# FermiNet
# ---------------------------------------------------------------------------------------
def ferminet_orbitals(params, electrons, atoms, nspins, envelope_type="abs_isotropic", use_last_layer=False):
    """``(ndets, n, n)`` orbital matrices (``app/molecule/wavefunction/ferminet.py:110-138`` ``orbitals``)."""
    p = params["params"]
    n_layers = (len(p["backbone_layer"]) + 1) // 2
    emb = molecule_features(electrons, atoms, rescale=False)
    h_one, _ = fermi_layers(p["backbone_layer"], emb["ae_features"], emb["ee_features"], nspins, n_layers,
                            use_last_layer)
    orb = orbital_projection(p["orbital_layer"], h_one, nspins)
    return _apply_envelope(orb, p, emb, nspins, envelope_type)


def ferminet_logpsi(params, electrons, atoms, nspins, envelope_type="abs_isotropic", use_last_layer=False):
    """``app/molecule/wavefunction/ferminet.py:76-108`` -> ``(sign, logpsi)``."""
    return logdet_sum(ferminet_orbitals(params, electrons, atoms, nspins, envelope_type, use_last_layer))


# ---------------------------------------------------------------------------------------
# flax building blocks used by LapNet / Psiformer
# ---------------------------------------------------------------------------------------
def layer_norm(p, x, eps):
    """flax ``nn.LayerNorm`` (0.12.6): ``var = mean(x^2) - mean(x)^2`` (``use_fast_variance=True``),
    ``y = (x - mean) * rsqrt(var + eps) * scale + bias``, reduction over the last axis."""
    mu = L.mean(x, dim=-1, keepdim=True)
    mu2 = L.mean(L.square(x), dim=-1, keepdim=True)
    var = mu2 - L.square(mu)
    y = (x - mu) * L.rsqrt(var + eps)
    return y * p["scale"] + p["bias"]


def softmax_last(logits):
    m = L.stop_gradient(logits).max(dim=-1, keepdim=True).values
    w = L.exp(logits - m)
    return w / L.sum_(w, dim=-1, keepdim=True)


def attention_core(q, k, v):
    """``softmax(q k^T / sqrt(d)) v`` on ``(n, heads, d)`` operands (``_attention.py:20-34``; the same
    contraction order flax ``dot_product_attention`` uses)."""
    d = q.shape[-1]
    logits = L.einsum("ihd,jhd->hij", q, k) / math.sqrt(d)
    w = softmax_last(logits)
    return L.einsum("hij,jhd->ihd", w, v)


# ---------------------------------------------------------------------------------------
# LapNet
# ---------------------------------------------------------------------------------------
def lapnet_backbone(p, ae_features, nspins, heads):
    """``backbone/lapnet/_backbone.py:191-263``.  ``use_layernorm=True`` shows up as three extra sub-trees per layer
    (``qk_layernorm``, ``value_layernorm``, ``post_attention_layernorm``, epsilon 1e-6, ``:62-64,81-111,121``)."""
    n_up, n_dn = nspins
    n = n_up + n_dn
    num_layers = sum(1 for k in p if k.startswith("layers_"))
    spins = torch.cat([torch.ones(n_up, dtype=F64), -torch.ones(n_dn, dtype=F64)])
    feats = L.cat([ae_features, spins[:, None]], dim=-1)
    ip = p["input_projection"]
    hs = L.dense(feats, ip["kernel"], ip.get("bias"))
    hd = hs
    qks = []
    for li in range(num_layers):
        lp = p[f"layers_{li}"]
        hs_in = layer_norm(lp["qk_layernorm"], hs, 1e-6) if "qk_layernorm" in lp else hs
        qk = L.dense(hs_in, lp["qk_projection"]["kernel"], lp["qk_projection"].get("bias"))
        half = qk.shape[-1] // 2
        qks.append((qk[:, :half], qk[:, half:]))
        j = 0
        while f"qk_update_layers_{j}" in lp:
            u = lp[f"qk_update_layers_{j}"]
            hs = hs + L.tanh(L.dense(hs, u["kernel"], u.get("bias")))
            j += 1
    for li in range(num_layers):
        lp = p[f"layers_{li}"]
        q, k = qks[li]
        v_in = layer_norm(lp["value_layernorm"], hd, 1e-6) if "value_layernorm" in lp else hd
        v = L.dense(v_in, lp["value_projection"]["kernel"], lp["value_projection"].get("bias"))
        dh = q.shape[-1] // heads
        rs = lambda t: L.reshape(t, n, heads, dh)  # noqa: E731
        att = L.reshape(attention_core(rs(q), rs(k), rs(v)), n, heads * dh)
        att = L.dense(att, lp["output_projection"]["kernel"], lp["output_projection"].get("bias"))
        res = hd + att
        res_n = layer_norm(lp["post_attention_layernorm"], res, 1e-6) if "post_attention_layernorm" in lp else res
        hd = res + L.tanh(L.dense(res_n, lp["value_update"]["kernel"], lp["value_update"].get("bias")))
    return hd


def lapnet_orbitals(params, electrons, atoms, nspins, heads=4, envelope_type="abs_isotropic"):
    p = params["params"]
    emb = molecule_features(electrons, atoms, rescale=True)
    h = lapnet_backbone(p["backbone_layer"], emb["ae_features"], nspins, heads)
    return _apply_envelope(orbital_projection(p["orbital_layer"], h, nspins), p, emb, nspins, envelope_type), emb


def lapnet_logpsi(params, electrons, atoms, nspins, heads=4, envelope_type="abs_isotropic"):
    """``app/molecule/wavefunction/lapnet.py:117-135``."""
    p = params["params"]
    orb, emb = lapnet_orbitals(params, electrons, atoms, nspins, heads, envelope_type)
    sign, lp = logdet_sum(orb)
    if "jastrow_layer" in p:
        lp = lp + simple_ee_jastrow(p["jastrow_layer"], emb["r_ee"], nspins)
    return sign, lp


# ---------------------------------------------------------------------------------------
# Psiformer
# ---------------------------------------------------------------------------------------
def psiformer_backbone(p, ae_features, nspins, layer_norm_mode="pre"):
    """``backbone/psiformer.py:143-187``; layers ``:60-99`` in the three LayerNorm modes (pre is the default)."""
    n_up, n_dn = nspins
    n = n_up + n_dn
    spins = torch.cat([torch.ones(n_up, dtype=F64), -torch.ones(n_dn, dtype=F64)])
    feats = L.cat([ae_features, spins[:, None]], dim=-1)
    x = L.dense(feats, p["Dense_0"]["kernel"], p["Dense_0"].get("bias"))
    li = 0
    while f"PsiformerLayer_{li}" in p:
        lp = p[f"PsiformerLayer_{li}"]
        mha = lp["MultiHeadDotProductAttention_0"]
        x_in = layer_norm(lp["LayerNorm_0"], x, 1e-5) if layer_norm_mode == "pre" else x
        q = L.dense(x_in, mha["query"]["kernel"], mha["query"].get("bias"))
        k = L.dense(x_in, mha["key"]["kernel"], mha["key"].get("bias"))
        v = L.dense(x_in, mha["value"]["kernel"], mha["value"].get("bias"))
        att = attention_core(q, k, v)  # (n, heads, d)
        ok = mha["out"]["kernel"]  # (heads, d, out)
        att = L.reshape(att, n, -1)
        out = L.matmul(att, ok.reshape(-1, ok.shape[-1]))
        if "bias" in mha["out"]:
            out = out + mha["out"]["bias"]
        x = x + out
        if layer_norm_mode == "pre":
            m = layer_norm(lp["LayerNorm_1"], x, 1e-5)
        elif layer_norm_mode == "post":
            m = x = layer_norm(lp["LayerNorm_0"], x, 1e-5)
        else:
            m = x
        j = 0
        while f"Dense_{j}" in lp:
            m = L.tanh(L.dense(m, lp[f"Dense_{j}"]["kernel"], lp[f"Dense_{j}"].get("bias")))
            j += 1
        x = x + m
        if layer_norm_mode == "post":
            x = layer_norm(lp["LayerNorm_1"], x, 1e-5)
        li += 1
    return x


def psiformer_orbitals(params, electrons, atoms, nspins, layer_norm_mode="pre", envelope_type="abs_isotropic"):
    p = params["params"]
    emb = molecule_features(electrons, atoms, rescale=True)
    h = psiformer_backbone(p["backbone_layer"], emb["ae_features"], nspins, layer_norm_mode)
    return _apply_envelope(orbital_projection(p["orbital_layer"], h, nspins), p, emb, nspins, envelope_type), emb


def psiformer_logpsi(params, electrons, atoms, nspins, layer_norm_mode="pre", envelope_type="abs_isotropic"):
    """``app/molecule/wavefunction/psiformer.py:137-167``."""
    p = params["params"]
    orb, emb = psiformer_orbitals(params, electrons, atoms, nspins, layer_norm_mode, envelope_type)
    sign, lp = logdet_sum(orb)
    if "jastrow_layer" in p:
        lp = lp + simple_ee_jastrow(p["jastrow_layer"], emb["r_ee"], nspins)
    return sign, lp


# ---------------------------------------------------------------------------------------
# periodic FermiNet (solid)
# ---------------------------------------------------------------------------------------
_SYM_MATS = {
    "minimal": [[1, 0, 0], [0, 1, 0], [0, 0, 1]],
    "fcc": [[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]],
    "bcc": [[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, -1, 0], [1, 0, -1], [0, 1, -1]],
    "hexagonal": [[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, -1, 0]],
}


def get_symmetry_lat(lattice: torch.Tensor, sym_type: str = "minimal"):
    """``geometry/pbc.py:347-381``: reciprocal vectors expanded by the integer direction table of the symmetry type
    (unknown types fall back to the identity there; here they raise), ``av = pinv(bv)^T``."""
    mat = torch.tensor(_SYM_MATS[sym_type], dtype=F64)
    bv = mat @ (2 * math.pi * torch.linalg.inv(lattice).T)
    av = torch.linalg.pinv(bv).T
    return av, bv


def wrap_positions(pos, lattice: torch.Tensor):
    """``geometry/pbc.py:97-111``.  ``% 1.0`` is piecewise-linear with unit slope: under derivative
    tracking it is a constant integer shift of the fractional coordinate."""
    inv = torch.linalg.inv(lattice)
    frac = L.matmul(pos, inv)
    shift = torch.floor(L.value(frac)).detach()
    return L.matmul(frac - shift, lattice)


def tri_distance(xea, a: torch.Tensor, b: torch.Tensor):
    """``geometry/pbc.py:282-324``."""
    w = L.matmul(xea, b.T)  # (..., l)
    sg, cg = L.sin(w), L.cos(w)
    rel = L.cat([L.matmul(sg, a), L.matmul(cg, a)], dim=-1)
    metric = a @ a.T
    un_r = lambda t: L.linear_map(lambda u: u[..., :, None], t)  # noqa: E731
    un_c = lambda t: L.linear_map(lambda u: u[..., None, :], t)  # noqa: E731
    omc = 1.0 - cg
    vec = un_r(sg) * un_c(sg) + un_r(omc) * un_c(omc)
    sd = L.sqrt(L.sum_(L.sum_(vec * metric, dim=-1), dim=-1))
    return sd, rel


def _scaled_f(w):
    """``geometry/pbc.py:205-220``: ``f(w) = |w| (1 - |w / pi|^3 / 4)``."""
    aw = L.abs_(w)
    z = aw * (1.0 / math.pi)
    return aw * (1.0 - z * z * z * 0.25)


def _scaled_g(w):
    """``geometry/pbc.py:223-240``: ``g(w) = w (1 - 3/2 |w / pi| + 1/2 |w / pi|^2)``."""
    z = L.abs_(w) * (1.0 / math.pi)
    return w * (1.0 - z * 1.5 + z * z * 0.5)


def nu_distance(xea, a: torch.Tensor, b: torch.Tensor):
    """``geometry/pbc.py:243-279``: polynomial periodic distance; ``(w + pi) // (2 pi)`` is a constant integer shift
    under derivative tracking."""
    w = L.matmul(xea, b.T)
    shift = torch.floor((L.value(w) + math.pi) / (2 * math.pi)).detach()
    w = w - shift * (2 * math.pi)
    fw = _scaled_f(w) * torch.linalg.norm(a, dim=-1)
    r1 = L.sum_(fw * fw, dim=-1)
    sg = _scaled_g(w)
    rel = L.matmul(sg, a)
    metric = a @ a.T
    off = metric * (1.0 - torch.eye(a.shape[0], dtype=F64))
    un_r = lambda t: L.linear_map(lambda u: u[..., :, None], t)  # noqa: E731
    un_c = lambda t: L.linear_map(lambda u: u[..., None, :], t)  # noqa: E731
    r2 = L.sum_(L.sum_(un_r(sg) * un_c(sg) * off, dim=-1), dim=-1)
    return L.sqrt(r1 + r2), rel


def solid_features(electrons, prim_atoms, sim_lattice, prim_lattice, distance_type="tri", sym_type="minimal"):
    """``wavefunction/input/atomic.py:108-147``; ``distance_type`` 'tri' | 'nu' (``geometry/pbc.py:327-344``)."""
    dist = {"tri": tri_distance, "nu": nu_distance}[distance_type]
    sim_av, sim_bv = get_symmetry_lat(sim_lattice, sym_type)
    prim_av, prim_bv = get_symmetry_lat(prim_lattice, sym_type)
    n = electrons.shape[0]
    pe = wrap_positions(electrons, prim_lattice)
    ae_disp = L.linear_map(lambda t: t[:, None, :], pe) - prim_atoms
    r_ae, ae_vec = dist(ae_disp, prim_av, prim_bv)
    se = wrap_positions(electrons, sim_lattice)
    ee_disp = L.linear_map(lambda t: t[:, None, :], se) - L.linear_map(lambda t: t[None, :, :], se)
    eye = torch.eye(n, dtype=F64)
    r_ee, ee_vec = dist(ee_disp + eye[..., None], sim_av, sim_bv)
    r_ee = r_ee * (1.0 - eye)
    ee_vec = ee_vec * (1.0 - eye)[..., None]
    un = lambda t: L.linear_map(lambda u: u[..., None], t)  # noqa: E731
    ae_features = L.reshape(L.cat([un(r_ae), ae_vec], dim=-1), n, -1)
    ee_features = L.cat([un(r_ee), ee_vec], dim=-1)
    return dict(ae_features=ae_features, ee_features=ee_features, r_ae=r_ae, ae_vec=ae_vec)


def solid_logpsi(params, electrons, prim_atoms, nspins, sim_lattice, prim_lattice, klist, distance_type="tri",
                 sym_type="minimal"):
    """``app/solid/wavefunction.py:91-147`` -> complex ``logpsi``."""
    p = params["params"]
    n_layers = (len(p["backbone_layer"]) + 1) // 2
    emb = solid_features(electrons, prim_atoms, sim_lattice, prim_lattice, distance_type, sym_type)
    h_one, _ = fermi_layers(p["backbone_layer"], emb["ae_features"], emb["ee_features"], nspins, n_layers)
    orb_r = orbital_projection(p["real_orbital_layer"], h_one, nspins)
    orb_i = orbital_projection(p["imag_orbital_layer"], h_one, nspins)
    orb = L.to_complex(orb_r) + L.to_complex(orb_i) * 1j
    env = envelope(p["envelope_layer"], emb["r_ae"], nspins)
    orb = orb * L.to_complex(env)
    phase = L.exp(L.to_complex(L.matmul(electrons, klist.T)) * 1j)  # (n, n_orb)
    orb = orb * phase
    _, lp = logdet_sum(orb)
    return lp


# ---------------------------------------------------------------------------------------
# hydrogen atom demo wavefunction
# ---------------------------------------------------------------------------------------
def hydrogen_logpsi(alpha, electrons):
    """``app/hydrogen_atom.py:28-35``: ``log psi = alpha * |r|`` (one electron)."""
    r = _norm_last(electrons)
    return L.sum_(r * alpha, dim=0)


# ---------------------------------------------------------------------------------------
# parameter construction (Flax tree layout)
# ---------------------------------------------------------------------------------------
def _dense(g, fan_in, fan_out, bias="zeros", dtype=F64):
    d = {"kernel": torch.randn(fan_in, fan_out, generator=g, dtype=dtype) / math.sqrt(fan_in)}
    if bias == "zeros":
        d["bias"] = torch.zeros(fan_out, dtype=dtype)
    elif bias == "normal":
        d["bias"] = torch.randn(fan_out, generator=g, dtype=dtype)
    elif bias == "small":
        d["bias"] = 0.1 * torch.randn(fan_out, generator=g, dtype=dtype)
    return d


def _orbital_params(g, hidden, ndets, nspins, bias=False):
    n = sum(nspins)
    active = [s for s in nspins if s > 0]

    def dg():
        d = {"kernel": torch.randn(hidden, ndets, n, generator=g, dtype=F64) / math.sqrt(hidden)}
        if bias:
            d["bias"] = 0.1 * torch.randn(ndets, n, generator=g, dtype=F64)
        return d

    if len(active) > 1:
        return {"SplitChannelDense_0": {"DenseGeneral_0": dg(), "DenseGeneral_1": dg()}}
    return {"DenseGeneral_0": dg()}


def _envelope_params(g, n_atoms, ndets, nspins, split, jitter):
    n = sum(nspins)

    def env():
        pi = torch.ones(n, n_atoms, ndets, dtype=F64)
        sigma = torch.ones(n, n_atoms, ndets, dtype=F64)
        if jitter:
            pi = pi + jitter * torch.randn(pi.shape, generator=g, dtype=F64)
            sigma = sigma + jitter * torch.randn(sigma.shape, generator=g, dtype=F64)
        return {"pi": pi, "sigma": sigma}

    if split and nspins[0] > 0 and nspins[1] > 0:
        return {"_env_up": env(), "_env_down": env()}
    return {"_env": env()}


def init_ferminet_params(nspins, n_atoms, ndets=16, hidden_single=(256,) * 4, hidden_double=(32,) * 4,
                         seed=0, jitter=0.2, feat_per_atom=4, ee_feat=4, bias="small"):
    """Seeded parameters in the tree of ``app/molecule/wavefunction/ferminet.py:53-74``.

    Kernels are LeCun-normal like flax's default; biases / envelope parameters are jittered away from
    the flax defaults (0 / 1) so that parity tests exercise every term.
    """
    g = torch.Generator().manual_seed(seed)
    nch = sum(1 for s in nspins if s > 0)
    bb = {}
    d1, d2 = feat_per_atom * n_atoms, ee_feat
    idx = 0
    nl = len(hidden_single)
    for layer in range(nl):
        fan_in = d1 * (1 + nch) + d2 * nch
        bb[f"Dense_{idx}"] = _dense(g, fan_in, hidden_single[layer], bias)
        idx += 1
        if layer < nl - 1:
            bb[f"Dense_{idx}"] = _dense(g, d2, hidden_double[layer], bias)
            idx += 1
            d2 = hidden_double[layer]
        d1 = hidden_single[layer]
    return {"params": {
        "backbone_layer": bb,
        "orbital_layer": _orbital_params(g, d1, ndets, nspins),
        "envelope_layer": _envelope_params(g, n_atoms, ndets, nspins, True, jitter),
    }}


def init_solid_params(nspins, n_prim_atoms, ndets=16, hidden_single=(256,) * 4, hidden_double=(32,) * 4,
                      seed=0, jitter=0.2, distance_type="tri"):
    """Tree of ``app/solid/wavefunction.py:58-89`` (``tri`` features: 7 per atom / pair; ``nu``: 4)."""
    fw = 4 if distance_type == "nu" else 7
    base = init_ferminet_params(nspins, n_prim_atoms, ndets, hidden_single, hidden_double, seed, jitter,
                                feat_per_atom=fw, ee_feat=fw)["params"]
    g = torch.Generator().manual_seed(seed + 7919)
    h = hidden_single[-1]
    return {"params": {
        "backbone_layer": base["backbone_layer"],
        "real_orbital_layer": base["orbital_layer"],
        "imag_orbital_layer": _orbital_params(g, h, ndets, nspins),
        "envelope_layer": base["envelope_layer"],
    }}


def init_lapnet_params(nspins, n_atoms, ndets=16, num_layers=4, heads=4, heads_dim=64, num_local_updates=2,
                       seed=0, jitter=0.2, jastrow=True):
    """Tree of ``backbone/lapnet/_backbone.py:40-64,169-189`` (biases N(0,1) as in ``:44``)."""
    g = torch.Generator().manual_seed(seed)
    d = heads * heads_dim
    bb = {"input_projection": _dense(g, 4 * n_atoms + 1, d, "normal")}
    for li in range(num_layers):
        lp = {
            "qk_projection": _dense(g, d, 2 * d, "normal"),
            "value_projection": _dense(g, d, d, "normal"),
            "output_projection": _dense(g, d, d, "normal"),
            "value_update": _dense(g, d, d, "normal"),
        }
        nlu = num_local_updates if li < num_layers - 1 else 0
        for j in range(nlu):
            lp[f"qk_update_layers_{j}"] = _dense(g, d, d, "normal")
        bb[f"layers_{li}"] = lp
    p = {
        "backbone_layer": bb,
        "orbital_layer": _orbital_params(g, d, ndets, nspins),
        "envelope_layer": _envelope_params(g, n_atoms, ndets, nspins, True, jitter),
    }
    if jastrow:
        p["jastrow_layer"] = {"alpha_par": torch.tensor([1.0 + 0.3 * jitter], dtype=F64),
                              "alpha_anti": torch.tensor([1.0 - 0.4 * jitter], dtype=F64)}
    return {"params": p}


def init_psiformer_params(nspins, n_atoms, ndets=16, num_layers=4, heads=4, heads_dim=64, mlp_hidden=(256,),
                          seed=0, jitter=0.2, jastrow=True):
    """Tree of ``backbone/psiformer.py:60-99,175-185``."""
    g = torch.Generator().manual_seed(seed)
    d = heads * heads_dim
    bb = {"Dense_0": _dense(g, 4 * n_atoms + 1, d, "small")}

    def ln():
        return {"scale": 1.0 + jitter * torch.randn(d, generator=g, dtype=F64),
                "bias": jitter * torch.randn(d, generator=g, dtype=F64)}

    for li in range(num_layers):
        def proj():
            return {"kernel": torch.randn(d, heads, heads_dim, generator=g, dtype=F64) / math.sqrt(d),
                    "bias": 0.1 * torch.randn(heads, heads_dim, generator=g, dtype=F64)}
        lp = {
            "LayerNorm_0": ln(),
            "MultiHeadDotProductAttention_0": {
                "query": proj(), "key": proj(), "value": proj(),
                "out": {"kernel": torch.randn(heads, heads_dim, d, generator=g, dtype=F64) / math.sqrt(d),
                        "bias": 0.1 * torch.randn(d, generator=g, dtype=F64)},
            },
            "LayerNorm_1": ln(),
        }
        fan = d
        j = 0
        for hdim in mlp_hidden:
            lp[f"Dense_{j}"] = _dense(g, fan, hdim, "small")
            fan = hdim
            j += 1
        lp[f"Dense_{j}"] = _dense(g, fan, d, "small")
        bb[f"PsiformerLayer_{li}"] = lp
    p = {
        "backbone_layer": bb,
        "orbital_layer": _orbital_params(g, d, ndets, nspins),
        "envelope_layer": _envelope_params(g, n_atoms, ndets, nspins, True, jitter),
    }
    if jastrow:
        p["jastrow_layer"] = {"alpha_par": torch.tensor([1.0 + 0.3 * jitter], dtype=F64),
                              "alpha_anti": torch.tensor([1.0 - 0.4 * jitter], dtype=F64)}
    return {"params": p}


def tree_map(fn, tree):
    if isinstance(tree, dict):
        return {k: tree_map(fn, v) for k, v in tree.items()}
    return fn(tree)


def tree_leaves(tree):
    """Leaves in ``jax.tree.leaves`` order (dict keys sorted)."""
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree):
            out.extend(tree_leaves(tree[k]))
        return out
    return [tree]
