"""``jaqmc_b200_bind_param_leaves`` (the leaves -> descriptor mapping an XLA-FFI shim uses, csrc/leaves.cu) against
the ctypes marshaller (``jaqmc_b200/_marshal.py``) on trees flattened in ``jax.tree.leaves`` order (dict keys sorted at
every level), including the reference-generated trees of tests/golden/ref_*.npz.  No GPU needed: host-only code of the
CUDA library."""

import ctypes as C
import os

import pytest
import torch

import helpers as H
import test_reference_fixtures as R
from jaqmc_b200 import _abi
from jaqmc_b200 import _marshal as M
from oracle import networks as ON

LIB = os.path.join(H.ROOT, "jaqmc_b200", "_C", "libjaqmc_b200.so")
PRESENT = C.c_void_p(1)


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.skip("CUDA library not built")
    return _abi.bind(C.CDLL(LIB))


def flat_sorted(tree, prefix="params"):
    """(path, tensor) in jax.tree.leaves order."""
    out = []
    for k in sorted(tree):
        v = tree[k]
        out.extend(flat_sorted(v, f"{prefix}/{k}") if isinstance(v, dict) else [(f"{prefix}/{k}", v)])
    return out


def pointer_fields(struct):
    """{dotted field name: integer pointer value} of a ctypes params struct (arrays and nested structs unrolled)."""
    out = {}

    def walk(obj, prefix):
        for name, typ in obj._fields_:
            v = getattr(obj, name)
            if isinstance(v, C.Structure):
                walk(v, f"{prefix}{name}.")
            elif isinstance(v, C.Array):
                for i, e in enumerate(v):
                    if isinstance(e, C.Array):
                        for j, f in enumerate(e):
                            out[f"{prefix}{name}[{i}][{j}]"] = f or 0
                    else:
                        out[f"{prefix}{name}[{i}]"] = e or 0
            else:
                out[f"{prefix}{name}"] = v or 0

    walk(struct, "")
    return out


def check(lib, handle, params_tree, preset):
    """Bind the sorted leaves of ``params_tree`` into a fresh params struct and compare with the marshaller's."""
    wf = handle.struct
    cfg_ptr = wf.config
    ptype = type(C.cast(wf.params, C.POINTER(_params_type(wf.kind))).contents)
    fresh = ptype()
    preset(fresh)
    leaves = flat_sorted(params_tree["params"] if "params" in params_tree else params_tree)
    n = lib.jaqmc_b200_param_leaf_count(wf.kind, cfg_ptr, C.byref(fresh))
    assert n == len(leaves), (n, [p for p, _ in leaves])
    path = C.create_string_buffer(512)
    nel, rank, dims = C.c_int64(), C.c_int32(), (C.c_int64 * 4)()
    for i, (pth, t) in enumerate(leaves):
        _abi.check(lib, lib.jaqmc_b200_param_leaf_info(wf.kind, cfg_ptr, C.byref(fresh), i, path, 512, C.byref(nel),
                                                        C.byref(rank), dims))
        assert path.value.decode() == pth, (i, path.value.decode(), pth)
        assert nel.value == t.numel() and tuple(dims[:rank.value]) == tuple(t.shape), (pth, tuple(dims[:rank.value]), t.shape)
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for _, t in leaves])
    sizes = (C.c_int64 * n)(*[t.numel() for _, t in leaves])
    _abi.check(lib, lib.jaqmc_b200_bind_param_leaves(wf.kind, cfg_ptr, C.byref(fresh), ptrs, sizes, n))
    want = pointer_fields(C.cast(wf.params, C.POINTER(ptype)).contents)
    got = pointer_fields(fresh)
    want.pop("klist", None)
    got.pop("klist", None)
    assert got == want
    # a wrong leaf count / size is refused with a message naming the leaf
    sizes[0] += 1
    fresh2 = ptype()
    preset(fresh2)
    rc = lib.jaqmc_b200_bind_param_leaves(wf.kind, cfg_ptr, C.byref(fresh2), ptrs, sizes, n)
    assert rc == _abi.ERR_INVALID_ARGUMENT and b"leaf 0" in lib.jaqmc_b200_last_error()
    assert lib.jaqmc_b200_bind_param_leaves(wf.kind, cfg_ptr, C.byref(fresh), ptrs, None, n - 1) == _abi.ERR_INVALID_ARGUMENT


def _params_type(kind):
    return {_abi.WF_FERMINET: _abi.FerminetParams, _abi.WF_LAPNET: _abi.LapnetParams, _abi.WF_PSIFORMER: _abi.PsiformerParams,
            _abi.WF_SOLID_FERMINET: _abi.SolidParams, _abi.WF_HYDROGEN: _abi.HydrogenParams}[kind]


def test_ferminet_leaf_order_including_env_down_before_env_up(lib):
    atoms, charges, nspins = H.molecule("N2")
    hs, hd = (32,) * 6, (8,) * 6            # 11 Dense layers: "Dense_10" sorts before "Dense_2"
    p = H.to_f32(ON.init_ferminet_params(nspins, 2, 4, hs, hd, seed=1))
    h = M.ferminet_handle(p, nspins, 2, 4, hs, hd)
    paths = [q for q, _ in flat_sorted(p["params"])]
    assert paths.index("params/envelope_layer/_env_down/pi") < paths.index("params/envelope_layer/_env_up/pi")
    assert paths.index("params/backbone_layer/Dense_10/bias") < paths.index("params/backbone_layer/Dense_2/bias")
    check(lib, h, p, lambda s: None)


def test_lapnet_psiformer_solid_hydrogen_leaf_order(lib):
    atoms, charges, nspins = H.molecule("LiH")
    p = H.to_f32(ON.init_lapnet_params(nspins, 2, 4, 3, 2, 8, 2, seed=3))
    h = M.lapnet_handle(p, nspins, 2, 4, 3, 2, 8, 2)

    def lap_preset(s):
        s.input_bias = PRESENT
        for l in range(3):
            s.qk_bias[l] = PRESENT
        s.head.jastrow_alpha_par = PRESENT

    check(lib, h, p, lap_preset)
    p = H.to_f32(ON.init_psiformer_params(nspins, 2, 4, 2, 2, 8, (16, 24), seed=5))
    h = M.psiformer_handle(p, nspins, 2, 4, 2, 2, 8, (16, 24))

    def psi_preset(s):
        s.input_bias = PRESENT
        for l in range(2):
            s.q_bias[l] = PRESENT
        s.head.jastrow_alpha_par = PRESENT
        if "bias" in str(flat_sorted(p["params"]["orbital_layer"])):
            s.head.orbital_bias[0] = PRESENT

    check(lib, h, p, psi_preset)
    prim, sim, patoms, cell_atoms, cell_charges, sn, klist = H.solid_system("fcc_lih_221")
    p = H.to_f32(ON.init_solid_params(sn, 2, 2, (16, 16), (8, 8), seed=2))
    h = M.solid_handle(p, sn, 2, sim, prim, torch.as_tensor(klist, dtype=torch.float32), 2, (16, 16), (8, 8))
    check(lib, h, p, lambda s: None)
    p = {"params": {"alpha": torch.tensor([-0.8])}}
    check(lib, M.hydrogen_handle(p), p, lambda s: None)


@pytest.mark.parametrize("name", ["ferminet_lih_last_layer", "ferminet_lih_diagonal", "ferminet_lih_null",
                                  "ferminet_lih_nosplit", "ferminet_h_single_channel", "lapnet_lih_layernorm",
                                  "lapnet_li_nojastrow", "psiformer_he_null", "psiformer_lih_post"])
def test_reference_generated_trees(lib, name):
    """Trees produced by the reference's own ``init_params`` (tests/golden/ref_*.npz)."""
    import test_emu_reference_fixtures as E

    meta, p64, z = R.load(name)
    params = H.to_f32(p64)
    wf, _, _ = E.handles(meta, {"params": params["params"]} if "params" in params else params, z)
    tree = params["params"] if "params" in params else params
    paths = [q for q, _ in flat_sorted(tree)]

    def preset(s):
        if meta["kind"] == "lapnet":
            s.input_bias = PRESENT
            for l in range(meta["kwargs"]["num_layers"]):
                s.qk_bias[l] = PRESENT
            if any("jastrow_layer" in q for q in paths):
                s.head.jastrow_alpha_par = PRESENT
        if meta["kind"] == "psiformer":
            s.input_bias = PRESENT
            for l in range(meta["kwargs"]["num_layers"]):
                s.q_bias[l] = PRESENT
            s.head.jastrow_alpha_par = PRESENT
            if any(q.endswith("DenseGeneral_0/bias") for q in paths):
                s.head.orbital_bias[0] = PRESENT

    # the emu handles were built from a float32 copy: rebuild the marshalled handle on THIS tree for pointer equality
    kw = meta["kwargs"]
    nspins = tuple(meta["nspins"])
    A = z["atoms"].shape[0]
    env = kw.get("envelope", "abs_isotropic")
    if meta["kind"] == "ferminet":
        h = M.ferminet_handle(params, nspins, A, kw["ndets"], kw["hidden_dims_single"], kw["hidden_dims_double"], env,
                              kw.get("orbitals_spin_split", True), kw.get("use_last_layer", False))
    elif meta["kind"] == "lapnet":
        h = M.lapnet_handle(params, nspins, A, kw["ndets"], kw["num_layers"], kw["num_heads"], kw["heads_dim"],
                            kw.get("num_local_updates", 2), env, True, kw.get("jastrow", "simple_ee") == "simple_ee",
                            kw.get("use_layernorm", False))
    else:
        h = M.psiformer_handle(params, nspins, A, kw["ndets"], kw["num_layers"], kw["num_heads"], kw["heads_dim"],
                               tuple(kw["mlp_hidden_dims"]), kw.get("layer_norm_mode", "pre"), env, True, True,
                               kw.get("jastrow", "simple_ee") == "simple_ee")
    check(lib, h, params, preset)
