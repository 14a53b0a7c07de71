"""Packing of the wavefunction configuration into the XLA-FFI attributes (``kind``, ``config`` int32 words, ``fconfig``
floats, ``optional`` bit mask) that ``ffi/xla_ffi_shim.cc`` unpacks into the structs of ``include/jaqmc_b200.h``.
Pure Python / numpy: importable (and tested) without jax."""
from __future__ import annotations

import numpy as np

MAX_LAYERS = 8
MAX_MLP = 4
ENVELOPE = {"isotropic": 0, "abs_isotropic": 1, "null": 2, "diagonal": 3}
LAYERNORM = {"pre": 0, "post": 1, "null": 2}
WF_FERMINET, WF_LAPNET, WF_PSIFORMER, WF_SOLID_FERMINET, WF_HYDROGEN = 1, 2, 3, 4, 5
OPT_INPUT_BIAS, OPT_BACKBONE_BIAS, OPT_ORBITAL_BIAS, OPT_JASTROW = 1, 2, 4, 8


def _pad(seq, n):
    seq = [int(v) for v in seq]
    if len(seq) > n:
        raise ValueError(f"at most {n} entries are supported, got {len(seq)}")
    return seq + [0] * (n - len(seq))


def pack_config(kind: str, **f):
    """``(kind_id, config int32[], fconfig float32[], optional_mask)`` -- field order = struct order of jaqmc_b200.h."""
    env = ENVELOPE[str(f.get("envelope", "abs_isotropic"))]
    n_up, n_dn = (int(v) for v in f["nspins"])
    if kind in ("ferminet", "solid"):
        hs, hd = list(f["hidden_dims_single"]), list(f["hidden_dims_double"])
        if len(hs) != len(hd):
            raise ValueError("hidden_dims_single and hidden_dims_double must have the same length")
        words = [n_up, n_dn, int(f["n_atoms"]), int(f["ndets"]), len(hs), *_pad(hs, MAX_LAYERS), *_pad(hd, MAX_LAYERS), env,
                 int(bool(f.get("orbitals_spin_split", True)) and n_up > 0 and n_dn > 0), int(bool(f.get("use_last_layer", False)))]
        if kind == "ferminet":
            return WF_FERMINET, np.asarray(words, np.int32), np.zeros(0, np.float32), 0
        lat = np.concatenate([np.asarray(f["simulation_lattice"], np.float32).reshape(9),
                              np.asarray(f["primitive_lattice"], np.float32).reshape(9)])
        # two trailing words: geometry/pbc.py DistanceType / SymmetryType (JAQMC_DISTANCE_*, JAQMC_SYMMETRY_*)
        words += [{"tri": 0, "nu": 1}[str(f.get("distance_type", "tri"))],
                  {"minimal": 0, "fcc": 1, "bcc": 2, "hexagonal": 3}[str(f.get("sym_type", "minimal"))]]
        return WF_SOLID_FERMINET, np.asarray(words, np.int32), lat, 0
    if kind == "lapnet":
        words = [n_up, n_dn, int(f["n_atoms"]), int(f["ndets"]), int(f["num_layers"]), int(f["num_heads"]), int(f["heads_dim"]),
                 int(f.get("num_local_updates", 2)), env, int(bool(f.get("rescale", True))), int(bool(f.get("use_layernorm", False)))]
        opt = (OPT_INPUT_BIAS * bool(f.get("use_input_bias", True)) | OPT_BACKBONE_BIAS * bool(f.get("use_backbone_bias", True))
               | OPT_ORBITAL_BIAS * bool(f.get("use_orbital_bias", False)) | OPT_JASTROW * (f.get("jastrow", "simple_ee") == "simple_ee"))
        return WF_LAPNET, np.asarray(words, np.int32), np.zeros(0, np.float32), int(opt)
    if kind == "psiformer":
        mlp = list(f.get("mlp_hidden_dims", (256,)))
        words = [n_up, n_dn, int(f["n_atoms"]), int(f["ndets"]), int(f["num_layers"]), int(f["num_heads"]), int(f["heads_dim"]),
                 len(mlp), *_pad(mlp, MAX_MLP), LAYERNORM[str(f.get("layer_norm_mode", "pre"))], env,
                 int(bool(f.get("orbitals_spin_split", True)) and n_up > 0 and n_dn > 0), int(bool(f.get("rescale", True)))]
        opt = (OPT_INPUT_BIAS * bool(f.get("input_bias", True)) | OPT_BACKBONE_BIAS * bool(f.get("with_bias", True))
               | OPT_ORBITAL_BIAS * bool(f.get("bias_orbitals", False)) | OPT_JASTROW * (f.get("jastrow", "simple_ee") == "simple_ee"))
        return WF_PSIFORMER, np.asarray(words, np.int32), np.zeros(0, np.float32), int(opt)
    if kind == "hydrogen":
        return WF_HYDROGEN, np.asarray([int(f.get("n_electrons", 1))], np.int32), np.zeros(0, np.float32), 0
    raise KeyError(kind)
