"""ctypes mirror of ``include/jaqmc_b200.h`` (structures + prototypes).

``bind(cdll)`` attaches argument / result types to a loaded library.  The product loads the CUDA build
through ``jaqmc_b200._lib``; the CPU test-suite binds the host-emulation build (``tests/emu``) with the same
prototypes so the two cannot drift apart.
"""

from __future__ import annotations

import ctypes as C

MAX_LAYERS = 8

OK = 0
ERR_INVALID_ARGUMENT = 1
ERR_WORKSPACE_TOO_SMALL = 2
ERR_CUDA = 3
ERR_UNSUPPORTED = 4

ENVELOPE = {"isotropic": 0, "abs_isotropic": 1, "null": 2, "diagonal": 3}

WF_FERMINET = 1
WF_LAPNET = 2
WF_PSIFORMER = 3
WF_SOLID_FERMINET = 4
WF_HYDROGEN = 5

FloatP = C.c_void_p  # device (or, for the emulation build, host) pointer to float32


class FerminetConfig(C.Structure):
    _fields_ = [
        ("n_up", C.c_int32),
        ("n_dn", C.c_int32),
        ("n_atoms", C.c_int32),
        ("ndets", C.c_int32),
        ("n_layers", C.c_int32),
        ("hidden_single", C.c_int32 * MAX_LAYERS),
        ("hidden_double", C.c_int32 * MAX_LAYERS),
        ("envelope_type", C.c_int32),
        ("orbitals_spin_split", C.c_int32),
        ("use_last_layer", C.c_int32),
    ]


class FerminetParams(C.Structure):
    _fields_ = [
        ("single_kernel", FloatP * MAX_LAYERS),
        ("single_bias", FloatP * MAX_LAYERS),
        ("double_kernel", FloatP * MAX_LAYERS),
        ("double_bias", FloatP * MAX_LAYERS),
        ("orbital_kernel", FloatP * 2),
        ("env_pi", FloatP * 2),
        ("env_sigma", FloatP * 2),
    ]


class HeadParams(C.Structure):
    _fields_ = [
        ("orbital_kernel", FloatP * 2),
        ("orbital_bias", FloatP * 2),
        ("env_pi", FloatP * 2),
        ("env_sigma", FloatP * 2),
        ("jastrow_alpha_par", FloatP),
        ("jastrow_alpha_anti", FloatP),
    ]


class LapnetConfig(C.Structure):
    _fields_ = [
        ("n_up", C.c_int32),
        ("n_dn", C.c_int32),
        ("n_atoms", C.c_int32),
        ("ndets", C.c_int32),
        ("num_layers", C.c_int32),
        ("num_heads", C.c_int32),
        ("heads_dim", C.c_int32),
        ("num_local_updates", C.c_int32),
        ("envelope_type", C.c_int32),
        ("rescale", C.c_int32),
        ("use_layernorm", C.c_int32),
    ]


class LapnetParams(C.Structure):
    _fields_ = [
        ("input_kernel", FloatP),
        ("input_bias", FloatP),
        ("qk_kernel", FloatP * MAX_LAYERS),
        ("qk_bias", FloatP * MAX_LAYERS),
        ("value_kernel", FloatP * MAX_LAYERS),
        ("value_bias", FloatP * MAX_LAYERS),
        ("output_kernel", FloatP * MAX_LAYERS),
        ("output_bias", FloatP * MAX_LAYERS),
        ("update_kernel", FloatP * MAX_LAYERS),
        ("update_bias", FloatP * MAX_LAYERS),
        ("qk_update_kernel", (FloatP * 4) * MAX_LAYERS),
        ("qk_update_bias", (FloatP * 4) * MAX_LAYERS),
        ("qk_ln_scale", FloatP * MAX_LAYERS),
        ("qk_ln_bias", FloatP * MAX_LAYERS),
        ("value_ln_scale", FloatP * MAX_LAYERS),
        ("value_ln_bias", FloatP * MAX_LAYERS),
        ("post_ln_scale", FloatP * MAX_LAYERS),
        ("post_ln_bias", FloatP * MAX_LAYERS),
        ("head", HeadParams),
    ]


LAYERNORM = {"pre": 0, "post": 1, "null": 2}
MAX_MLP = 4


class PsiformerConfig(C.Structure):
    _fields_ = [
        ("n_up", C.c_int32),
        ("n_dn", C.c_int32),
        ("n_atoms", C.c_int32),
        ("ndets", C.c_int32),
        ("num_layers", C.c_int32),
        ("num_heads", C.c_int32),
        ("heads_dim", C.c_int32),
        ("n_mlp_hidden", C.c_int32),
        ("mlp_hidden", C.c_int32 * MAX_MLP),
        ("layer_norm_mode", C.c_int32),
        ("envelope_type", C.c_int32),
        ("orbitals_spin_split", C.c_int32),
        ("rescale", C.c_int32),
    ]


class PsiformerParams(C.Structure):
    _fields_ = [
        ("input_kernel", FloatP),
        ("input_bias", FloatP),
        ("ln0_scale", FloatP * MAX_LAYERS),
        ("ln0_bias", FloatP * MAX_LAYERS),
        ("q_kernel", FloatP * MAX_LAYERS),
        ("q_bias", FloatP * MAX_LAYERS),
        ("k_kernel", FloatP * MAX_LAYERS),
        ("k_bias", FloatP * MAX_LAYERS),
        ("v_kernel", FloatP * MAX_LAYERS),
        ("v_bias", FloatP * MAX_LAYERS),
        ("out_kernel", FloatP * MAX_LAYERS),
        ("out_bias", FloatP * MAX_LAYERS),
        ("ln1_scale", FloatP * MAX_LAYERS),
        ("ln1_bias", FloatP * MAX_LAYERS),
        ("mlp_kernel", (FloatP * MAX_MLP) * MAX_LAYERS),
        ("mlp_bias", (FloatP * MAX_MLP) * MAX_LAYERS),
        ("head", HeadParams),
    ]


DISTANCE = {"tri": 0, "nu": 1}                                   # geometry/pbc.py DistanceType
SYMMETRY = {"minimal": 0, "fcc": 1, "bcc": 2, "hexagonal": 3}     # geometry/pbc.py SymmetryType


class SolidConfig(C.Structure):
    _fields_ = [("net", FerminetConfig), ("simulation_lattice", C.c_float * 9), ("primitive_lattice", C.c_float * 9),
                ("distance_type", C.c_int32), ("sym_type", C.c_int32)]


class SolidParams(C.Structure):
    _fields_ = [
        ("net", FerminetParams),
        ("real_orbital_kernel", FloatP * 2),
        ("imag_orbital_kernel", FloatP * 2),
        ("klist", FloatP),
    ]


class HydrogenConfig(C.Structure):
    _fields_ = [("n_electrons", C.c_int32)]


class HydrogenParams(C.Structure):
    _fields_ = [("alpha", FloatP)]


class Wavefunction(C.Structure):
    _fields_ = [("kind", C.c_int32), ("config", C.c_void_p), ("params", C.c_void_p)]


class System(C.Structure):
    _fields_ = [("atoms", FloatP), ("charges", FloatP), ("n_atoms", C.c_int32)]


class Ewald(C.Structure):
    _fields_ = [
        ("lattice", FloatP),
        ("inv_lattice", FloatP),
        ("mic_shifts", FloatP),
        ("images", FloatP),
        ("gpoints", FloatP),
        ("gweight", FloatP),
        ("n_images", C.c_int32),
        ("center_image", C.c_int32),
        ("n_g", C.c_int32),
        ("mic_kind", C.c_int32),
        ("alpha", C.c_float),
        ("self_const_factor", C.c_float),
        ("ijconst", C.c_float),
    ]


PROTOTYPES = {
    "jaqmc_b200_local_energy_complex": (
        C.c_int,
        [C.POINTER(Wavefunction), C.POINTER(System), C.c_void_p, FloatP, FloatP, C.c_int32, FloatP, C.c_int64, FloatP,
         FloatP, FloatP, FloatP, FloatP, FloatP, FloatP, C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "jaqmc_b200_ewald": (
        C.c_int,
        [C.POINTER(Ewald), FloatP, C.c_int64, C.c_int32, FloatP, FloatP, C.c_int32, FloatP, C.c_void_p],
    ),
    "jaqmc_b200_workspace_bytes": (C.c_size_t, [C.POINTER(Wavefunction), C.c_int64, C.c_int]),
    "jaqmc_b200_logpsi": (
        C.c_int,
        [C.POINTER(Wavefunction), C.POINTER(System), FloatP, C.c_int64, FloatP, FloatP, C.c_void_p, C.c_size_t,
         C.c_void_p],
    ),
    "jaqmc_b200_local_energy": (
        C.c_int,
        [C.POINTER(Wavefunction), C.POINTER(System), FloatP, C.c_int64, FloatP, FloatP, FloatP, FloatP, FloatP,
         FloatP, FloatP, FloatP, C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "jaqmc_b200_coulomb": (C.c_int, [C.POINTER(System), FloatP, C.c_int64, C.c_int32, FloatP, C.c_void_p]),
    "jaqmc_b200_mh_step": (
        C.c_int,
        [C.POINTER(Wavefunction), C.POINTER(System), FloatP, FloatP, C.c_int32, FloatP, FloatP, FloatP, C.c_int32,
         C.c_int64, FloatP, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "jaqmc_b200_mh_step_pbc": (
        C.c_int,
        [C.POINTER(Wavefunction), C.POINTER(System), FloatP, FloatP, C.c_int32, FloatP, FloatP, FloatP, C.c_int32,
         C.c_int64, FloatP, C.c_void_p, C.POINTER(C.c_float), C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "jaqmc_b200_orbitals": (
        C.c_int,
        [C.POINTER(Wavefunction), C.POINTER(System), FloatP, C.c_int64, FloatP, C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "jaqmc_b200_psi_ratios": (
        C.c_int,
        [C.POINTER(Wavefunction), C.POINTER(System), FloatP, C.c_int64, C.c_int32, C.c_void_p, FloatP, FloatP, FloatP,
         C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "jaqmc_b200_ferminet_vjp_workspace_bytes": (C.c_size_t, [C.POINTER(FerminetConfig), C.c_int64]),
    "jaqmc_b200_ferminet_logpsi_vjp": (
        C.c_int,
        [C.POINTER(FerminetConfig), C.POINTER(FerminetParams), C.POINTER(System), FloatP, C.c_int64, FloatP,
         C.POINTER(FerminetParams), FloatP, FloatP, C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "jaqmc_b200_dense_fl": (
        C.c_int,
        [FloatP, FloatP, FloatP, FloatP, FloatP, FloatP, FloatP, FloatP, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
         C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "jaqmc_b200_layernorm_fl": (
        C.c_int,
        [FloatP, FloatP, FloatP, FloatP, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_void_p],
    ),
    "jaqmc_b200_attention_fl": (
        C.c_int,
        [FloatP, FloatP, FloatP, FloatP, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
         C.c_void_p],
    ),
    "jaqmc_b200_mh_propose": (C.c_int, [FloatP, FloatP, FloatP, FloatP, C.c_int64, C.c_void_p]),
    "jaqmc_b200_mh_accept": (
        C.c_int,
        [FloatP, FloatP, FloatP, FloatP, FloatP, C.c_int64, C.c_int32, FloatP, C.c_void_p, C.c_void_p],
    ),
    "jaqmc_b200_param_leaf_count": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p]),
    "jaqmc_b200_param_leaf_info": (
        C.c_int,
        [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_char_p, C.c_size_t, C.POINTER(C.c_int64),
         C.POINTER(C.c_int32), C.POINTER(C.c_int64)],
    ),
    "jaqmc_b200_bind_param_leaves": (
        C.c_int,
        [C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int32],
    ),
    "jaqmc_b200_launch_count": (C.c_int64, []),
    "jaqmc_b200_reset_launch_count": (None, []),
    "jaqmc_b200_profile_enable": (None, [C.c_int]),
    "jaqmc_b200_profile_fetch": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "jaqmc_b200_last_error": (C.c_char_p, []),
    "jaqmc_b200_version": (C.c_char_p, []),
}


def bind(cdll: C.CDLL) -> C.CDLL:
    """Attach prototypes; raises ``AttributeError`` if the library lacks a declared symbol."""
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(cdll, name)
        fn.restype = restype
        fn.argtypes = argtypes
    return cdll


class JaqmcB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"jaqmc_b200 error {code}: {message}")
        self.code = code


def check(cdll: C.CDLL, rc: int) -> None:
    if rc != OK:
        raise JaqmcB200Error(rc, cdll.jaqmc_b200_last_error().decode())
