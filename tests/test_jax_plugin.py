"""What can be checked of the JAX-side boundary without jax / jaxlib (neither is installable in the build container):
the XLA-FFI shim type-checks against a declaration-only stand-in of xla/ffi/api/ffi.h, the plugin's Python files
byte-compile, and the attribute packing of ``jaqmc_b200_jax._config`` reproduces the ctypes structs of the product's own
marshaller word for word (the shim memcpy's those words into the C structs)."""

import ctypes as C
import os
import py_compile
import subprocess

import numpy as np
import pytest
import torch

import helpers as H
from jaqmc_b200 import _abi
from jaqmc_b200 import _marshal as M
from oracle import networks as ON


def test_xla_ffi_shim_type_checks():
    r = subprocess.run(["make", "-C", os.path.join(H.ROOT, "ffi"), "check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "error" not in r.stderr.lower(), r.stderr


def test_plugin_files_compile():
    d = os.path.join(H.ROOT, "jaqmc_b200_jax")
    for f in sorted(os.listdir(d)):
        if f.endswith(".py"):
            py_compile.compile(os.path.join(d, f), doraise=True)


def _words(struct, n_int_bytes):
    return np.frombuffer(bytes(struct)[:n_int_bytes], dtype=np.int32)


def test_config_packing_matches_the_c_structs():
    from jaqmc_b200_jax._config import pack_config

    atoms, charges, nspins = H.molecule("LiH")
    hs, hd = (32, 32, 32), (8, 8, 8)
    p = H.to_f32(ON.init_ferminet_params(nspins, 2, 4, hs, hd, seed=1))
    h = M.ferminet_handle(p, nspins, 2, 4, hs, hd, "isotropic", True, False)
    cfg = C.cast(h.struct.config, C.POINTER(_abi.FerminetConfig)).contents
    kind, words, fcfg, opt = pack_config("ferminet", nspins=nspins, n_atoms=2, ndets=4, hidden_dims_single=hs,
                                         hidden_dims_double=hd, envelope="isotropic")
    assert kind == _abi.WF_FERMINET and np.array_equal(words, _words(cfg, C.sizeof(cfg)))
    p = H.to_f32(ON.init_lapnet_params(nspins, 2, 4, 3, 2, 8, 2, seed=3))
    h = M.lapnet_handle(p, nspins, 2, 4, 3, 2, 8, 2)
    cfg = C.cast(h.struct.config, C.POINTER(_abi.LapnetConfig)).contents
    kind, words, fcfg, opt = pack_config("lapnet", nspins=nspins, n_atoms=2, ndets=4, num_layers=3, num_heads=2, heads_dim=8)
    assert kind == _abi.WF_LAPNET and np.array_equal(words, _words(cfg, C.sizeof(cfg))) and opt == 1 | 2 | 8
    p = H.to_f32(ON.init_psiformer_params(nspins, 2, 4, 2, 2, 8, (16, 24), seed=5))
    h = M.psiformer_handle(p, nspins, 2, 4, 2, 2, 8, (16, 24), "post")
    cfg = C.cast(h.struct.config, C.POINTER(_abi.PsiformerConfig)).contents
    kind, words, fcfg, opt = pack_config("psiformer", nspins=nspins, n_atoms=2, ndets=4, num_layers=2, num_heads=2,
                                         heads_dim=8, mlp_hidden_dims=(16, 24), layer_norm_mode="post")
    assert kind == _abi.WF_PSIFORMER and np.array_equal(words, _words(cfg, C.sizeof(cfg)))
    prim, sim, patoms, cell_atoms, cell_charges, sn, klist = H.solid_system("fcc_lih_221")
    p = H.to_f32(ON.init_solid_params(sn, 2, 2, (16, 16), (8, 8), seed=2))
    p = H.to_f32(ON.init_solid_params(sn, 2, 2, (16, 16), (8, 8), seed=2, distance_type="nu"))
    h = M.solid_handle(p, sn, 2, sim, prim, torch.as_tensor(klist, dtype=torch.float32), 2, (16, 16), (8, 8),
                       distance_type="nu", sym_type="bcc")
    cfg = C.cast(h.struct.config, C.POINTER(_abi.SolidConfig)).contents
    kind, words, fcfg, opt = pack_config("solid", nspins=sn, n_atoms=2, ndets=2, hidden_dims_single=(16, 16),
                                         hidden_dims_double=(8, 8), simulation_lattice=sim, primitive_lattice=prim,
                                         distance_type="nu", sym_type="bcc")
    # the shim's `config` attribute: the embedded FermiNet config words, then distance_type and sym_type; the two
    # lattices travel as 18 floats in `fconfig`
    n_int = C.sizeof(_abi.FerminetConfig)
    assert kind == _abi.WF_SOLID_FERMINET and np.array_equal(words[:-2], _words(cfg, n_int))
    assert list(words[-2:]) == [cfg.distance_type, cfg.sym_type] == [1, 2]
    assert np.array_equal(fcfg, np.frombuffer(bytes(cfg)[n_int:n_int + 72], dtype=np.float32))
