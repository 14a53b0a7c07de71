"""Registration of the XLA-FFI targets of ``libjaqmc_b200_ffi.so`` (built from ffi/xla_ffi_shim.cc) and thin
``jax.ffi.ffi_call`` wrappers.  Needs jax >= 0.5 (``jax.ffi``); written for the reference's pinned jax==0.9.1."""
from __future__ import annotations

import ctypes
import os

import jax
import jax.numpy as jnp
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("JAQMC_B200_FFI_LIB", os.path.join(_HERE, "..", "jaqmc_b200", "_C", "libjaqmc_b200_ffi.so"))
TARGETS = ("jaqmc_b200_ffi_logpsi", "jaqmc_b200_ffi_orbitals", "jaqmc_b200_ffi_local_energy",
           "jaqmc_b200_ffi_local_energy_complex", "jaqmc_b200_ffi_mh_step", "jaqmc_b200_ffi_coulomb", "jaqmc_b200_ffi_ewald")
_registered = False


def register():
    """Load the shim and register every handler for the CUDA platform.  No CPU registration: there is no CPU path."""
    global _registered
    if _registered:
        return
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(f"jaqmc_b200: {_LIB_PATH} not found -- build it with `make -C ffi FFI_INCLUDE=$(python -c "
                           "'import jax.ffi; print(jax.ffi.include_dir())')` after `python __graft_entry__.py`")
    lib = ctypes.CDLL(_LIB_PATH)
    for name in TARGETS:
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA")
    _registered = True


def _wf_attrs(packed):
    kind, config, fconfig, optional = packed
    return dict(kind=np.int32(kind), config=np.asarray(config, np.int32), fconfig=np.asarray(fconfig, np.float32),
                optional=np.int32(optional))


def leaves(params):
    """``jax.tree.leaves`` order == the order ``jaqmc_b200_bind_param_leaves`` expects (dict keys sorted)."""
    return [jnp.asarray(x, jnp.float32) for x in jax.tree.leaves(params)]


def logpsi(packed, params, electrons, atoms, extra=()):
    """One walker: ``electrons (n, 3)`` -> ``(logpsi (), sign ())``; ``vmap`` adds the walker axis without replicating
    the parameters (``vmap_method="expand_dims"``)."""
    register()
    call = jax.ffi.ffi_call("jaqmc_b200_ffi_logpsi",
                            (jax.ShapeDtypeStruct((), jnp.float32), jax.ShapeDtypeStruct((), jnp.float32)),
                            vmap_method="expand_dims")
    return call(electrons.astype(jnp.float32), atoms.astype(jnp.float32), *extra, *leaves(params), **_wf_attrs(packed))


def orbitals(packed, params, electrons, atoms, ndets, extra=(), complex_valued=False):
    register()
    n = electrons.shape[-2]
    shape = (ndets, n, n, 2) if complex_valued else (ndets, n, n)
    call = jax.ffi.ffi_call("jaqmc_b200_ffi_orbitals", jax.ShapeDtypeStruct(shape, jnp.float32), vmap_method="expand_dims")
    out = call(electrons.astype(jnp.float32), atoms.astype(jnp.float32), *extra, *leaves(params), **_wf_attrs(packed))
    return jax.lax.complex(out[..., 0], out[..., 1]) if complex_valued else out


def local_energy(packed, params, electrons, atoms, charges):
    """One walker -> dict(logpsi, sign, grad (3n,), lap, e_kin, e_pot, e_loc); ``sums`` (3,) are per-call partial sums."""
    register()
    n = electrons.shape[-2]
    f = lambda *s: jax.ShapeDtypeStruct(s, jnp.float32)  # noqa: E731
    call = jax.ffi.ffi_call("jaqmc_b200_ffi_local_energy", (f(), f(), f(3 * n), f(), f(), f(), f(), f(3)),
                            vmap_method="expand_dims")
    out = call(electrons.astype(jnp.float32), atoms.astype(jnp.float32), charges.astype(jnp.float32), *leaves(params),
               **_wf_attrs(packed))
    return dict(zip(("logpsi", "sign", "grad", "lap", "e_kin", "e_pot", "e_loc", "sums"), out))


def mh_step(packed, params, electrons, atoms, normals, uniforms, stddev, lattice=None, extra=()):
    """Batched: ``electrons (W, n, 3)``, ``normals (S, W, n, 3)``, ``uniforms (S, W)`` -> ``(electrons, logpsi, n_accept,
    accepted (S, W) u8)``.  The walker buffer is aliased in/out (donated state, workflow/stage/vmc.py:304)."""
    register()
    S, W = uniforms.shape
    call = jax.ffi.ffi_call(
        "jaqmc_b200_ffi_mh_step",
        (jax.ShapeDtypeStruct(electrons.shape, jnp.float32), jax.ShapeDtypeStruct((W,), jnp.float32),
         jax.ShapeDtypeStruct((1,), jnp.float32), jax.ShapeDtypeStruct((S, W), jnp.uint8)),
        input_output_aliases={0: 0})
    lat = np.zeros(0, np.float32) if lattice is None else np.asarray(lattice, np.float32).reshape(9)
    return call(electrons.astype(jnp.float32), atoms.astype(jnp.float32), normals.astype(jnp.float32),
                uniforms.astype(jnp.float32), jnp.reshape(stddev, (1,)).astype(jnp.float32), *extra, *leaves(params),
                lattice=lat, **_wf_attrs(packed))


def coulomb(electrons, atoms, charges):
    register()
    call = jax.ffi.ffi_call("jaqmc_b200_ffi_coulomb", jax.ShapeDtypeStruct((), jnp.float32), vmap_method="expand_dims")
    return call(electrons.astype(jnp.float32), atoms.astype(jnp.float32), charges.astype(jnp.float32))
