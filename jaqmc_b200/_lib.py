"""Loader of the sm_100a CUDA library.  There is no CPU path: a missing library is a hard error."""

from __future__ import annotations

import ctypes
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libjaqmc_b200.so")
_cdll = None


def cuda_library() -> ctypes.CDLL:
    """The CUDA build of the kernels (``jaqmc_b200/_C/libjaqmc_b200.so``, built by ``__graft_entry__.build()``)."""
    global _cdll
    if _cdll is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"jaqmc_b200: CUDA library not found at {LIB_PATH}. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback."
            )
        lib = _abi.bind(ctypes.CDLL(LIB_PATH))
        if b"sm_100a" not in lib.jaqmc_b200_version():
            raise RuntimeError(f"jaqmc_b200: {LIB_PATH} is not the sm_100a build: {lib.jaqmc_b200_version()!r}")
        _cdll = lib
    return _cdll
