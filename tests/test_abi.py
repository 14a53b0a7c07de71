"""The built CUDA library exports every symbol ``include/jaqmc_b200.h`` declares, and the ctypes mirror covers the
header (no compute calls: runs without a GPU)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jaqmc_b200.h")
LIB = os.path.join(ROOT, "jaqmc_b200", "_C", "libjaqmc_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jaqmc_b200_\w+)\s*\(", src)))


def test_header_declares_the_entry_points():
    names = declared_symbols()
    for must in ("jaqmc_b200_logpsi", "jaqmc_b200_local_energy", "jaqmc_b200_local_energy_complex", "jaqmc_b200_mh_step",
                 "jaqmc_b200_mh_step_pbc", "jaqmc_b200_coulomb", "jaqmc_b200_ewald", "jaqmc_b200_dense_fl", "jaqmc_b200_attention_fl", "jaqmc_b200_layernorm_fl",
                 "jaqmc_b200_workspace_bytes"):
        assert must in names


@pytest.mark.skipif(not os.path.exists(LIB), reason="CUDA library not built (run __graft_entry__.build())")
def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in ctypes.cast(lib.jaqmc_b200_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()


def test_ctypes_mirror_covers_the_header():
    from jaqmc_b200 import _abi

    assert sorted(_abi.PROTOTYPES) == declared_symbols()


def test_product_refuses_to_run_without_cuda():
    """No CPU fallback: the product runtime raises instead of computing on the host."""
    import torch

    from jaqmc_b200._runtime import runtime

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        runtime()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        runtime("cpu")
