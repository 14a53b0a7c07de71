"""Builds the HOST-EMULATION library of the kernels -- TEST INFRASTRUCTURE ONLY.

The kernel sources under ``jaqmc_b200/csrc`` are written in block-stride style (see ``common.cuh``), so
the same files compile with g++ (``-DJAQMC_HOST_EMU``: one thread per block, blocks run sequentially).
The CPU test-suite uses this build to check the kernels' arithmetic and the host-side orchestration
against the float64 oracle without a GPU.  It is never loaded by the ``jaqmc_b200`` package: the product
path is the nvcc build and fails loudly when that is missing.  The tcgen05 kernel (``dense_tc.cu``) has no
host build; the emulation routes every dense layer through the CUDA-core kernel's scalar twin.
"""

from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "jaqmc_b200", "csrc")
OUT_DIR = os.path.join(ROOT, "tests", "emu", "_build")
LIB = os.path.join(OUT_DIR, "libjaqmc_b200_emu.so")
EMU_SKIP = {"dense_tc.cu"}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") and f not in EMU_SKIP)


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in sources()]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(ROOT, "include", "jaqmc_b200.h"))
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DJAQMC_HOST_EMU", "-ffp-contract=off", "-o", LIB]
    for s in srcs:
        cmd += ["-x", "c++", s]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
