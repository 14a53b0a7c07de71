"""Times ``jaqmc_b200_dense_fl`` alone (CUDA events) on the FermiNet-N2 launch shapes at 4096 walkers.

usage: python scripts/dense_bench.py [case ...]     (GPU box only; quick iteration on the tcgen05 kernel)
"""

from __future__ import annotations

import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from jaqmc_b200 import _abi  # noqa: E402
from jaqmc_b200._lib import cuda_library  # noqa: E402

W = int(os.environ.get("WALKERS", 4096))
CASES = {
    # name: G, C, k0, k1, N, groups_per_walker, act, res_mode, bias, addend
    "main": (W * 14, 44, 256, 64, 256, 14, 1, 1, True, True),
    "orbital": (W * 7, 44, 256, 0, 224, 7, 0, 0, False, False),
    "layer1": (W * 14, 44, 32, 0, 256, 14, 1, 0, True, False),
    "mean": (W, 44, 512, 0, 256, 1, 0, 0, False, False),
    "psi_dense": (W * 14, 44, 256, 0, 256, 14, 1, 2, True, False),
    "value_main": (W * 14, 1, 256, 64, 256, 14, 1, 1, True, True),
    "value_plain": (W * 14, 1, 256, 0, 256, 14, 0, 0, False, False),
    "value_big": (W * 14 * 16, 1, 256, 64, 256, 14, 1, 1, True, True),
}


def main():
    lib = cuda_library()
    dev = "cuda"
    names = sys.argv[1:] or ["main", "orbital"]
    for name in names:
        G, Cc, k0, k1, N, gpw, act, res_mode, wb, wa = CASES[name]
        x = torch.randn(G, Cc, k0, device=dev)
        x2 = torch.randn(G, Cc, k1, device=dev) if k1 else None
        k = torch.randn(k0, N, device=dev) / (k0 + k1) ** 0.5
        k2 = torch.randn(k1, N, device=dev) / (k0 + k1) ** 0.5 if k1 else None
        bias = torch.randn(N, device=dev) if wb else None
        addend = torch.randn(G // gpw, Cc, N, device=dev) if wa else None
        res = torch.randn(G, Cc, N, device=dev) if res_mode else None
        out = torch.empty(G, Cc, N, device=dev)
        ws = torch.empty(2 * (k0 + k1) * N * 4 + 1024, dtype=torch.uint8, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)  # noqa: E731
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        for mode, label in ((1, "pair"), (2, "stream")):
            def call():
                rc = lib.jaqmc_b200_dense_fl(p(x), p(x2), p(k), p(k2), p(bias), p(addend), p(res), p(out), G, Cc, k0, k1,
                                             N, gpw, act, res_mode, mode, p(ws), ws.numel(), st)
                _abi.check(lib, rc)

            for _ in range(3):
                call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                call()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            flops = 2.0 * G * Cc * (k0 + k1) * N
            print(f"{name:10s} {label:6s} rows {G * Cc:9d} K {k0 + k1:4d} N {N:4d}: {ms:7.3f} ms  "
                  f"{flops / ms / 1e9:7.1f} TFLOP/s (x3 products)", flush=True)


if __name__ == "__main__":
    main()
