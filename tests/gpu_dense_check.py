"""Stand-alone GPU check of ``jaqmc_b200_dense_fl``: tcgen05 3xTF32 kernel vs the CUDA-core FP32 kernel vs float64.

Run it under ``timeout -s KILL`` before the pytest suite on a GPU box: a pipeline deadlock in a hand-written
mbarrier/tcgen05 kernel hangs the device, and this script is the cheap canary.  ``test_gpu_dense.py`` imports
``run_case`` for the pytest version.
"""

from __future__ import annotations

import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from jaqmc_b200 import _abi  # noqa: E402

CASES = [
    # name, G, C, k0, k1, N, groups_per_walker, act, res_mode, bias, addend
    ("n2_layer", 14 * 9, 44, 256, 64, 256, 14, 1, 1, True, True),
    ("n2_orbital", 14 * 5, 44, 256, 0, 224, 14, 0, 0, False, False),
    ("n2_mean", 40, 44, 512, 0, 256, 1, 0, 0, False, False),
    ("value_only", 1000, 1, 256, 64, 256, 14 * 0 + 8, 1, 1, True, True),
    ("li_layer", 3 * 50, 11, 256, 64, 256, 3, 1, 1, True, True),
    ("local1", 14 * 7, 5, 256, 0, 512 // 2, 14, 1, 2, True, False),
    ("minimal", 37, 8, 64, 0, 64, 1, 1, 0, True, False),
    ("wide_c", 6, 128, 64, 32, 128, 2, 1, 1, True, True),
    ("wide_c_many", 42 * 8, 128, 64, 0, 128, 42, 1, 0, True, False),   # both accumulator buffers, 128-row groups
    ("wide_n", 32 * 6, 98, 256, 0, 512, 32, 0, 0, False, False),        # four feature blocks (CTA-pair kernel only)
    ("n2_big", 14 * 300, 44, 256, 64, 256, 14, 1, 1, True, True),
    ("ragged_n48", 3 * 200, 11, 256, 0, 48, 3, 0, 0, False, False),     # N not a multiple of 32 (Li orbitals: 16 x 3)
    ("ragged_n80", 5 * 120, 17, 64, 32, 80, 5, 1, 0, True, True),
    ("ragged_n208", 14 * 60, 44, 256, 0, 208, 14, 1, 2, True, False),
    ("ragged_value", 3000, 1, 256, 0, 48, 3, 1, 0, True, True),
    ("k32_many", 14 * 400, 44, 32, 0, 256, 14, 1, 0, True, False),     # one K chunk per item, many items per CTA pair
    ("k64_many", 14 * 400, 11, 64, 0, 128, 14, 1, 2, True, False),
]


def reference(x, x2, k, k2, bias, addend, res, gpw, act, res_mode):
    d = torch.float64
    y = x.to(d) @ k.to(d)
    if x2 is not None:
        y = y + x2.to(d) @ k2.to(d)
    G, Cc, N = y.shape
    if addend is not None:
        y = (y.reshape(G // gpw, gpw, Cc, N) + addend.to(d)[:, None]).reshape(G, Cc, N)
    if bias is not None:
        y[:, 0] += bias.to(d)
    if act == 1:
        t = torch.tanh(y[:, 0])
        d1 = 1 - t * t
        out = torch.empty_like(y)
        out[:, 0] = t
        if Cc > 1:
            j = y[:, 1:Cc - 1]
            out[:, 1:Cc - 1] = d1[:, None] * j
            out[:, Cc - 1] = d1 * y[:, Cc - 1] - 2 * t * d1 * (j * j).sum(1)
        y = out
    if res_mode == 1:
        y = (res.to(d) + y) / (2.0 ** 0.5)
    elif res_mode == 2:
        y = res.to(d) + y
    return y


def run_case(lib, name, G, Cc, k0, k1, N, gpw, act, res_mode, with_bias, with_addend, seed=0, dev="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32).to(dev)  # noqa: E731
    x = r(G, Cc, k0)
    x2 = r(G, Cc, k1) if k1 else None
    k = r(k0, N) / (k0 + k1) ** 0.5
    k2 = r(k1, N) / (k0 + k1) ** 0.5 if k1 else None
    bias = 0.3 * r(N) if with_bias else None
    addend = r(G // gpw, Cc, N) if with_addend else None
    res = r(G, Cc, N) if res_mode else None
    ws = torch.empty(2 * (k0 + k1) * N * 4 + 1024, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)  # noqa: E731
    outs = {}
    for use_tc in (0, 1, 2):  # CUDA cores | tcgen05 (CTA pair, resident weights, when eligible) | tcgen05 streaming
        out = torch.full((G, Cc, N), float("nan"), dtype=torch.float32, device=dev)
        rc = lib.jaqmc_b200_dense_fl(p(x), p(x2), p(k), p(k2), p(bias), p(addend), p(res), p(out), G, Cc, k0, k1, N, gpw,
                                     act, res_mode, use_tc, p(ws), ws.numel(),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _abi.check(lib, rc)
        torch.cuda.synchronize()
        outs[use_tc] = out
    ref = reference(x, x2, k, k2, bias, addend, res, gpw, act, res_mode)
    scale = ref.abs().max().item() + 1e-30
    e_simt = (outs[0].double() - ref).abs().max().item() / scale
    e_tc = max((outs[m].double() - ref).abs().max().item() / scale for m in (1, 2))
    nan_tc = int(torch.isnan(outs[1]).sum() + torch.isnan(outs[2]).sum())
    if os.environ.get("DENSE_CHECK_VERBOSE"):
        print("   pair %.2e  streaming %.2e" % tuple((outs[m].double() - ref).abs().max().item() / scale for m in (1, 2)))
    return e_simt, e_tc, nan_tc


def main():
    from jaqmc_b200._lib import cuda_library

    lib = cuda_library()
    bad = 0
    for case in CASES:
        e_simt, e_tc, nan_tc = run_case(lib, *case)
        ok = nan_tc == 0 and e_tc < 3e-5 and e_simt < 5e-6
        print(f"{case[0]:12s} simt {e_simt:.2e}  tc {e_tc:.2e}  nan {nan_tc}  {'ok' if ok else 'FAIL'}", flush=True)
        bad += not ok
    print("dense check:", "PASS" if not bad else f"{bad} FAILED")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
