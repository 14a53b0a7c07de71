"""``jaqmc_b200_dense_fl`` on the GPU: the tcgen05 3xTF32 kernels (CTA pair with resident weights, weight-streaming
single CTA) and the CUDA-core FP32 kernel against a float64 reference of the forward-Laplacian dense rule
(reference laplacian/primitives/dot_general.py:377-407 + elementwise.py:42-72), over the shape cases that exercise
ragged output widths, 128-row groups, several feature blocks and single-chunk contractions.

``tests/gpu_dense_check.py`` is the stand-alone version of the same cases (a cheap canary to run under ``timeout``
before the suite: a pipeline deadlock in a hand-written mbarrier kernel hangs the device).
"""

import pytest

import gpu_dense_check as D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", D.CASES, ids=[c[0] for c in D.CASES])
def test_dense_fl_kernels_match_float64(case):
    from jaqmc_b200._lib import cuda_library

    e_simt, e_tc, nan_tc = D.run_case(cuda_library(), *case)
    assert nan_tc == 0
    assert e_simt < 5e-6, e_simt      # FP32 CUDA cores
    assert e_tc < 3e-5, e_tc          # 3xTF32: hi*hi + lo*hi + hi*lo, TMEM accumulation rounds toward zero
