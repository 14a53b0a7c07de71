#!/bin/bash
# compute-sanitizer over the kernels added late in r2 (k_logdet_tiny, staged k_logdet_small, k_logdet_c staging / tile-major M,
# k_logdet_combine_c_warp, warp-per-matrix k_minv_small<32>).  Run on a GPU box: bash scripts/sanitize_late.sh out_dir
set -u
OUT=${1:-gpurun_out/sanitize_late}
mkdir -p "$OUT"
S=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $S --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_det_sizes.py tests/test_gpu_solid.py tests/test_gpu_gradients.py tests/test_gpu_ferminet.py -m gpu -q -x \
   -k "determinant_sizes or cubic_h2 or lih_221 or Ar or Zn or Li" > "$OUT/memcheck_late.log" 2>&1
timeout 900 $S --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_det_sizes.py tests/test_gpu_solid.py tests/test_gpu_gradients.py -m gpu -q -x \
   -k "n6_d4 or n12_d2 or n16_d4 or n17_d2 or cubic_h2 or lih_221 or Ar" > "$OUT/racecheck_late.log" 2>&1
timeout 600 $S --tool initcheck --error-exitcode 7 python -m pytest tests/test_gpu_det_sizes.py tests/test_gpu_solid.py tests/test_gpu_gradients.py -m gpu -q -x \
   -k "determinant_sizes or cubic_h2 or lih_221 or Ar" > "$OUT/initcheck_late.log" 2>&1
for f in "$OUT"/*_late.log; do echo "== $f"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" $f | tail -3; done
