#!/bin/bash
# compute-sanitizer passes over the GPU tests that exercise the hand-written kernels with small inputs (memcheck is
# ~10x, racecheck ~50x slower than a plain run).  Run on a GPU box from the repo root:  bash scripts/sanitize.sh out_dir
set -u
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
S=/usr/local/cuda/bin/compute-sanitizer
$S --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_attention_op.py tests/test_gpu_layernorm_op.py -m gpu -q -x \
   -k "n17 or n42 or n48 or n14 or C44 or C128 or C23" > "$OUT/memcheck_ops.log" 2>&1
$S --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ferminet.py tests/test_gpu_gradients.py tests/test_gpu_solid.py -m gpu -q -x \
   -k "Ar or small or options or vjp_matches_autograd_small" > "$OUT/memcheck_nets.log" 2>&1
$S --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_attention_op.py -m gpu -q -x \
   -k "n17-dense or n42-dense or n14-local" > "$OUT/racecheck_attn.log" 2>&1
$S --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_ferminet.py tests/test_gpu_solid.py tests/test_gpu_gradients.py \
   tests/test_gpu_layernorm_op.py -m gpu -q -x -k "Ar or cubic_h2 or Li or C44 or C11 or vjp_matches_autograd_small" > "$OUT/racecheck_nets.log" 2>&1
$S --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_attention_nets.py tests/test_gpu_sampling.py -m gpu -q -x \
   -k "small or sampling" > "$OUT/racecheck_nets2.log" 2>&1
grep -H -E "ERROR SUMMARY|RACECHECK SUMMARY" "$OUT"/*.log
$S --tool initcheck --error-exitcode 7 python -m pytest tests/test_gpu_ferminet.py tests/test_gpu_solid.py tests/test_gpu_attention_nets.py \
   tests/test_gpu_gradients.py -m gpu -q -x -k "small or cubic_h2 or Ar or options" > "$OUT/initcheck.log" 2>&1
grep -H -E "ERROR SUMMARY" "$OUT/initcheck.log"
