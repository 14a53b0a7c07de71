#!/bin/bash
# canary for the tcgen05 kernel: hard-killed on a hang so the box is never left stuck
TAG=${1:-dense}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 180 python tests/gpu_dense_check.py > $OUT/dense_check.log 2>&1
echo "dense check exit $?"
cat $OUT/dense_check.log | tail -30
