#!/bin/bash
# ncu --set full on selected kernels of one FermiNet-N2 step: scripts/gpu_prof.sh <tag> <regex1> [regex2 ...]
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for K in "$@"; do
  NAME=$(echo $K | tr -cd 'a-zA-Z0-9_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-4} -c 1 -o $OUT/prof_$NAME python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-equilibrate --no-graph --no-vmc > $OUT/ncu_$NAME.log 2>&1
  echo "ncu $K exit $?"
done
ls -la $OUT
