#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list (+ optional full capture of one kernel).
# usage: scripts/gpu_round.sh <tag> [ncu-kernel-regex] [ncu-skip]
TAG=${1:-r1}
KREGEX=${2:-}
KSKIP=${3:-9}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== dense canary" ; timeout -s KILL 180 python tests/gpu_dense_check.py > $OUT/dense_check.log 2>&1 ; RC=$? ; tail -12 $OUT/dense_check.log
if [ $RC -ne 0 ]; then
  echo "dense canary rc=$RC: continuing with the tcgen05 kernel disabled"
  export JAQMC_B200_DISABLE_TC=1
  nvidia-smi > $OUT/after_canary_smi.txt 2>&1 || { echo "GPU unresponsive after canary"; exit 1; }
fi
if [ -z "$SKIP_TESTS" ]; then
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1 ; echo "pytest exit $?" ; grep -E "passed|failed|scaled|err" $OUT/pytest_gpu.log | tail -40
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "smoke exit $?" ; tail -5 $OUT/smoke.log
fi
echo "== bench" ; timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench exit $?" ; tail -3 $OUT/bench.err ; cat $OUT/bench.json
for WL in $EXTRA_WORKLOADS; do
  echo "== bench $WL" ; timeout 900 python bench.py --steps 3 --warmup 3 --workload $WL --cpu-seconds 8 > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err ; echo "bench exit $?" ; tail -3 $OUT/bench_$WL.err ; cat $OUT/bench_$WL.json
done
echo "== ncu launch list" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-equilibrate --no-graph --no-vmc --evals-per-step 1 > $OUT/ncu_bench.log 2>&1 ; echo "ncu exit $?"
if [ -n "$KREGEX" ]; then
  echo "== ncu full $KREGEX" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $KSKIP -c 2 -o $OUT/prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-equilibrate --no-graph --no-vmc --evals-per-step 1 --walkers ${NCU_WALKERS:-4096} > $OUT/ncu_full.log 2>&1 ; echo "ncu full exit $?"; tail -3 $OUT/ncu_full.log
fi
ls -la $OUT
