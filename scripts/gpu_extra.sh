#!/bin/bash
# Secondary workloads + a full ncu capture of the main FermiNet-N2 layer launch inside the real step.
TAG=${1:-x1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_tc_pair -s 7 -c 1 -o $OUT/main_layer python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-equilibrate --no-graph --no-vmc > $OUT/ncu_main.log 2>&1; echo "ncu exit $?"
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 exit $?"; tail -2 $OUT/bench_n2.err
for WL in ${WORKLOADS:-li n2-lapnet n2-psiformer}; do
  timeout 900 python bench.py --steps 3 --warmup 3 --workload $WL --cpu-seconds 6 > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err ; echo "bench $WL exit $?" ; tail -2 $OUT/bench_$WL.err
done
ls -la $OUT
