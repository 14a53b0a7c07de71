mkdir -p gpurun_out/r2x
run() { ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_logdet_small -c 4 --csv --log-file gpurun_out/r2x/ldm_$1.csv python bench.py --steps 1 --warmup 1 --evals-per-step 1 --no-cpu-baseline --no-vmc --no-equilibrate --no-graph > gpurun_out/r2x/bm_$1.log 2>&1; }
JAQMC_B200_LOGDET_PREFETCH=1 run 2
JAQMC_B200_LOGDET_SLABS=3 run 3
JAQMC_B200_LOGDET_SLABS=4 run 4
JAQMC_B200_LOGDET_SLABS=3 python -m pytest tests -m gpu -x -q -k "logdet or parity" 2>&1 | tail -3 > gpurun_out/r2x/tests3.log
