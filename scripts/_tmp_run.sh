mkdir -p gpurun_out/r2y
python bench.py --steps 5 --warmup 3 > gpurun_out/r2y/bench_n2.json 2> gpurun_out/r2y/bench_n2.err
JAQMC_B200_LOGDET_SLABS=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-vmc > gpurun_out/r2y/bench_n2_ns1.json 2> gpurun_out/r2y/bench_n2_ns1.err
