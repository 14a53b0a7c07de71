mkdir -p gpurun_out/r2bg
python bench.py --walkers 512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2bg/n2_512.json 2> gpurun_out/r2bg/n2_512.err
