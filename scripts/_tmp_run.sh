mkdir -p gpurun_out/r2ao
python bench.py > gpurun_out/r2ao/bench_n2.json 2> gpurun_out/r2ao/bench_n2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ao/bench_n2_ref.json 2> gpurun_out/r2ao/bench_n2_ref.err
