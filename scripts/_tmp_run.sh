mkdir -p gpurun_out/r2af
timeout 900 python -m pytest tests/test_gpu_attention_op.py -m gpu -q 2>&1 | tail -2 > gpurun_out/r2af/op.log
A="--workload benzene-psiformer --walkers 512 --steps 3 --warmup 3 --no-cpu-baseline --no-vmc"
python bench.py $A > gpurun_out/r2af/bz.json 2> gpurun_out/r2af/bz.err
