mkdir -p gpurun_out/r2ag
timeout 900 python -m pytest tests/test_gpu_attention_op.py tests/test_gpu_attention_nets.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r2ag/tests.log
for wl in n2-psiformer n2-lapnet; do
A="--workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-vmc"
python bench.py $A > gpurun_out/r2ag/${wl}_mma.json 2> gpurun_out/r2ag/${wl}_mma.err
JAQMC_B200_ATTENTION_WARP=1 python bench.py $A > gpurun_out/r2ag/${wl}_warp.json 2> gpurun_out/r2ag/${wl}_warp.err
python bench.py $A > gpurun_out/r2ag/${wl}_mma2.json 2> gpurun_out/r2ag/${wl}_mma2.err
done
