mkdir -p gpurun_out/r2ab
timeout 900 python -m pytest tests/test_gpu_layernorm_op.py -m gpu -q -s 2>&1 | grep -E "C=[0-9]+ F=|passed|failed|Error|assert|FAILED" | sed 's/^\.*//' > gpurun_out/r2ab/ln_op.log
A="--workload benzene-psiformer --walkers 512 --steps 3 --warmup 3 --no-cpu-baseline --no-vmc"
python bench.py $A > gpurun_out/r2ab/bz_rows.json 2> gpurun_out/r2ab/bz_rows.err
JAQMC_B200_LAYERNORM_SMEM=1 python bench.py $A > gpurun_out/r2ab/bz_smem.json 2> gpurun_out/r2ab/bz_smem.err
B="--workload n2-psiformer --steps 3 --warmup 3 --no-cpu-baseline --no-vmc"
python bench.py $B > gpurun_out/r2ab/n2p_rows.json 2> gpurun_out/r2ab/n2p_rows.err
JAQMC_B200_LAYERNORM_SMEM=1 python bench.py $B > gpurun_out/r2ab/n2p_smem.json 2> gpurun_out/r2ab/n2p_smem.err
