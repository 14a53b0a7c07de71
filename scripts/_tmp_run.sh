mkdir -p gpurun_out/r2bb
timeout 1500 python -m pytest tests -m gpu -q -k "ferminet or attention_nets" 2>&1 | tail -2 > gpurun_out/r2bb/tests.log
A="--workload benzene-psiformer --walkers 512 --steps 3 --warmup 3 --no-cpu-baseline --no-vmc"
python bench.py $A > gpurun_out/r2bb/bz.json 2> gpurun_out/r2bb/bz.err
