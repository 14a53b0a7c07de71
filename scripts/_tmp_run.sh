mkdir -p gpurun_out/r2bd
timeout 1500 python -m pytest tests -m gpu -q -k "solid" 2>&1 | tail -2 > gpurun_out/r2bd/tests.log
A="--workload lih-solid --walkers 512 --steps 3 --warmup 3 --no-cpu-baseline --no-vmc"
python bench.py $A > gpurun_out/r2bd/lih.json 2> gpurun_out/r2bd/lih.err
