mkdir -p gpurun_out/r2az
timeout 1500 python -m pytest tests -m gpu -q -k "solid" 2>&1 | tail -2 > gpurun_out/r2az/tests.log
A="--workload lih-solid --walkers 512 --steps 3 --warmup 3 --no-cpu-baseline --no-vmc"
python bench.py $A > gpurun_out/r2az/lih.json 2> gpurun_out/r2az/lih.err
python bench.py $A > gpurun_out/r2az/lih2.json 2> gpurun_out/r2az/lih2.err
