mkdir -p gpurun_out/r2ax
A="--workload lih-solid --walkers 512 --steps 3 --warmup 3 --no-cpu-baseline --no-vmc"
python bench.py $A > gpurun_out/r2ax/lih_async.json 2> gpurun_out/r2ax/lih_async.err
JAQMC_B200_LOGDET_PLAIN_STAGING=1 python bench.py $A > gpurun_out/r2ax/lih_plain.json 2> gpurun_out/r2ax/lih_plain.err
python bench.py $A > gpurun_out/r2ax/lih_async2.json 2> gpurun_out/r2ax/lih_async2.err
