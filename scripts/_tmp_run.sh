mkdir -p gpurun_out/r2ar
timeout 1500 python -m pytest tests -m gpu -q -s -k "gradient or estimators" 2>&1 | grep -E "^\.*vjp|passed|failed|FAILED|Error" | sed 's/^\.*//' > gpurun_out/r2ar/tests_mma.log
JAQMC_B200_GEMM_TN_SIMT=1 timeout 1500 python -m pytest tests -m gpu -q -s -k "gradient or estimators" 2>&1 | grep -E "^\.*vjp|passed|failed|FAILED|Error" | sed 's/^\.*//' > gpurun_out/r2ar/tests_simt.log
python scripts/profile_vjp.py > gpurun_out/r2ar/vjp_mma.log 2>&1
JAQMC_B200_GEMM_TN_SIMT=1 python scripts/profile_vjp.py > gpurun_out/r2ar/vjp_simt.log 2>&1
python scripts/profile_vjp.py > gpurun_out/r2ar/vjp_mma2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_gemm_tn -c 34 --csv --log-file gpurun_out/r2ar/gemm.csv python scripts/profile_vjp.py --calls 1 > gpurun_out/r2ar/ncu.log 2>&1
