mkdir -p gpurun_out/r2ah
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2ah/gpu_tests.log
