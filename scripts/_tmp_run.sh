mkdir -p gpurun_out/r2aa
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2aa/gpu_tests.log
