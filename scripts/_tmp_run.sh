mkdir -p gpurun_out/r2be
timeout 1500 python -m pytest tests -m gpu -q -k "solid or reference_fixtures or leaf" 2>&1 | tail -6 > gpurun_out/r2be/tests.log
