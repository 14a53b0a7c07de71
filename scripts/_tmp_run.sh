mkdir -p gpurun_out/r2as
timeout 1500 python -m pytest tests -m gpu -q -k "gradient or estimators" 2>&1 | tail -1 > gpurun_out/r2as/tests.log
python scripts/profile_vjp.py > gpurun_out/r2as/vjp.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_env_param_grad|k_pair_mean_bwd|k_tanh_bwd|k_spin_mean_bwd" -c 30 --csv --log-file gpurun_out/r2as/k.csv python scripts/profile_vjp.py --calls 1 > gpurun_out/r2as/ncu.log 2>&1
