mkdir -p gpurun_out/r2am
timeout 1500 python -m pytest tests -m gpu -q -k "sampling or ferminet or attention_nets or observables or reference_fixtures or estimators" 2>&1 | tail -2 > gpurun_out/r2am/tests.log
python scripts/profile_mh.py > gpurun_out/r2am/mh_fused.log 2>&1
JAQMC_B200_UNFUSED_ENV_LOGDET=1 python scripts/profile_mh.py > gpurun_out/r2am/mh_unfused.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_logdet_value_warp" -c 2 --csv --log-file gpurun_out/r2am/fused.csv python scripts/profile_mh.py --eager --calls 1 > gpurun_out/r2am/ncu0.log 2>&1
JAQMC_B200_UNFUSED_ENV_LOGDET=1 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_logdet_value_warp" -c 2 --csv --log-file gpurun_out/r2am/unfused.csv python scripts/profile_mh.py --eager --calls 1 > gpurun_out/r2am/ncu1.log 2>&1
