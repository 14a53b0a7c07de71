mkdir -p gpurun_out/r2an
timeout 1500 python -m pytest tests -m gpu -q -k "sampling or ferminet or observables or reference_fixtures or estimators" 2>&1 | tail -2 > gpurun_out/r2an/tests.log
python scripts/profile_mh.py > gpurun_out/r2an/mh_new.log 2>&1
JAQMC_B200_PAIR_LAYER_ONE_PAIR=1 python scripts/profile_mh.py > gpurun_out/r2an/mh_old.log 2>&1
python scripts/profile_mh.py > gpurun_out/r2an/mh_new2.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:k_pair_layer" -c 3 --csv --log-file gpurun_out/r2an/pl.csv python scripts/profile_mh.py --eager --calls 1 > gpurun_out/r2an/ncu0.log 2>&1
