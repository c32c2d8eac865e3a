set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_tfm_launches.csv python scripts/tfm_bwd_prof.py 10000 1 > gpurun_out/ncu_tfm.log 2>&1; tail -2 gpurun_out/ncu_tfm.log
