set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transformer_backward_kernel -c 2 -o gpurun_out/r2_tfm_bwd2 python scripts/tfm_bwd_prof.py 10000 0 > gpurun_out/ncu_tfm.log 2>&1; tail -2 gpurun_out/ncu_tfm.log
