set -x
timeout 300 python scripts/vmc_c5_phases.py 1048576 MADE 2>&1 | grep -v "^-\|^$" | cut -c1-200 | head -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"made_forward_kernel|made_backward_kernel|batch_reduce_gemm" -c 3 -o gpurun_out/r2_made_dmma python scripts/vmc_c5_phases.py 262144 MADE > gpurun_out/ncu_made.log 2>&1; tail -2 gpurun_out/ncu_made.log
