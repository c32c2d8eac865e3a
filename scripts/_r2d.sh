#!/bin/bash
# ncu --set full of the bit-sliced fused kernel on a 1M-row window of an 8.4M-key table: the per-rank work of the 8-GPU bench (run under gpurun)
mkdir -p gpurun_out
ONLY_BS=1 REPS=3 ncu --set full --clock-control none --import-source on -k regex:fused_eloc_bs --launch-skip 1 -c 1 -o gpurun_out/prof_r2b_fused_8m -f python scripts/fused_bigtable.py 8388608 > gpurun_out/ncu_r2b_fused_8m.log 2>&1
tail -3 gpurun_out/ncu_r2b_fused_8m.log
