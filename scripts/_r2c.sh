#!/bin/bash
# ncu --set full of the table insert kernels at 8.4M keys (run under gpurun)
mkdir -p gpurun_out
SIZES=8388608 REPS=1 ncu --set full --clock-control none --import-source on -k regex:'hash_build' -c 4 -o gpurun_out/hash_build_8m -f python scripts/table_build_time.py > gpurun_out/hash_build_ncu.log 2>&1
tail -3 gpurun_out/hash_build_ncu.log
ls -la gpurun_out/
