#!/bin/bash
# last check of the round on one B200 (run under gpurun): all GPU tests, then the fused kernel against an 8.4M-key table
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ONLY_BS=1 REPS=5 python scripts/fused_bigtable.py 8388608 2>&1 | tee gpurun_out/fused_bigtable_final.txt
