#!/bin/bash
# last check of the round on one B200 (run under gpurun): all GPU tests, smoke(), table build timing with the shipped insert kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/table_build_time.py 2>&1 | tee gpurun_out/table_build_time_final.txt
