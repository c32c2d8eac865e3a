#!/bin/bash
# round-2 final evidence on one B200 (run under gpurun): all GPU tests, table build timing, the bench line, its launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/table_build_time.py 2>&1 | tee gpurun_out/table_build_time.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err; echo bench rc=$?
python scripts/bench_brief.py gpurun_out/r2b_bench_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_launch_r2b.log 2>&1; echo ncu rc=$?
