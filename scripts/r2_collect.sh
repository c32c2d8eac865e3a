#!/bin/bash
# Round-2 evidence on one B200 (run under gpurun): GPU tests, timings of the new kernels, the bench line, the ncu launch list of
# the bench and one full ncu capture of the enumeration kernels + the dominant fused kernel.
python -m pytest tests -m gpu -q 2>&1 | tail -2
python scripts/pair_join_time.py > gpurun_out/r2_pair_join.txt 2>&1
python scripts/full_mode_time.py > gpurun_out/r2_full_mode.txt 2>&1
python scripts/enum_time.py 16384 65536 > gpurun_out/r2_enum_time.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_launch_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"enum_|fused_eloc_bs" -c 3 -o gpurun_out/prof_r2_final -f python scripts/enum_prof.py 16384 > gpurun_out/ncu_r2_final.log 2>&1; tail -1 gpurun_out/ncu_r2_final.log
