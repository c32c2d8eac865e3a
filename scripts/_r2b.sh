#!/bin/bash
# round-2 re-entry check: table build with byte-table hashes and four keys in flight (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_hamiltonian.py tests/test_gpu_vmc.py -x -q 2>&1 | tail -3
python scripts/table_build_time.py 2>&1 | tee gpurun_out/table_build_time.txt
SIZES=8388608 REPS=1 ncu --set full --clock-control none --import-source on -k regex:'hash_build|filter_count' -c 4 -o gpurun_out/hash_build_8m_v2 -f python scripts/table_build_time.py > gpurun_out/hash_build_ncu.log 2>&1
tail -2 gpurun_out/hash_build_ncu.log
