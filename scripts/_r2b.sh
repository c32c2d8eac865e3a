#!/bin/bash
# round-2 re-entry check: all GPU tests + where the table build spends its time + the fused kernel on an 8M-key table (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/table_build_time.py 2>&1 | tee gpurun_out/table_build_time.txt
python scripts/fused_bigtable.py 2>&1 | tee gpurun_out/fused_bigtable.txt
REPS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/table_build_launches.csv python scripts/table_build_time.py > gpurun_out/table_build_ncu.log 2>&1
python - <<'P'
import csv
rows = [r for r in csv.reader(open('gpurun_out/table_build_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[1:]:
    if any(t in r[ki] for t in ('filter', 'hash', 'part_')): print(r[ki][:60], r[vi])
P
