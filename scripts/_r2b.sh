#!/bin/bash
# round-2 re-entry check: sampler with unmasked levels + where the table build spends its time (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_anqs.py -x -q 2>&1 | tail -5
python scripts/table_build_time.py 2>&1 | tee gpurun_out/table_build_time.txt
REPS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/table_build_launches.csv python scripts/table_build_time.py > gpurun_out/table_build_ncu.log 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/table_build_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[1:]:
    if 'filter' in r[ki] or 'hash' in r[ki] or 'fill' in r[ki].lower() or 'memset' in r[ki].lower(): print(r[ki][:60], r[vi])
P
