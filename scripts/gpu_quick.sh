#!/bin/bash
# quick GPU check: parity tests + two bench sizes (run under gpurun)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --n-unq 65536 --no-cpu-baseline > gpurun_out/bench_64k.log 2>&1; python scripts/bench_brief.py gpurun_out/bench_64k.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1; python scripts/bench_brief.py gpurun_out/bench_1m.log
