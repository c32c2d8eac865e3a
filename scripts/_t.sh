timeout 900 python -m pytest tests/test_gpu_anqs.py -x -q 2>&1 | tail -2
timeout 300 python scripts/vmc_c5_phases.py 1048576 MADE 2>&1 | grep "rows\|anqs::" | cut -c1-250
