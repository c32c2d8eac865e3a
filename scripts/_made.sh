timeout 900 python -m pytest tests/test_gpu_anqs.py tests/test_gpu_nade.py -x -q 2>&1 | tail -3
timeout 300 python scripts/vmc_c5_phases.py 1048576 MADE 2>&1 | grep "rows\|made_forward\|made_backward\|batch_reduce_gemm" | cut -c1-220
