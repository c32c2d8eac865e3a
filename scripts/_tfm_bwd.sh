set -x
timeout 300 python -m pytest tests/test_gpu_transformer.py -x -q 2>&1 | tail -8
timeout 300 python scripts/tfm_bwd_prof.py 10000 3 2>&1 | tail -4
