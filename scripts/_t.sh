timeout 900 python -m pytest tests/test_gpu_hamiltonian.py tests/test_gpu_vmc.py -x -q 2>&1 | tail -3
timeout 600 python scripts/bench_vmc_sharded.py --steps 20 2>&1 | grep "per-iteration\|vmc_iterations" | cut -c1-700
