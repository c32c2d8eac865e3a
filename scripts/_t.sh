timeout 900 python -m pytest tests/test_gpu_hamiltonian.py -x -q 2>&1 | tail -3
bash scripts/_ab.sh
