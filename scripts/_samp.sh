timeout 900 python -m pytest tests/test_gpu_anqs.py tests/test_gpu_nade.py tests/test_gpu_vmc.py tests/test_gpu_transformer.py -x -q 2>&1 | tail -3
timeout 300 python scripts/sampler_c5_phases.py 2>&1 | grep "samples ->\|split_\|made_tc\|emit_children" | cut -c1-220
