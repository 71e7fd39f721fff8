timeout 600 python -m pytest tests/test_gpu_tf32.py -q --timeout 300 2>&1 | tail -12
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_golden.py -q -x --timeout 200 2>&1 | tail -2
timeout 200 python scripts/bench_configs.py fno3d fno2d 2>&1 | tail -3
