timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_golden.py -q -x --timeout 300 2>&1 | tail -15
