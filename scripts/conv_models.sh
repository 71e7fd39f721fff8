timeout 400 python -m pytest tests/test_gpu_tf32_conv.py tests/test_gpu_models.py tests/test_gpu_unet.py tests/test_gpu_golden.py -q -x --timeout 100 2>&1 | tail -4 | cut -c1-300
timeout 200 python scripts/bench_configs.py resnet2d dilated2d 2>/dev/null | cut -c1-200
