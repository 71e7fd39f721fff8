timeout 200 python -m pytest tests/test_gpu_tf32_conv.py -q -x --timeout 60 2>&1 | tail -3 | cut -c1-200
MODE=neumann timeout 120 python scripts/run_conv.py; MODE=neumann DIL=8 timeout 120 python scripts/run_conv.py
