timeout 120 python -m pytest tests/test_gpu_tf32_conv.py -q -x --timeout 50 2>&1 | tail -3 | cut -c1-300
timeout 60 python scripts/run_conv.py; MODE=dirichlet DIL=2 timeout 60 python scripts/run_conv.py
