timeout 300 python -m pytest tests/test_gpu_tf32_conv.py -q -x --timeout 60 2>&1 | tail -3 | cut -c1-200
timeout 120 python scripts/run_conv.py; MODE=dirichlet DIL=2 timeout 120 python scripts/run_conv.py; MODE=neumann DIL=8 timeout 120 python scripts/run_conv.py
