timeout 300 python -m pytest tests/test_gpu_tf32.py -q -x --timeout 120 2>&1 | tail -3
timeout 120 python scripts/run_block.py
