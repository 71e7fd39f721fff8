timeout 300 python -m pytest tests/test_gpu_tf32_conv.py -q -x --timeout 120 2>&1 | tail -5
python scripts/run_conv.py; DIL=8 python scripts/run_conv.py; MODE=dirichlet python scripts/run_conv.py; PREC=fp32 python scripts/run_conv.py
