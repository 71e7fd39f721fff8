timeout 600 python -m pytest tests/test_gpu_tf32.py -q -x --timeout 300 2>&1 | tail -4
for f in 0 1 2; do echo "fuse=$f"; PDEQX_FUSE_R2C=$f timeout 120 python scripts/run_block.py; done
