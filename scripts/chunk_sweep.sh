for c in 0 1 2 4 8 16; do echo "chunk=$c"; PDEQX_BWD_CHUNK=$c timeout 120 python scripts/run_block.py; done
timeout 300 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_models.py -q -x --timeout 200 2>&1 | tail -2
