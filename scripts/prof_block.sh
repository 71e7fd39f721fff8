timeout 300 python -m pytest tests/test_gpu_tf32.py -q --timeout 200 2>&1 | tail -8
python scripts/run_block.py && ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/block_launches.csv python scripts/run_block.py > /dev/null 2>&1
