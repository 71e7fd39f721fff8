timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -4
timeout 120 python scripts/run_block.py
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -c 1800 gpurun_out/bench_r1_n1.json; tail -3 gpurun_out/bench_r1_n1.err
bash scripts/prof_traffic.sh
