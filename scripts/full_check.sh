timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -4
timeout 120 python scripts/run_block.py
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -c 1500 gpurun_out/bench_r1_n1.json; tail -3 gpurun_out/bench_r1_n1.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; head -c 600 gpurun_out/bench_r1_ref.json
