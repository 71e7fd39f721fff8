timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32_v4.json 2> gpurun_out/bench_tf32_v4.err; head -c 330 gpurun_out/bench_tf32_v4.json; echo; tail -2 gpurun_out/bench_tf32_v4.err
