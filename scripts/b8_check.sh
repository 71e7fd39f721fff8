timeout 300 python -m pytest tests/test_gpu_tf32.py -q -x --timeout 100 2>&1 | tail -3 | cut -c1-300
for gb in 8 64; do timeout 120 python bench.py --global-batch $gb --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b$gb.json 2> gpurun_out/bench_b$gb.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b$gb.json')); print('B=$gb:', d['value'], d['ms_per_step'])"; done
