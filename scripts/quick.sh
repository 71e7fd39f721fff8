timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32_v5.json 2> gpurun_out/bench_tf32_v5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tf32_v5.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['physics_conv'])"
timeout 300 python scripts/bench_resnet.py 2>&1 | tail -2
