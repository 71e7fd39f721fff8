timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['physics_conv']['frac'], d['cpu_baseline']['value'])"
timeout 300 python scripts/bench_configs.py resnet2d dilated2d unet1d > gpurun_out/r1_configs_conv.jsonl 2>/dev/null; cut -c1-200 gpurun_out/r1_configs_conv.jsonl
