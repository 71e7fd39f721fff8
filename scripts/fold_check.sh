timeout 600 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_models.py -q -x --timeout 300 2>&1 | tail -6
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; tail -2 gpurun_out/bench_tmp.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tmp.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['config']['launch'], d['gpu_launches'])"
