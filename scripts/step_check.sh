timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tf32.py tests/test_gpu_models.py tests/test_gpu_golden.py -q -x --timeout 300 2>&1 | tail -4
timeout 120 python scripts/run_block.py
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tmp.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/step_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
