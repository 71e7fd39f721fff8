timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tf32.py -q -x --timeout 100 2>&1 | tail -3 | cut -c1-300
timeout 120 python bench.py --global-batch 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b8.json')); print('B=8:', d['value'], d['ms_per_step'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'mode_gemm|c2c_kernel|adam' -s 100 -c 60 --csv --log-file gpurun_out/b8_modes.csv python bench.py --global-batch 8 --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
