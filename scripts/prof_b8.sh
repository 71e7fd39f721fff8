python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --global-batch 8 > gpurun_out/plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 650 -c 240 --csv --log-file gpurun_out/b8_step_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --global-batch 8 > /dev/null 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --global-batch 8 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('B=8 1GPU graph:', d['value'], d['ms_per_step'])"
