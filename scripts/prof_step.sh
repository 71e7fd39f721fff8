python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 240 --csv --log-file gpurun_out/r1_step_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
