python bench.py --steps 3 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32_v3.json 2> gpurun_out/bench_tf32_v3.err && ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/step_launches.csv python bench.py --steps 3 --warmup 3 --precision tf32 --no-cpu-baseline > /dev/null 2>&1
head -c 400 gpurun_out/bench_tf32_v3.json; echo
