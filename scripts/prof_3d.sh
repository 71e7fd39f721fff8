STEPS=2 python scripts/bench_configs.py fno3d > gpurun_out/plain_3d.log 2>&1 &&
STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 200 --csv --log-file gpurun_out/fno3d_launches.csv python scripts/bench_configs.py fno3d > /dev/null 2>&1
