STEPS=2 python scripts/bench_configs.py unet1d > gpurun_out/plain_unet.log 2>&1 &&
STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 600 --csv --log-file gpurun_out/unet_launches.csv python scripts/bench_configs.py unet1d > /dev/null 2>&1
