STEPS=1 python scripts/run_conv.py > gpurun_out/plain_conv.log 2>&1 && cat gpurun_out/plain_conv.log &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/conv_launches.csv python scripts/run_conv.py > /dev/null 2>&1
