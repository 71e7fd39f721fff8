python scripts/run_conv.py && ncu --set full --clock-control none --import-source on -k regex:conv_rows_kernel -s 2 -c 1 -o gpurun_out/conv_v2 python scripts/run_conv.py > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
