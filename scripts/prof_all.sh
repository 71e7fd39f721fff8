set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 240 --csv --log-file gpurun_out/r1_step_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
ITERS=1 python scripts/run_block.py > gpurun_out/plain_block.log 2>&1 &&
ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'slab_kernel|wgrad_kernel' -s 3 -c 4 -o gpurun_out/r1_block python scripts/run_block.py > gpurun_out/ncu_block.log 2>&1
tail -2 gpurun_out/ncu_block.log
python scripts/run_conv.py > gpurun_out/plain_conv.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_rows -s 4 -c 1 -o gpurun_out/r1_conv python scripts/run_conv.py > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
ls -la gpurun_out/*.ncu-rep
