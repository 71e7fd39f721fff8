set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 680 -c 240 --csv --log-file gpurun_out/r1_step_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
ITERS=1 python scripts/run_block.py > gpurun_out/plain_block.log 2>&1 &&
ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'slab_kernel|wgrad_kernel' -s 3 -c 4 -o gpurun_out/r1_block python scripts/run_block.py > gpurun_out/ncu_block.log 2>&1
tail -2 gpurun_out/ncu_block.log
ITERS=1 timeout 900 ncu --set full --clock-control none -k regex:'r2c_kernel|c2c_kernel|mode_gemm' -s 8 -c 8 -o gpurun_out/r1_block_small python scripts/run_block.py > gpurun_out/ncu_block2.log 2>&1
tail -2 gpurun_out/ncu_block2.log
python scripts/run_conv.py > gpurun_out/plain_conv.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_rows|conv_wgrad_kernel' -s 8 -c 3 -o gpurun_out/r1_conv python scripts/run_conv.py > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
ls -la gpurun_out/*.ncu-rep
