set -x
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'slab_kernel|wgrad_kernel' -s 15 -c 15 -o gpurun_out/r1_step_slab python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_step_slab.log 2>&1
tail -2 gpurun_out/ncu_step_slab.log
python scripts/run_conv.py > gpurun_out/plain_conv.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_rows' -s 8 -c 2 -o gpurun_out/r1_conv_pair python scripts/run_conv.py > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
timeout 600 python scripts/bench_configs.py > gpurun_out/r1_configs.jsonl 2> gpurun_out/r1_configs.err; cat gpurun_out/r1_configs.jsonl | cut -c1-220
