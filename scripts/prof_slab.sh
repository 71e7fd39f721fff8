ITERS=1 python scripts/run_block.py && ITERS=1 ncu --set full --clock-control none --import-source on -k regex:slab_kernel -s 3 -c 3 -o gpurun_out/slab_v2 python scripts/run_block.py > gpurun_out/ncu_slab.log 2>&1
tail -2 gpurun_out/ncu_slab.log
python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32_v2.json 2> gpurun_out/bench_tf32_v2.err; cat gpurun_out/bench_tf32_v2.json | head -c 600
