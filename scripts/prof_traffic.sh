# DRAM bytes + duration of every slab / wgrad / r2c launch of one eagerly launched train step (after bench has run plainly)
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'slab_kernel|wgrad_kernel|r2c_kernel' -s 60 -c 24 --csv --log-file gpurun_out/r1_slab_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
