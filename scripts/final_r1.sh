set -x
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -c 600 gpurun_out/bench_r1_n1.json; tail -2 gpurun_out/bench_r1_n1.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 240 --csv --log-file gpurun_out/r1_step_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
python scripts/run_conv.py > gpurun_out/plain_conv.log 2>&1 && cat gpurun_out/plain_conv.log &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_rows|conv_wgrad_shift' -s 7 -c 3 -o gpurun_out/r1_conv_pair python scripts/run_conv.py > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
timeout 600 python scripts/bench_configs.py > gpurun_out/r1_configs.jsonl 2> gpurun_out/r1_configs.err; cut -c1-200 gpurun_out/r1_configs.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
