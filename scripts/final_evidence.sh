#!/bin/bash
# One-GPU evidence pass of a round (run on the GPU box through gpurun): GPU tests, smoke, bench lines, DRAM traffic of the
# slab / chain launches (quoted by bench.py as roofline.traffic), launch list, CUPTI timelines. Everything lands in gpurun_out/.
set -u
T=${TAG:-r4_final}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; tail -2 $O/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
scripts/profile_step.sh traffic > $O/${T}_traffic.log 2>&1; cp $O/slab_traffic.csv profiles/slab_traffic.csv; head -3 $O/slab_traffic.csv
timeout 600 python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err; cut -c1-300 $O/${T}_bench_n1.json
timeout 300 python bench.py --config fno3d --no-cpu-baseline > $O/${T}_bench_n1_fno3d.json 2>/dev/null
timeout 300 python bench.py --precision fp32 --no-cpu-baseline --steps 5 > $O/${T}_bench_n1_fp32.json 2>/dev/null
for b in 64 8; do timeout 300 python scripts/graph_gaps.py $b 2>&1 | grep -v Warning | grep -v _warn_once > $O/${T}_graph_timeline_b$b.txt; done
timeout 600 python scripts/bench_configs.py 2>/dev/null | grep '^{' > $O/${T}_configs_1gpu.jsonl; cat $O/${T}_configs_1gpu.jsonl | cut -c1-200
COUNT=216 scripts/profile_step.sh launches > /dev/null 2>&1; cp $O/step_launches.csv $O/${T}_step_launches_tf32.csv; cp $O/step_launches_summary.txt $O/${T}_step_launches_tf32_summary.txt; head -12 $O/step_launches_summary.txt
