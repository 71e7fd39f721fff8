#!/bin/bash
# per-launch duration / DRAM bytes / shared-store conflicts of the channel-mix kernels of one eager train step (cold caches)
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,smsp__inst_executed_op_shared_st.sum \
  --clock-control none ${NCU_CACHE:-} -k regex:mode_gemm -s 36 -c 12 --csv --log-file gpurun_out/modemix.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph "$@" > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/modemix.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); idi=h.index('ID'); gi=h.index('Grid Size')
d={}
for r in rows[1:]: d.setdefault((int(r[idi]), r[ki][5:45], r[gi]),{})[r[mi].split('__')[1][:28]]=r[vi]
for k,v in sorted(d.items()): print(k, v)
P
