#!/bin/bash
# ncu passes over one bench.py train step (run on the GPU box through gpurun; every pass only after the same command has
# exited 0 without ncu; --clock-control none; numbers printed under ncu are never bench values).
#
#   scripts/profile_step.sh launches [bench args]   per-launch durations of ~3 eagerly launched steps -> gpurun_out/step_launches.csv
#   scripts/profile_step.sh traffic  [bench args]   DRAM bytes + duration of every slab_kernel / chain_kernel launch of one step
#                                                   -> gpurun_out/slab_traffic.csv (header: kernel-source hash, config, batch);
#                                                   copy to profiles/slab_traffic.csv: bench.py quotes it as roofline.traffic
#   scripts/profile_step.sh full KERNEL_REGEX [bench args]   one `--set full` capture of the first 12 matching launches after
#                                                   the warm-up steps -> gpurun_out/full_<regex>.ncu-rep
# SLAB_PER_STEP (default 9) = slab_kernel + chain_kernel launches per train step of the build under test (the launches bench.py's
# roofline block averages: forward x4, rank-one cotangent, chained reverse passes x3, lifting reduction).
set -u
mode=$1; shift
out=gpurun_out; mkdir -p $out
N=${SLAB_PER_STEP:-9}
case $mode in
launches)
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph "$@" > $out/plain_bench.log 2>&1 || { tail -5 $out/plain_bench.log; exit 1; }
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-700} -c ${COUNT:-260} --csv --log-file $out/step_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph "$@" > /dev/null 2>&1
  python scripts/launch_summary.py $out/step_launches.csv > $out/step_launches_summary.txt; cat $out/step_launches_summary.txt ;;
traffic)
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph "$@" > $out/plain_bench.log 2>&1 || { tail -5 $out/plain_bench.log; exit 1; }
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'slab_kernel|chain_kernel' \
      -s $((3 * N)) -c $N --csv --log-file $out/slab_traffic.raw.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph "$@" > /dev/null 2>&1
  cfg=fno2d; per=64
  prev=""; for a in "$@"; do [ "$prev" = "--config" ] && cfg=$a; [ "$prev" = "--global-batch" ] && per=$a; prev=$a; done
  [ "$cfg" = "fno3d" ] && [ "$per" = "64" ] && per=32
  { echo "# source_sha=$(python bench.py --print-source-sha)"; echo "# config=$cfg"; echo "# per_gpu_batch=$per";
    echo "# command=scripts/profile_step.sh traffic $*"; grep '^"' $out/slab_traffic.raw.csv; } > $out/slab_traffic.csv
  head -4 $out/slab_traffic.csv ;;
full)
  regex=$1; shift
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph "$@" > $out/plain_bench.log 2>&1 || { tail -5 $out/plain_bench.log; exit 1; }
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$regex -s ${SKIP:-36} -c ${COUNT:-12} -o $out/full_${regex//[^a-zA-Z0-9_]/_} -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph "$@" > $out/ncu_full.log 2>&1; tail -3 $out/ncu_full.log ;;
*) echo "unknown mode $mode"; exit 2 ;;
esac
