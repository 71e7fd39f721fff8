"""Markdown summary of an .ncu-rep (read on the CPU box): per captured launch, the metrics the roofline discussion needs.
    python scripts/ncu_summary.py gpurun_out/r1_block.ncu-rep > profiles/r1_block_ncu.md"""
import csv, io, subprocess, sys, re
WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("smsp__cycles_active.avg", "SMSP cycles active"),
]
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(head)}
    print(f"## {rep}\n")
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
        print(f"### {name}\n")
        print("| metric | value |\n|---|---|")
        for key, label in WANT:
            if key in idx and r[idx[key]] != "":
                print(f"| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |")
        print()
