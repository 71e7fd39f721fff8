#!/usr/bin/env python
"""bench.py -- ClassicFNO 2D 256^2 train-step throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision tf32|fp32] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one Adam training step (forward, reverse pass, gradient all-reduce, Adam) of
ClassicFNO(2, 1, 1, hidden 64, 16 modes/axis, 4 blocks, gelu) on synthetic fields x, y ~ N(0,1) of
shape (64, 1, 256, 256): BASELINE.json configs[3], global batch 64 sharded by sample over the N ranks.

  value : samples/s, whole job, inputs already resident in HBM, CUDA events, max over ranks
  e2e   : same metric through the public API with HOST (pinned) inputs: every step copies its shard
          host->device and reads the loss back, both inside the timed region
  roofline : the dominant kernel (last stage of the spectral block: inverse DFT along the contiguous
          axis + 1x1 bypass + bias + gelu), algorithmic bytes / its CUDA-event time vs measured HBM peak
  cpu_baseline : the NumPy oracle (CPU port of the reference; JAX is not installable here) on a bounded
          sample of the same workload, on this box's host cores (rank 0, N = 1 only)

--impl reference times that CPU port alone (the reference itself cannot be imported: no jax/equinox
in the image, DESIGN.md "Reference arm").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: NCCL's own banner / debug output goes to stderr
# (NCCL's version banner is a bare printf): file descriptor 1 is pointed at stderr for the whole process and the result
# line is written to a private duplicate of the original stdout.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())

# The workloads: BASELINE.json configs[3] (the metric's configuration, default) and configs[4] (--config fno3d).
CONFIGS = {
    "fno2d": dict(workload="ClassicFNO 2D periodic, 256x256 grid, hidden 64, 16 modes/axis, 4 blocks, global batch 64 (BASELINE.json configs[3])",
                  metric="ClassicFNO 2D 256^2 train samples/s", ndim=2, grid=256, hidden=64, modes=16, blocks=4, global_batch=64),
    "fno3d": dict(workload="ClassicFNO 3D periodic, 64x64x64 grid, hidden 32, 12 modes/axis, 4 blocks, global batch 32 (BASELINE.json configs[4])",
                  metric="ClassicFNO 3D 64^3 train samples/s", ndim=3, grid=64, hidden=32, modes=12, blocks=4, global_batch=32),
}
L2_NOTE = ("GPU arm: inputs larger than L2, no flush between steps (every activation tensor of a step is >= 134 MB per rank "
           "at every rank count up to 8; L2 is 126 MB)")


def config_block(args):
    """`config` of the JSON line: identical in both arms (the reference arm runs on the GPU arm's configuration)."""
    return {"workload": args.cfg["workload"], "global_batch": args.global_batch, "l2": L2_NOTE}


TRAFFIC_CSV = os.path.join(ROOT, "profiles", "slab_traffic.csv")


def kernel_source_sha():
    """sha256 over the kernel sources (csrc/ + the public header): an ncu capture under profiles/ is only quoted by the
    bench line if it was taken from exactly these sources (there is no .git on the GPU box)."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "pdequinox_b200", "csrc")
    files = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(ROOT, "include", "pdeqx_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def measured_slab_traffic(config, per_gpu_batch):
    """(bytes per slab launch, note) from profiles/slab_traffic.csv -- dram__bytes_read.sum + dram__bytes_write.sum of every
    tc::slab_kernel launch of ONE train step, captured by `scripts/profile_step.sh traffic` -- or (None, why) when the
    capture does not belong to this build / workload. The file starts with `# key=value` lines: source_sha, config,
    per_gpu_batch."""
    import csv
    try:
        lines = open(TRAFFIC_CSV).read().splitlines()
    except OSError:
        return None, "no capture: profiles/slab_traffic.csv is absent"
    meta = dict(ln[1:].strip().split("=", 1) for ln in lines if ln.startswith("#") and "=" in ln)
    if meta.get("source_sha") != kernel_source_sha():
        return None, f"stale capture: taken from kernel sources {meta.get('source_sha')}, this build is {kernel_source_sha()}"
    if meta.get("config") != config or int(meta.get("per_gpu_batch", -1)) != per_gpu_batch:
        return None, f"capture is for {meta.get('config')} at {meta.get('per_gpu_batch')} samples per GPU"
    rows = list(csv.reader(ln for ln in lines if ln.startswith('"')))
    head = rows[0]
    ki, mi, vi = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value")
    ii = head.index("ID")
    total, ids = 0.0, set()
    for r in rows[1:]:
        if ("slab_kernel" in r[ki] or "chain_kernel" in r[ki]) and r[mi] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(r[vi].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[r[head.index("Metric Unit")]]
            ids.add(r[ii])
    if not ids:
        return None, "capture holds no slab_kernel / chain_kernel launch"
    return total / len(ids), (f"ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the {len(ids)} slab_kernel / chain_kernel launches of one "
                              f"train step (profiles/slab_traffic.csv, kernel sources {meta['source_sha']})")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PDEQX_PRECISION", "tf32"), choices=["tf32", "fp32"])
    ap.add_argument("--config", default="fno2d", choices=sorted(CONFIGS), help="fno2d = BASELINE.json configs[3] (the metric), fno3d = configs[4]")
    ap.add_argument("--global-batch", type=int, default=0, help="0 = the configuration's own global batch")
    ap.add_argument("--cpu-sample", type=int, default=0, help="samples (= host threads) per CPU-baseline step; 0 = all cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap-allreduce", action="store_true", help="bucketed all-reduce on a side stream during the reverse pass")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "allreduce"],
                    help="gradient exchange of the N > 1 runs: peer-memory kernel (default where available) or NCCL all-reduce")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying the captured CUDA graph")
    ap.add_argument("--print-source-sha", action="store_true", help="print the kernel-source hash that keys profiles/slab_traffic.csv and exit")
    a = ap.parse_args()
    if a.print_source_sha:
        os.write(_RESULT_FD, (kernel_source_sha() + "\n").encode())
        raise SystemExit(0)
    a.cfg = CONFIGS[a.config]
    if a.global_batch <= 0:
        a.global_batch = a.cfg["global_batch"]
    return a


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's train step (NumPy; pocketfft + BLAS on the host cores)
# ------------------------------------------------------------------------------------------------
def cpu_workers():
    """Host threads of the CPU arm: one sample per thread (the path shards by sample), bounded by memory (~4 GB each)."""
    cores = len(os.sched_getaffinity(0))
    try:
        import psutil
        cores = min(cores, max(1, int(psutil.virtual_memory().available // (4 << 30))))
    except Exception:
        pass
    return max(1, min(cores, 32))


def cpu_train_step_rate(cfg, samples, workers, steps, warmup):
    """Train steps of the oracle port on `samples` samples each, in waves of `workers` host threads, one sample per thread
    (NumPy's pocketfft, einsum and BLAS calls release the GIL; BLAS itself is pinned to one thread per call), gradients
    averaged over the samples, one Adam update per step. Returns (samples/s, seconds per step)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import nn as onn
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    D, grid, modes, blocks = cfg["ndim"], cfg["grid"], cfg["modes"], cfg["blocks"]
    rng = np.random.default_rng(0)
    p = onn.init_fno(rng, D, 1, 1, hidden_channels=cfg["hidden"], num_modes=modes, num_blocks=blocks)
    state = onn.adam_init(p)
    shape = (samples, 1) + (grid,) * D
    x = rng.standard_normal(shape).astype(np.float32)
    y = rng.standard_normal(shape).astype(np.float32)

    def one(i):
        return onn.fno_loss_and_grad(p, x[i:i + 1], y[i:i + 1], (modes,) * D, blocks)

    times = []
    limiter = threadpool_limits(limits=1) if (threadpool_limits and workers > 1) else None
    try:
        with ThreadPoolExecutor(max_workers=workers) as pool:
            for it in range(warmup + steps):
                t0 = time.perf_counter()
                acc = None
                for w0 in range(0, samples, workers):  # one wave = one sample per host thread; summed as it completes
                    for _, g in pool.map(one, range(w0, min(w0 + workers, samples))):
                        acc = dict(g) if acc is None else {k: acc[k] + g[k] for k in acc}
                g = {k: acc[k].reshape(p[k].shape) / samples for k in acc}
                p, state = onn.adam_step(p, g, state)
                if it >= warmup:
                    times.append(time.perf_counter() - t0)
    finally:
        if limiter is not None and hasattr(limiter, "restore_original_limits"):
            limiter.restore_original_limits()
    dt = sum(times) / len(times)
    return samples / dt, dt


# seconds of host work one sample of a train step costs on one thread (measured on the round-1 GPU box: 16 samples on 16
# threads per 1.9-2.0 s step); only used to size the bounded sample of the reference arm before anything is timed
CPU_SECONDS_PER_SAMPLE = {"fno2d": 2.0, "fno3d": 6.0}
REFERENCE_BUDGET_S = 270.0


def run_reference(args):
    """--impl reference: the CPU port of the reference's train step on this box's host cores, on the configuration,
    --steps and --warmup of the GPU arm. A step covers the full global batch whenever (steps + warmup) such steps fit
    REFERENCE_BUDGET_S of host time; otherwise the largest whole number of waves that does (stated in `sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.cfg
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    workers = args.cpu_sample if args.cpu_sample > 0 else cpu_workers()
    workers = min(workers, args.global_batch)
    wave_s = CPU_SECONDS_PER_SAMPLE[args.config]
    waves_full = -(-args.global_batch // workers)
    waves = max(1, min(waves_full, int(REFERENCE_BUDGET_S / ((steps + warmup) * wave_s))))
    samples = min(args.global_batch, waves * workers)
    rate, dt = cpu_train_step_rate(cfg, samples, workers, steps, warmup)
    sample = (f"{samples} of the {args.global_batch} samples per step in {waves} wave(s) of {workers} host threads (one sample per "
              f"thread), {steps} timed steps of {dt:.1f} s after {warmup} warm-up; NumPy port of the reference train step "
              "(rfftn/einsum/irfftn forward, analytic reverse pass, Adam); jax/equinox are not installable in this image so the "
              "reference itself cannot run")
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": rate, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(args),
        "details": {"timing": "host perf_counter, CPU only"},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def shutdown(world, graphed):
    """Tear down in an order that cannot hang: captured graphs that contain NCCL kernels are released before the
    communicator, and a watchdog ends the process if the communicator teardown itself blocks (the result line has
    already been printed and flushed by then)."""
    import torch
    import torch.distributed as dist
    if graphed is not None:
        graphed.graph.reset()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        threading.Timer(20.0, lambda: os._exit(0)).start()
        try:
            dist.barrier()
            dist.destroy_process_group()
        finally:
            os._exit(0)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import pdequinox_b200 as pdeqx
    from pdequinox_b200 import _lib
    from pdequinox_b200.train import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.gpus != world and rank == 0:
        print(f"# note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    if args.global_batch % world:
        raise SystemExit(f"global batch {args.global_batch} is not divisible by {world} ranks")
    per = args.global_batch // world
    L = _lib.lib()
    pdeqx.set_precision(args.precision)

    cfg = args.cfg
    D, GRID, HIDDEN, MODES, BLOCKS = cfg["ndim"], cfg["grid"], cfg["hidden"], cfg["modes"], cfg["blocks"]
    spatial = (GRID,) * D
    model = pdeqx.ClassicFNO(D, 1, 1, hidden_channels=HIDDEN, num_modes=MODES, num_blocks=BLOCKS, key=0)
    step = TrainStep(model, lr=3e-4, overlap=args.overlap_allreduce, exchange=args.exchange)
    rng = np.random.default_rng(1234)  # same global batch on every rank, sharded by sample
    xg = rng.standard_normal((args.global_batch, 1) + spatial, dtype=np.float32)
    yg = rng.standard_normal((args.global_batch, 1) + spatial, dtype=np.float32)
    xh = torch.from_numpy(xg[rank * per:(rank + 1) * per]).pin_memory()
    yh = torch.from_numpy(yg[rank * per:(rank + 1) * per]).pin_memory()
    xd, yd = xh.cuda(), yh.cuda()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up (eager), then capture the whole step (forward, reverse pass, all-reduce, Adam) into one CUDA graph ----
    for _ in range(max(args.warmup, 3)):
        step(xd, yd)
    torch.cuda.synchronize()
    launches0 = L.pdeqx_launch_count()
    step(xd, yd)
    launches_per_step = L.pdeqx_launch_count() - launches0
    graphed = None
    if not args.no_graph:
        try:
            from pdequinox_b200.train import GraphedTrainStep
            graphed = GraphedTrainStep(step, xd, yd, warmup=1)
        except Exception as exc:  # capture is an optimisation of the launch path, never a different computation
            print(f"# CUDA-graph capture failed ({type(exc).__name__}: {exc}); timing the eager launch path", file=sys.stderr)
            graphed = None
    run_step = (lambda: graphed()) if graphed is not None else (lambda: step(xd, yd))
    # clock settling: a fresh box boosts for the first tenths of a second of load and then drops to its sustained state; the
    # W warm-up steps above are too short to reach it (20 steps of 7 ms measured 3 % slower when they came first than the same
    # steps half a second later). Replay untimed steps for ~0.4 s so that the timed region sees the sustained clocks.
    # (the number of steps is agreed between the ranks: they synchronise inside every step)
    # The clock sampler starts BEFORE the settling steps: launching nvidia-smi (NVML initialisation) next to running kernels
    # slowed the first tenths of a second of them by ~3 % (device-timed value 6.45 ms/step against 6.26 ms in the end-to-end
    # arm of the same run, which started after the sampler had stopped); its samples cover the settling load and the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    settle_steps = max(2, min(400, int(400.0 / max(timed(run_step, 3) / 3.0, 0.05))))
    for _ in range(settle_steps):
        run_step()
    torch.cuda.synchronize()
    settle_steps += 3

    # ---- device-resident arm ----
    ms = timed(run_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = launches_per_step * args.steps
    value = args.global_batch * args.steps / (ms * 1e-3)

    # ---- end-to-end arm: host buffers in, loss out, every step ----
    xin, yin = torch.empty_like(xd), torch.empty_like(yd)
    loss_host = torch.zeros(1).pin_memory()

    if graphed is not None:
        graphed.prefetch(xh, yh)  # input pipeline primed: every step below issues the copy of the batch after it

    def e2e_step():
        if graphed is not None:
            loss = graphed()              # consumes the staged batch (waits for its host -> device copy), one graph launch
            graphed.prefetch(xh, yh)      # next batch: pinned host -> device on the copy stream, under this step's kernels
        else:
            xin.copy_(xh, non_blocking=True)
            yin.copy_(yh, non_blocking=True)
            loss = step(xin, yin)
        loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller owns the loss value before the next step

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = args.global_batch * args.steps / (ms_e2e * 1e-3)

    # ---- per-kernel durations: the same steps launched eagerly, each tagged launch bracketed by CUDA events on its stream ----
    probe_steps = min(args.steps, 3)
    L.pdeqx_timing_enable(1)
    ms_probe = timed(lambda: step(xd, yd), probe_steps)
    slab_ms, slab_n, slab_bytes_total = _lib.C.c_double(), _lib.C.c_int64(), _lib.C.c_double()
    _lib.check(L.pdeqx_timing_read_ex(_lib.K_SPECTRAL_SLAB, _lib.C.byref(slab_ms), _lib.C.byref(slab_n), _lib.C.byref(slab_bytes_total)))
    exchange_us = None
    if step.peer is not None:
        ex_ms, ex_n = _lib.C.c_double(), _lib.C.c_int64()
        _lib.check(L.pdeqx_timing_read(_lib.K_DP_EXCHANGE, _lib.C.byref(ex_ms), _lib.C.byref(ex_n)))
        exchange_us = ex_ms.value * 1e3 / max(ex_n.value, 1)
    L.pdeqx_timing_enable(0)

    if step.peer is not None:
        step.peer.check()  # every rank arrived at every flag barrier of the exchange
    if rank != 0:
        shutdown(world, graphed)
        return

    # ---- roofline of the dominant kernel ----
    # tc::slab_kernel = the last stage of every pass over a spectral block (forward, pre-activation cotangent, data gradient).
    # Its launcher states the ALGORITHMIC bytes of each launch -- every tensor the launch reads or writes, once (SURVEY.md 8d:
    # x in, y out, plus the saved z / the cotangent where the variant has them; the folded ends of the network read or write a
    # single-channel field instead) -- and the library brackets every launch with CUDA events on its stream.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_source = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    n_pix = GRID ** D
    roofline = None
    if slab_n.value > 0:
        avg_s = slab_ms.value * 1e-3 / slab_n.value
        slab_bytes = slab_bytes_total.value / slab_n.value
        achieved = slab_bytes / avg_s / 1e9
        traffic, traffic_note = measured_slab_traffic(args.config, per)
        roofline = {"bound": "hbm", "kernel": "tc::slab_kernel / tc::chain_kernel, last stage of every pass over a spectral block (c2r DFT + 1x1 bypass "
                              "+ bias + gelu forward; chained data-gradient x gelu' reverse), mean over their %d launches per step" % (slab_n.value // probe_steps),
                    "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "peak_source": peak_source,
                    "algorithmic_bytes_per_launch": slab_bytes, "avg_launch_us": avg_s * 1e6,
                    "launches_timed": slab_n.value, "share_of_step": slab_ms.value / ms_probe,
                    "timed_in": f"{probe_steps} eagerly launched steps after the graph-replayed region (events cannot bracket "
                                "kernels inside a graph replay); share_of_step is relative to that pass",
                    "traffic": traffic, "traffic_source": traffic_note}

    # ---- op-level roofline: one whole ClassicSpectralBlock (hidden -> hidden, this configuration's grid and per-GPU batch),
    # forward and reverse pass, every kernel of the op (r2c, leading-axis DFTs, channel mix, slab; cotangent, weight
    # gradients, adjoint chain, data gradient) inside the CUDA-event bracket, against SURVEY.md 8(d)'s algorithmic bytes:
    #   forward: read x, write y, read the complex weights;  reverse: read gy, write gx, read w, write gw (+ saved modes)
    spectral_block = None
    if world == 1:
        from pdequinox_b200.train import ParameterArena
        blk = pdeqx.ClassicSpectralBlock(D, HIDDEN, HIDDEN, activation=model.blocks[0].activation, num_modes=MODES, key=2)
        ParameterArena(blk)
        xb = torch.randn((per, HIDDEN) + spatial, device="cuda")
        gyb = torch.randn((per, HIDDEN) + spatial, device="cuda")
        tape = pdeqx._module.Tape()

        reps, ms_f, ms_b = 5, 0.0, 0.0
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for it in range(2 + reps):
            tape.stack.clear()
            ev[0].record()
            blk._fwd(xb, tape)
            ev[1].record()
            blk._bwd(gyb, tape, True)
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                ms_f += ev[0].elapsed_time(ev[1]) / reps
                ms_b += ev[1].elapsed_time(ev[2]) / reps
        G = 2 ** (D - 1)
        w_bytes = 8 * G * HIDDEN * HIDDEN * MODES ** D
        act_bytes = 4 * per * (HIDDEN + HIDDEN) * n_pix
        modes_saved = 8 * per * HIDDEN * G * MODES ** D
        fwd_bytes = act_bytes + w_bytes + 4 * HIDDEN * (HIDDEN + 1)
        bwd_bytes = act_bytes + 2 * w_bytes + modes_saved
        spectral_block = {
            "op": f"ClassicSpectralBlock {HIDDEN}->{HIDDEN}, grid {'x'.join(map(str, spatial))}, {MODES} modes/axis, batch {per} (spectral conv + "
                  "1x1 bypass + bias + activation as one op; every kernel of the op inside the bracket)",
            "bound": "hbm", "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_source,
            "fwd": {"us": ms_f * 1e3, "algorithmic_bytes": fwd_bytes, "achieved": fwd_bytes / (ms_f * 1e-3) / 1e9,
                    "frac": fwd_bytes / (ms_f * 1e-3) / 1e9 / hbm_peak},
            "bwd": {"us": ms_b * 1e3, "algorithmic_bytes": bwd_bytes, "achieved": bwd_bytes / (ms_b * 1e-3) / 1e9,
                    "frac": bwd_bytes / (ms_b * 1e-3) / 1e9 / hbm_peak,
                    "note": "SURVEY.md 8(d) counts gy in + gx out only; the reverse pass also has to read x (bypass weight gradient, "
                            "recomputed pre-activation): with that tensor the floor is 1.5x these bytes"},
        }
        del xb, gyb, tape, blk
        torch.cuda.empty_cache()

    # ---- PhysicsConv roofline (BASELINE.json configs[2] layer: 3x3, 64 -> 64, 128x128, batch 128, periodic) ----
    physics = None
    if world == 1:
        conv = pdeqx.PhysicsConv(2, 64, 64, 3, key=1, boundary_mode="periodic")
        xc = torch.randn(128, 64, 128, 128, device="cuda")
        yc = torch.empty_like(xc)
        desc, _ = conv._desc(xc)
        nws = L.pdeqx_conv_workspace_bytes(desc)
        ws = torch.empty(nws // 4 + 1, device="cuda")

        def conv_fwd():
            _lib.check(L.pdeqx_conv_fwd(desc, _lib.ptr(xc), _lib.ptr(conv.weight), _lib.ptr(conv.bias), None, _lib.ACT_RELU,
                                        _lib.ptr(yc), _lib.ptr(ws), nws, _lib.stream()))
        for _ in range(3):
            conv_fwd()
        reps = 100  # ~25 ms back to back: long enough for the clocks to settle under this kernel
        ms_c = timed(conv_fwd, reps) / reps
        conv_bytes = 4 * 128 * (64 + 64) * 128 * 128 + 4 * 64 * (64 * 9 + 1)
        tflops = 2 * 64 * 64 * 9 * 128 * 128 * 128 / (ms_c * 1e-3) / 1e12
        # 144 flop per algorithmic byte puts this layer ON the TF32 ridge: the HBM fraction is the north star's target figure, the
        # tensor fraction says what actually bounds the kernel (TF32 dense peak taken as half of the measured bf16 figure, the
        # sustained one: the probe runs the kernel back to back)
        tf32_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 2250.0 * 0.6)) / 2.0
        physics = {"bound": "tensor (tf32) -- at the ridge: 144 flop/B", "kernel": "PhysicsConv 3x3 64->64 fwd + bias + relu, 128x128, batch 128, periodic",
                   "achieved": conv_bytes / (ms_c * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                   "frac": conv_bytes / (ms_c * 1e-3) / 1e9 / hbm_peak, "avg_launch_us": ms_c * 1e3,
                   "tflops": tflops, "tensor_peak_tflops": tf32_peak, "tensor_frac": tflops / tf32_peak,
                   "tensor_peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (derived: tf32 runs at half the bf16 rate)" if peaks else "fallback",
                   "tcgen05_passes": int(L.pdeqx_conv_tcgen05_passes(desc))}
        # the two reverse passes of the same layer (same algorithmic bytes: two activation-sized tensors each; same flops)
        gxc = torch.empty_like(xc)
        gwc, gbc = torch.empty_like(conv.weight), torch.empty_like(conv.bias)

        def conv_bwd_data():
            _lib.check(L.pdeqx_conv_bwd_data(desc, _lib.ptr(yc), _lib.ptr(conv.weight), None, None, _lib.ptr(gxc), _lib.ptr(ws), nws,
                                             _lib.stream()))

        def conv_bwd_filter():
            _lib.check(L.pdeqx_conv_bwd_filter(desc, _lib.ptr(xc), _lib.ptr(yc), _lib.ptr(gwc), _lib.ptr(gbc), _lib.ptr(ws), nws,
                                               _lib.stream()))
        for name, fn in (("bwd_data", conv_bwd_data), ("bwd_filter", conv_bwd_filter)):
            for _ in range(3):
                fn()
            ms_r = timed(fn, 50) / 50
            physics[name] = {"avg_us": ms_r * 1e3, "frac": conv_bytes / (ms_r * 1e-3) / 1e9 / hbm_peak,
                             "tensor_frac": 2 * 64 * 64 * 9 * 128 * 128 * 128 / (ms_r * 1e-3) / 1e12 / tf32_peak}
        physics["bwd_filter"]["note"] = ("filter + bias gradient: tcgen05 with the column-shifted cotangent operand in tensor memory, "
                                         "partial sums of every CTA reduced by a second kernel (inside the bracket)")
        del xc, yc, gxc

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        workers = args.cpu_sample if args.cpu_sample > 0 else cpu_workers()
        workers = min(workers, args.global_batch)
        rate, dt = cpu_train_step_rate(cfg, workers, workers, 5, 1)  # ~12 s of CPU work on a 16-core host
        cpu = {"value": rate, "unit": "samples/s", "cores": workers, "kind": "port",
               "sample": f"{workers} of the {args.global_batch} samples per step (one per host thread), 5 timed steps "
                         f"({dt:.1f} s each); NumPy port of the reference train step (jax/equinox not installable here)"}

    line = {
        "metric": cfg["metric"], "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "tf32 (fp32 accumulate, fp32 storage)" if args.precision == "tf32" else "f32",
        "data": "synthetic",
        "config": config_block(args),
        "details": {"per_gpu_batch": per,
                   "parallelism": f"dp{world} (batch-sharded; exchange of the %.1f MB gradient arena per step: %s)" % (
                       step.arena.size * 4 / 1e6,
                       "none (one rank)" if world == 1 else
                       "reduce-scatter + Adam on the rank's shard + all-gather in one kernel over NVLink peer memory" if step.peer is not None
                       else "one NCCL all-reduce + replicated Adam"),
                   "launch": "one CUDA graph replay per step" if graphed is not None else "eager kernel launches",
                   "precision": args.precision,
                   "untimed_steps_before_timing": max(args.warmup, 3) + 1 + settle_steps,
                   "exchange_us_per_step_rank0": exchange_us,  # peer-memory exchange incl. the wait for the slowest rank (eager probe steps)
                   "activation_tensor_mb_per_rank": per * HIDDEN * n_pix * 4 / 1e6},
        "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(xh.numel() * 4 + yh.numel() * 4) * world, "d2h_bytes_per_step": 4 * world},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "spectral_block": spectral_block,
        "physics_conv": physics,
        "cpu_baseline": cpu,
    }
    emit(line)
    shutdown(world, graphed)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
