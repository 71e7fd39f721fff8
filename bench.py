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

WORKLOAD = "ClassicFNO 2D periodic, 256x256 grid, hidden 64, 16 modes/axis, 4 blocks, global batch 64 (BASELINE.json configs[3])"
GRID, HIDDEN, MODES, BLOCKS, GLOBAL_BATCH = 256, 64, 16, 4, 64
METRIC = "ClassicFNO 2D 256^2 train samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PDEQX_PRECISION", "tf32"), choices=["tf32", "fp32"])
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH)
    ap.add_argument("--cpu-sample", type=int, default=2, help="samples per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's train step (NumPy; pocketfft + BLAS on the host cores)
# ------------------------------------------------------------------------------------------------
def cpu_train_step_rate(sample, steps, warmup):
    import numpy as np
    from oracle import nn as onn
    rng = np.random.default_rng(0)
    p = onn.init_fno(rng, 2, 1, 1, hidden_channels=HIDDEN, num_modes=MODES, num_blocks=BLOCKS)
    state = onn.adam_init(p)
    x = rng.standard_normal((sample, 1, GRID, GRID)).astype(np.float32)
    y = rng.standard_normal((sample, 1, GRID, GRID)).astype(np.float32)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, g = onn.fno_loss_and_grad(p, x, y, (MODES, MODES), BLOCKS)
        g = {k: v.reshape(p[k].shape) for k, v in g.items()}
        p, state = onn.adam_step(p, g, state)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    try:
        from threadpoolctl import threadpool_info
        cores = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        cores = len(os.sched_getaffinity(0))
    return sample / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warmup = 1 if args.warmup > 0 else 0
    rate, dt, cores = cpu_train_step_rate(args.cpu_sample, steps, warmup)
    sample = (f"{args.cpu_sample} of the {args.global_batch} samples per step, {steps} timed steps; NumPy port of the "
              "reference train step (rfftn/einsum/irfftn forward, analytic reverse pass, Adam); jax/equinox are not "
              "installable in this image so the reference itself cannot run")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "timing": "host perf_counter, CPU only"},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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

    model = pdeqx.ClassicFNO(2, 1, 1, hidden_channels=HIDDEN, num_modes=MODES, num_blocks=BLOCKS, key=0)
    step = TrainStep(model, lr=3e-4)
    rng = np.random.default_rng(1234)  # same global batch on every rank, sharded by sample
    xg = rng.standard_normal((args.global_batch, 1, GRID, GRID), dtype=np.float32)
    yg = rng.standard_normal((args.global_batch, 1, GRID, GRID), dtype=np.float32)
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

    # ---- device-resident arm ----
    for _ in range(max(args.warmup, 3)):
        step(xd, yd)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.pdeqx_launch_count()
    L.pdeqx_timing_enable(1)
    ms = timed(lambda: step(xd, yd), args.steps)
    launches = L.pdeqx_launch_count() - launches0
    slab_ms, slab_n = _lib.C.c_double(), _lib.C.c_int64()
    _lib.check(L.pdeqx_timing_read(_lib.K_SPECTRAL_SLAB, _lib.C.byref(slab_ms), _lib.C.byref(slab_n)))
    L.pdeqx_timing_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    value = args.global_batch * args.steps / (ms * 1e-3)

    # ---- end-to-end arm: host buffers in, loss out, every step ----
    xin, yin = torch.empty_like(xd), torch.empty_like(yd)
    loss_host = torch.zeros(1).pin_memory()

    def e2e_step():
        xin.copy_(xh, non_blocking=True)
        yin.copy_(yh, non_blocking=True)
        loss = step(xin, yin)
        loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller owns the loss value before the next step

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = args.global_batch * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (read x once + write y once [+ saved z]) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    n_pix = GRID * GRID
    # per launch: x [B,C,N] read, y [B,C,N] written, z [B,C,H,2m] read, tables + bypass weights (negligible)
    slab_bytes = 4 * per * HIDDEN * (2 * n_pix + GRID * 2 * MODES) + 4 * (HIDDEN * HIDDEN + HIDDEN + 2 * MODES * GRID)
    roofline = None
    if slab_n.value > 0:
        avg_s = slab_ms.value * 1e-3 / slab_n.value
        achieved = slab_bytes / avg_s / 1e9
        roofline = {"bound": "hbm", "kernel": "spectral block last stage (c2r DFT + 1x1 bypass + bias + gelu)",
                    "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                    "algorithmic_bytes_per_launch": slab_bytes, "avg_launch_us": avg_s * 1e6,
                    "launches_timed": slab_n.value, "share_of_step": slab_ms.value / ms, "traffic": None}

    # ---- PhysicsConv roofline (BASELINE.json configs[2] layer: 3x3, 64 -> 64, 128x128, batch 128, periodic) ----
    physics = None
    if world == 1:
        conv = pdeqx.PhysicsConv(2, 64, 64, 3, key=1, boundary_mode="periodic")
        xc = torch.randn(128, 64, 128, 128, device="cuda")
        yc = torch.empty_like(xc)
        desc, _ = conv._desc(xc)

        def conv_fwd():
            _lib.check(L.pdeqx_conv_fwd(desc, _lib.ptr(xc), _lib.ptr(conv.weight), _lib.ptr(conv.bias), None, _lib.ACT_RELU,
                                        _lib.ptr(yc), _lib.stream()))
        for _ in range(3):
            conv_fwd()
        reps = 10
        ms_c = timed(conv_fwd, reps) / reps
        conv_bytes = 4 * 128 * (64 + 64) * 128 * 128 + 4 * 64 * (64 * 9 + 1)
        physics = {"bound": "hbm (at the TF32 ridge: 144 flop/B)", "kernel": "PhysicsConv 3x3 64->64 fwd + bias + relu, 128x128, batch 128, periodic",
                   "achieved": conv_bytes / (ms_c * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                   "frac": conv_bytes / (ms_c * 1e-3) / 1e9 / hbm_peak, "avg_launch_us": ms_c * 1e3,
                   "tflops": 2 * 64 * 64 * 9 * 128 * 128 * 128 / (ms_c * 1e-3) / 1e12,
                   "tcgen05_passes": int(L.pdeqx_conv_tcgen05_passes(desc))}
        del xc, yc

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, dt, cores = cpu_train_step_rate(args.cpu_sample, 2, 1)
        cpu = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_sample} of the {args.global_batch} samples per step, 2 timed steps ({dt:.1f} s each); "
                         "NumPy port of the reference train step (jax/equinox not installable here)"}

    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "tf32 (fp32 accumulate, fp32 storage)" if args.precision == "tf32" else "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": args.global_batch, "per_gpu_batch": per,
                   "parallelism": f"dp{world} (batch-sharded, one NCCL all-reduce of the 67.2 MB gradient arena per step)",
                   "precision": args.precision,
                   "l2": "inputs larger than L2: every activation tensor is %.0f MB per rank" % (per * HIDDEN * n_pix * 4 / 1e6)},
        "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(xh.numel() * 4 + yh.numel() * 4) * world, "d2h_bytes_per_step": 4 * world},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "physics_conv": physics,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
