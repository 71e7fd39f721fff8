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

WORKLOAD = "ClassicFNO 2D periodic, 256x256 grid, hidden 64, 16 modes/axis, 4 blocks, global batch 64 (BASELINE.json configs[3])"
GRID, HIDDEN, MODES, BLOCKS, GLOBAL_BATCH = 256, 64, 16, 4, 64
METRIC = "ClassicFNO 2D 256^2 train samples/s"
# dram__bytes_read.sum + dram__bytes_write.sum of tc::slab_kernel at B = 64 per GPU, summed over the 12 launches of one train
# step (profiles/r1_slab_traffic.csv: forward 6.86 GB, cotangent passes 11.13 GB, data gradients 7.95 GB) and averaged per launch,
# as `achieved` averages them. Algorithmic mean: 2.199e9 -> no re-reads (part of z is still in L2).
SLAB_TRAFFIC_B64 = 25.93e9 / 12.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PDEQX_PRECISION", "tf32"), choices=["tf32", "fp32"])
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH)
    ap.add_argument("--cpu-sample", type=int, default=0, help="samples (= host threads) per CPU-baseline step; 0 = all cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap-allreduce", action="store_true", help="bucketed all-reduce on a side stream during the reverse pass")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying the captured CUDA graph")
    return ap.parse_args()


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


def cpu_train_step_rate(workers, steps, warmup):
    """One train step of the oracle port on `workers` samples, one sample per host thread (NumPy's pocketfft, einsum and
    BLAS calls release the GIL; BLAS itself is pinned to one thread per call), gradients averaged, one Adam update."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import nn as onn
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    rng = np.random.default_rng(0)
    p = onn.init_fno(rng, 2, 1, 1, hidden_channels=HIDDEN, num_modes=MODES, num_blocks=BLOCKS)
    state = onn.adam_init(p)
    x = rng.standard_normal((workers, 1, GRID, GRID)).astype(np.float32)
    y = rng.standard_normal((workers, 1, GRID, GRID)).astype(np.float32)

    def one(i):
        return onn.fno_loss_and_grad(p, x[i:i + 1], y[i:i + 1], (MODES, MODES), BLOCKS)

    times = []
    limiter = threadpool_limits(limits=1) if (threadpool_limits and workers > 1) else None
    try:
        with ThreadPoolExecutor(max_workers=workers) as pool:
            for it in range(warmup + steps):
                t0 = time.perf_counter()
                results = list(pool.map(one, range(workers)))
                g = {k: sum(r[1][k] for r in results).reshape(p[k].shape) / workers for k in results[0][1]}
                p, state = onn.adam_step(p, g, state)
                if it >= warmup:
                    times.append(time.perf_counter() - t0)
    finally:
        if limiter is not None:
            limiter.restore_original_limits() if hasattr(limiter, "restore_original_limits") else None
    dt = sum(times) / len(times)
    return workers / dt, dt, workers


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warmup = 1 if args.warmup > 0 else 0
    workers = args.cpu_sample if args.cpu_sample > 0 else cpu_workers()
    rate, dt, cores = cpu_train_step_rate(workers, steps, warmup)
    sample = (f"{workers} of the {args.global_batch} samples per step (one per host thread), {steps} timed steps of {dt:.1f} s; "
              "NumPy port of the reference train step (rfftn/einsum/irfftn forward, analytic reverse pass, Adam); "
              "jax/equinox are not installable in this image so the reference itself cannot run")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "timing": "host perf_counter, CPU only"},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
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
    from pdequinox_b200.conv._spectral_conv import get_plan

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
    step = TrainStep(model, lr=3e-4, overlap=args.overlap_allreduce)
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
    for _ in range(2):
        run_step()

    # ---- device-resident arm ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
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
    slab_ms, slab_n = _lib.C.c_double(), _lib.C.c_int64()
    _lib.check(L.pdeqx_timing_read(_lib.K_SPECTRAL_SLAB, _lib.C.byref(slab_ms), _lib.C.byref(slab_n)))
    L.pdeqx_timing_enable(0)

    if rank != 0:
        shutdown(world, graphed)
        return

    # ---- roofline of the dominant kernel (read x once + write y once [+ saved z]) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    n_pix = GRID * GRID
    # the slab kernel runs three times per block: forward (read x, z; write y), cotangent of the pre-activation (read x, z,
    # gy; write g) and data gradient (read g, z'; write gx) = 7 full-size tensor passes per block; z = [B,C,H,2m]; tables +
    # bypass weights are negligible. The folded ends of the network remove passes: the lifted first block never reads its
    # input (forward, cotangent) nor writes its input gradient; the last block's cotangent is rank one (gy is a single
    # channel) and, projected, its forward pass writes a single channel.
    act_code = pdeqx.nn.activation_code(model.blocks[0].activation)
    lifted = bool(model.blocks[0].lifted_ok(xd))
    rank1 = bool(L.pdeqx_spectral_block_rank1_ok(get_plan(2, per, HIDDEN, HIDDEN, (GRID, GRID), (MODES, MODES)).handle, 1, act_code))
    projected = bool(model.blocks[-1].projected_ok(xd))
    tensor_passes = 7 * BLOCKS - (3 if lifted else 0) - (1 if rank1 else 0) - (1 if projected else 0)
    slab_launches = 3 * BLOCKS
    small = 4 * (HIDDEN * HIDDEN + HIDDEN + 2 * MODES * GRID)
    one_tensor = 4 * per * HIDDEN * n_pix
    z_bytes = 4 * per * HIDDEN * GRID * 2 * MODES
    field = 4 * per * n_pix  # single-channel fields read by the folded variants
    n_fields = (3 if lifted else 0) + (1 if rank1 else 0) + (1 if projected else 0)
    slab_bytes = (tensor_passes * one_tensor + n_fields * field) / slab_launches + z_bytes + small  # mean over a step's launches
    roofline = None
    if slab_n.value > 0:
        avg_s = slab_ms.value * 1e-3 / slab_n.value
        achieved = slab_bytes / avg_s / 1e9
        roofline = {"bound": "hbm", "kernel": "tc::slab_kernel, spectral block last stage (c2r DFT + 1x1 bypass + bias + gelu / gelu' / adjoint), "
                              "mean over its %d launches per step (%d full-size tensor passes: lifting %s, projection %s)"
                              % (slab_launches, tensor_passes, "folded" if lifted else "separate", "folded" if projected else "separate"),
                    "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                    "algorithmic_bytes_per_launch": slab_bytes, "avg_launch_us": avg_s * 1e6,
                    "launches_timed": slab_n.value, "share_of_step": slab_ms.value / ms_probe,
                    "timed_in": f"{probe_steps} eagerly launched steps after the graph-replayed region (events cannot bracket "
                                "kernels inside a graph replay); share_of_step is relative to that pass",
                    "traffic": SLAB_TRAFFIC_B64 * per / 64.0,
                    "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the 12 launches of one step at 64 samples "
                                      "per GPU (profiles/r1_slab_traffic.csv), scaled by this run's samples per GPU"}

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
        workers = args.cpu_sample if args.cpu_sample > 0 else cpu_workers()
        rate, dt, cores = cpu_train_step_rate(workers, 5, 1)  # ~12 s of CPU work on a 16-core host
        cpu = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{workers} of the {args.global_batch} samples per step (one per host thread), 5 timed steps "
                         f"({dt:.1f} s each); NumPy port of the reference train step (jax/equinox not installable here)"}

    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "tf32 (fp32 accumulate, fp32 storage)" if args.precision == "tf32" else "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": args.global_batch, "per_gpu_batch": per,
                   "parallelism": f"dp{world} (batch-sharded, one NCCL all-reduce of the 67.2 MB gradient arena per step)",
                   "launch": "one CUDA graph replay per step" if graphed is not None else "eager kernel launches",
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
    emit(line)
    shutdown(world, graphed)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
