"""Idle time between the kernels of one graph-replayed configs[3] train step (CUPTI timeline through torch.profiler):
sum of kernel durations, span of the replay, and the gaps by predecessor kernel.
Usage: python scripts/graph_gaps.py [B] [fno2d|unet1d|resnet2d|dilated2d]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pdequinox_b200 as pdeqx
from pdequinox_b200.train import TrainStep, GraphedTrainStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pdeqx.set_precision("tf32")
which = sys.argv[2] if len(sys.argv) > 2 else "fno2d"
model, shape = {
    "fno2d": (lambda: pdeqx.ClassicFNO(2, 1, 1, hidden_channels=64, num_modes=16, num_blocks=4, key=0), (B, 1, 256, 256)),
    "unet1d": (lambda: pdeqx.ClassicUNet(1, 1, 1, boundary_mode="dirichlet", key=0), (B, 1, 64)),
    "resnet2d": (lambda: pdeqx.ClassicResNet(2, 1, 1, hidden_channels=64, key=0), (B, 1, 128, 128)),
    "dilated2d": (lambda: pdeqx.DilatedResNet(2, 1, 1, hidden_channels=64, key=0), (B, 1, 128, 128)),
}[which]
model = model()
step = TrainStep(model, lr=3e-4)
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).cuda()
y = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).cuda()
g = GraphedTrainStep(step, x, y, warmup=2)
for _ in range(5):
    g()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        g()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
print("device events:", len(ev))
if not ev:
    raise SystemExit("no device events recorded")
n = len(ev) // 3
last = ev[2 * n:]
busy = sum(e.time_range.end - e.time_range.start for e in last)
span = last[-1].time_range.end - last[0].time_range.start
print(f"last replay: {len(last)} device activities, busy {busy:.1f} us, span {span:.1f} us, idle {span - busy:.1f} us")
gaps = collections.defaultdict(lambda: [0, 0.0])
for a, b in zip(last[:-1], last[1:]):
    d = b.time_range.start - a.time_range.end
    k = a.name[:60]
    gaps[k][0] += 1
    gaps[k][1] += d
for k, (c, t) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t:8.1f} us idle after {c:3d} x {k}")
tot = collections.defaultdict(lambda: [0, 0.0])
for e in last:
    k = e.name.replace("void ", "").replace("pdeqx::", "").split("(")[0][:70]
    tot[k][0] += 1
    tot[k][1] += e.time_range.end - e.time_range.start
print(f"\nin-situ kernel time of one replayed {which} step (B = {B} per GPU):")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us {c:4d} launches {t / c:8.1f} us/launch {100 * t / busy:5.1f}%  {k}")
