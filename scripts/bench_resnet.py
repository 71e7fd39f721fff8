"""Secondary workload: BASELINE.json configs[2] -- ClassicResNet / DilatedResNet 2D periodic, 128x128, hidden 64,
batch 128: Adam train-step time (device-resident inputs, CUDA events)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pdequinox_b200 as pdeqx
from pdequinox_b200 import _lib
from pdequinox_b200.train import TrainStep

prec = os.environ.get("PREC", "tf32")
B = int(os.environ.get("B", "128"))
pdeqx.set_precision(prec)
out = {}
want = os.environ.get("MODELS", "ClassicResNet,DilatedResNet").split(",")
n_warm, n = int(os.environ.get("WARM", "3")), int(os.environ.get("N", "5"))
for name, ctor in [("ClassicResNet", lambda: pdeqx.ClassicResNet(2, 1, 1, hidden_channels=64, boundary_mode="periodic", key=0)),
                   ("DilatedResNet", lambda: pdeqx.DilatedResNet(2, 1, 1, hidden_channels=64, boundary_mode="periodic", key=0))]:
    if name not in want:
        continue
    model = ctor()
    step = TrainStep(model)
    x = torch.randn(B, 1, 128, 128, device="cuda"); y = torch.randn(B, 1, 128, 128, device="cuda")
    for _ in range(n_warm): step(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.lib().pdeqx_launch_count()
    e0.record()
    for _ in range(n): step(x, y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out[name] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3, "params": pdeqx.count_parameters(model),
                 "launches_per_step": (_lib.lib().pdeqx_launch_count() - l0) // n, "precision": prec, "batch": B,
                 "loss": float(step.loss.item())}
    del step, model
    torch.cuda.empty_cache()
print(json.dumps(out))
