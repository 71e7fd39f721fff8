"""Train-step throughput of every BASELINE.json config on ONE B200 (synthetic data of the named shape, random-init weights):
samples/s from CUDA events over graph-replayed steps (eager launches if capture fails), tf32 mode unless PREC=fp32.

    python scripts/bench_configs.py [names...]      names: fno1d unet1d resnet2d dilated2d fno2d fno3d
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pdequinox_b200 as pdeqx
from pdequinox_b200.train import TrainStep, GraphedTrainStep

prec = os.environ.get("PREC", "tf32")
pdeqx.set_precision(prec)
CONFIGS = {
    # name: (constructor, input shape (B, C, *S), BASELINE.json config it stands for)
    "fno1d": (lambda: pdeqx.ClassicFNO(1, 1, 1, key=0), (32, 1, 64), "configs[0] ClassicFNO 1D defaults, batch 32"),
    "unet1d": (lambda: pdeqx.ClassicUNet(1, 1, 1, boundary_mode="dirichlet", key=0), (32, 1, 64), "configs[1] ClassicUNet 1D dirichlet, batch 32"),
    "resnet2d": (lambda: pdeqx.ClassicResNet(2, 1, 1, hidden_channels=64, key=0), (128, 1, 128, 128), "configs[2] ClassicResNet 2D periodic 128^2 hidden 64, batch 128"),
    "dilated2d": (lambda: pdeqx.DilatedResNet(2, 1, 1, hidden_channels=64, key=0), (128, 1, 128, 128), "configs[2] DilatedResNet 2D periodic 128^2 hidden 64, batch 128"),
    "fno2d": (lambda: pdeqx.ClassicFNO(2, 1, 1, hidden_channels=64, num_modes=16, num_blocks=4, key=0), (64, 1, 256, 256), "configs[3] ClassicFNO 2D 256^2 hidden 64 modes 16, batch 64"),
    "fno3d": (lambda: pdeqx.ClassicFNO(3, 1, 1, hidden_channels=32, num_modes=12, num_blocks=4, key=0), (int(os.environ.get("B3D", "32")), 1, 64, 64, 64), "configs[4] ClassicFNO 3D 64^3 hidden 32 modes 12, batch 32 (one GPU)"),
}
names = sys.argv[1:] or list(CONFIGS)
steps = int(os.environ.get("STEPS", "10"))
for name in names:
    ctor, shape, label = CONFIGS[name]
    model = ctor()
    step = TrainStep(model)
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal(shape, dtype=np.float32)).cuda()
    y = torch.from_numpy(rng.standard_normal(shape, dtype=np.float32)).cuda()
    for _ in range(3):
        step(x, y)
    torch.cuda.synchronize()
    mode = "graph"
    try:
        run = GraphedTrainStep(step, x, y, warmup=1)
    except Exception as exc:
        mode = f"eager ({type(exc).__name__})"
        run = lambda: step(x, y)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"config": label, "params": pdeqx.count_parameters(model), "precision": prec, "launch": mode,
                      "ms_per_step": round(ms, 4), "samples_per_s": round(shape[0] / ms * 1e3, 1),
                      "loss_finite": bool(torch.isfinite(step.loss).all())}), flush=True)
    del run, step, model, x, y
    torch.cuda.empty_cache()
