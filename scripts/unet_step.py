"""configs[1] (ClassicUNet 1D dirichlet, batch 32): a few eager train steps for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pdequinox_b200 as pdeqx
from pdequinox_b200.train import TrainStep

pdeqx.set_precision(os.environ.get("PREC", "tf32"))
model = pdeqx.ClassicUNet(1, 1, 1, boundary_mode="dirichlet", key=0)
step = TrainStep(model)
x, y = torch.randn(32, 1, 64, device="cuda"), torch.randn(32, 1, 64, device="cuda")
for _ in range(int(os.environ.get("N", "3"))):
    step(x, y)
torch.cuda.synchronize()
print("loss", float(step.loss.item()))
