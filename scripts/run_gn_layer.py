"""One norm -> conv -> relu layer of DilatedResNet (BASELINE.json configs[2]: 64 channels, 128x128, batch 128, periodic) with the
GroupNorm folded into the convolution's passes: device time of every pass (for ncu: -k regex:'conv_rows_pair|conv_wgrad|gn_')."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pdequinox_b200 as pdeqx
from pdequinox_b200 import _lib, nn as pnn
from pdequinox_b200._module import Module, Tape
from pdequinox_b200.train import ParameterArena

B, d = int(os.environ.get("B", "128")), int(os.environ.get("DIL", "1"))
pdeqx.set_precision("tf32")


class Layer(Module):
    _fields = ("norm", "conv")

    def __init__(self):
        self.num_spatial_dims = 2
        self.norm = pnn.GroupNorm(1, 64)
        self.conv = pdeqx.PhysicsConv(2, 64, 64, 3, dilation=d, key=0, boundary_mode="periodic")


m = Layer()
ParameterArena(m)
x = torch.randn(B, 64, 128, 128, device="cuda").relu_()
ga = torch.randn(B, 64, 128, 128, device="cuda")
tape = Tape()
assert m.conv.gn_fusable(x)


def t(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


stats = m.norm.stats_of(x, tape)
out = {}
out["stats pass (first layer of a block only)"] = t(lambda: m.norm.stats_of(x, tape))
out["forward (conv + norm folded + relu + sums of y)"] = t(lambda: m.conv.apply_gn(x, m.norm, stats, tape, act=_lib.ACT_RELU))
res = {}


def bwd_conv():
    res["v"] = m.conv.apply_bwd_gn(x, ga, m.norm, stats, tape)


out["filter gradient + data gradient (norm folded)"] = t(bwd_conv)
v, parts, slots = res["v"]
out["norm reverse pass (elementwise, relu' + parameter sums)"] = t(lambda: m.norm.bwd_fused(x, v, stats, parts, slots, tape, mask_h=True))
gb = 4 * B * 64 * 128 * 128 / 1e9
for k, us in out.items():
    print(f"{k}: {us:.1f} us")
print(f"(one activation-sized tensor = {gb:.3f} GB = {gb / 6.5367 * 1e3:.0f} us at the measured HBM peak)")
