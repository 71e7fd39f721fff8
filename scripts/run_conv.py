"""PhysicsConv 3x3 64->64 at BASELINE configs[2] size (B=128, 128x128): device time of fwd / data grad / filter grad."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pdequinox_b200 as pdeqx
from pdequinox_b200.train import ParameterArena

B = int(os.environ.get("B", "128"))
prec = os.environ.get("PREC", "tf32")
d = int(os.environ.get("DIL", "1"))
mode = os.environ.get("MODE", "periodic")
pdeqx.set_precision(prec)
conv = pdeqx.PhysicsConv(2, 64, 64, 3, dilation=d, key=0, boundary_mode=mode)
arena = ParameterArena(conv)
x = torch.randn(B, 64, 128, 128, device="cuda")
gy = torch.randn(B, 64, 128, 128, device="cuda")
from pdequinox_b200 import _lib
L = _lib.lib()
desc, _ = conv._desc(x)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
y = torch.empty_like(x); gx = torch.empty_like(x)
nws = L.pdeqx_conv_workspace_bytes(desc)
ws = torch.empty(nws // 4 + 1, device="cuda")
gb = 4 * B * 64 * 128 * 128 * 2 / 1e9
tf = t(lambda: _lib.check(L.pdeqx_conv_fwd(desc, _lib.ptr(x), _lib.ptr(conv.weight), _lib.ptr(conv.bias), None, 1, _lib.ptr(y), _lib.ptr(ws), nws, _lib.stream())))
td = t(lambda: _lib.check(L.pdeqx_conv_bwd_data(desc, _lib.ptr(gy), _lib.ptr(conv.weight), None, None, _lib.ptr(gx), _lib.ptr(ws), nws, _lib.stream())))
tw = t(lambda: _lib.check(L.pdeqx_conv_bwd_filter(desc, _lib.ptr(x), _lib.ptr(gy), _lib.ptr(conv.grad("weight")), _lib.ptr(conv.grad("bias")), _lib.ptr(ws), nws, _lib.stream())), n=2)
print(f"conv 3x3 64->64 B={B} d={d} {mode} {prec}: fwd {tf:.1f} us ({gb/tf*1e6:.0f} GB/s), bwd_data {td:.1f} us ({gb/td*1e6:.0f} GB/s), bwd_filter {tw:.1f} us ({gb/tw*1e6:.0f} GB/s)")
