"""One ClassicSpectralBlock forward + reverse pass at the BASELINE configs[3] size (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pdequinox_b200 as pdeqx
from pdequinox_b200.train import ParameterArena

B = int(os.environ.get("B", "64"))
prec = os.environ.get("PREC", "tf32")
iters = int(os.environ.get("ITERS", "3"))
pdeqx.set_precision(prec)
blk = pdeqx.ClassicSpectralBlock(2, 64, 64, activation=pdeqx.nn.gelu, num_modes=16, key=0)
arena = ParameterArena(blk)
x = torch.randn(B, 64, 256, 256, device="cuda")
gy = torch.randn(B, 64, 256, 256, device="cuda")
tape = pdeqx._module.Tape()
for _ in range(iters):
    tape.stack.clear()
    y = blk._fwd(x, tape)
    gx = blk._bwd(gy, tape, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tape.stack.clear()
y = blk._fwd(x, tape)
e1.record()
torch.cuda.synchronize()
t_f = e0.elapsed_time(e1)
e0.record()
gx = blk._bwd(gy, tape, True)
e1.record()
torch.cuda.synchronize()
print(f"block fwd {t_f:.3f} ms, bwd {e0.elapsed_time(e1):.3f} ms (B={B}, {prec})")
# per-kernel device timing of one more forward + reverse pass
from pdequinox_b200 import _lib
L = _lib.lib()
L.pdeqx_timing_enable(1)
tape.stack.clear()
y = blk._fwd(x, tape)
gx = blk._bwd(gy, tape, True)
torch.cuda.synchronize()
names = {1: "slab", 2: "r2c", 3: "modes(c2c)", 4: "wgrad"}
out = []
for tag, nm in names.items():
    ms, n = _lib.C.c_double(), _lib.C.c_int64()
    L.pdeqx_timing_read(tag, _lib.C.byref(ms), _lib.C.byref(n))
    out.append(f"{nm} {ms.value * 1e3:.0f} us/{n.value}")
L.pdeqx_timing_enable(0)
print("   " + ", ".join(out))
