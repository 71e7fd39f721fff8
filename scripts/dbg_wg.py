import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import pdequinox_b200 as pdeqx
from pdequinox_b200 import _lib
from pdequinox_b200.train import ParameterArena
from oracle import conv as oconv
pdeqx.set_precision("tf32")
mode = os.environ.get("MODE", "periodic"); d = int(os.environ.get("DIL", "1"))
conv = pdeqx.PhysicsConv(2, 64, 64, 3, dilation=d, key=0, boundary_mode=mode)
arena = ParameterArena(conv)
rng = np.random.default_rng(0)
x = rng.standard_normal((2, 64, 16, 128)); gy = rng.standard_normal((2, 64, 16, 128))
xd = torch.from_numpy(x).float().cuda(); gyd = torch.from_numpy(gy).float().cuda()
L = _lib.lib(); desc, _ = conv._desc(xd)
print("passes", L.pdeqx_conv_tcgen05_passes(desc))
y = torch.empty_like(xd)
_lib.check(L.pdeqx_conv_fwd(desc, _lib.ptr(xd), _lib.ptr(conv.weight), _lib.ptr(conv.bias), None, 0, _lib.ptr(y), _lib.stream())); torch.cuda.synchronize(); print("fwd ok")
_lib.check(L.pdeqx_conv_bwd_filter(desc, _lib.ptr(xd), _lib.ptr(gyd), _lib.ptr(conv.grad("weight")), _lib.ptr(conv.grad("bias")), _lib.stream())); torch.cuda.synchronize(); print("bwd_filter ok")
w = conv.weight.cpu().numpy().astype(np.float64)
_, gw_r, gb_r = oconv.physics_conv_vjp(x, w, gy, boundary_mode=mode, dilation=d)
gw = conv.grad("weight").cpu().numpy(); gb = conv.grad("bias").cpu().numpy()
rel = lambda a, b: np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())
print("gw rel", rel(gw, gw_r), "gb rel", rel(gb, gb_r))
for tw in range(3):
    for th in range(3):
        print(th, tw, rel(gw[:, :, th, tw], gw_r[:, :, th, tw]))
