"""GPU diagnostics for the tcgen05 spectral kernels (not a test): isolates each stage."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pdequinox_b200 as pdeqx
from pdequinox_b200 import _lib
from oracle import spectral as ospec, conv as oconv


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def run_block(x, wr, wi, bw, bb, modes, act="identity"):
    D = len(modes)
    ci, co = x.shape[1], wr.shape[1]
    blk = pdeqx.ClassicSpectralBlock(D, ci, co, activation=getattr(pdeqx.nn, act), num_modes=modes, key=0)
    blk.spectral_conv.weights_real, blk.spectral_conv.weights_imag = dev(wr), dev(wi)
    blk.by_pass_conv.weight, blk.by_pass_conv.bias = dev(bw), dev(bb)
    y = blk._fwd(dev(x), None)
    torch.cuda.synchronize()
    return y.cpu().numpy()


pdeqx.set_precision("tf32")
rng = np.random.default_rng(0)
B, C, H, W = 2, 32, 16, 128
modes = (4, 16)
x = rng.standard_normal((B, C, H, W))
wr = rng.standard_normal((2, C, C) + modes) / 32
wi = rng.standard_normal((2, C, C) + modes) / 32
zero = np.zeros_like(wr)

# 1. plain spectral conv (r2c + slab K-major part)
conv = pdeqx.SpectralConv(2, C, C, modes, key=0)
conv.weights_real, conv.weights_imag = dev(wr), dev(wi)
y = pdeqx.vmap(conv)(dev(x)).cpu().numpy()
ref = ospec.spectral_conv_nd(x, wr, wi, modes)
print("plain spectral conv (tf32 path): rel", rel(y, ref))
_lib.lib().pdeqx_set_fast_paths(0)
y0 = pdeqx.vmap(conv)(dev(x)).cpu().numpy()
_lib.lib().pdeqx_set_fast_paths(1)
print("plain spectral conv (generic)  : rel", rel(y0, ref))

# 2. bypass only: W = identity, zero spectral weights
eye = np.eye(C)[:, :, None, None]
y = run_block(x, zero, zero, eye, np.zeros((C, 1, 1)), modes)
print("bypass identity: rel", rel(y, x))
if rel(y, x) > 1e-2:
    print(" y[0,0,0,:8] =", y[0, 0, 0, :8], "\n x[0,0,0,:8] =", x[0, 0, 0, :8])
    # decode with one-hot inputs
    for (c, w) in [(0, 0), (1, 0), (0, 1), (0, 33), (8, 5), (3, 100)]:
        xo = np.zeros((1, C, 1, W))
        xo[0, c, 0, w] = 1.0
        # need H multiple? use H=1 rows
        yo = run_block(xo, zero[:, :, :, :1], zero[:, :, :, :1], eye, np.zeros((C, 1, 1)), (1, 16))
        nz = np.argwhere(np.abs(yo) > 1e-3)
        print(f" one-hot (c={c}, w={w}) ->", [(int(a[1]), int(a[3]), float(yo[tuple(a)])) for a in nz[:6]])

# 3. random W
bw = rng.standard_normal((C, C, 1, 1)) / np.sqrt(C)
bb = rng.standard_normal((C, 1, 1))
y = run_block(x, zero, zero, bw, bb, modes)
print("bypass random W + bias: rel", rel(y, oconv.pointwise_conv(x, bw, bb)))
y = run_block(x, wr, wi, bw, bb, modes, act="gelu")
from oracle import nn as onn
print("full block gelu: rel", rel(y, onn.gelu_tanh(ospec.spectral_conv_nd(x, wr, wi, modes) + oconv.pointwise_conv(x, bw, bb))))
