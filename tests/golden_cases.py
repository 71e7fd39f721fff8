"""Loader + oracle evaluator for tests/golden/hotpath_golden.npz (fixtures produced by running the reference's
own source, see tests/golden/make_golden.py)."""
import json
import os

import numpy as np

from oracle import conv as oconv, nn as onn, spectral as ospec

HERE = os.path.dirname(os.path.abspath(__file__))
_META = json.load(open(os.path.join(HERE, "golden", "hotpath_golden.json")))
_NPZ = None
CASES = _META["cases"]
PARAMETER_COUNTS = _META["parameter_counts"]


def arrays(name):
    global _NPZ
    if _NPZ is None:
        _NPZ = np.load(os.path.join(HERE, "golden", "hotpath_golden.npz"))
    pre = name + "/"
    x, y = _NPZ[pre + "x"], _NPZ[pre + "y"]
    p = {k[len(pre) + 2:]: _NPZ[k] for k in _NPZ.files if k.startswith(pre + "p/")}
    return x, y, p


def oracle_eval(case, x, p):
    """The oracle's answer for one golden case (x is ONE sample, as the reference is called)."""
    kind, kw = case["kind"], case["kwargs"]
    xb = x[None]
    if kind in ("SpectralConv", "SpectralConvRaw"):
        y = ospec.spectral_conv_nd(xb, p["weights_real"], p["weights_imag"], tuple(kw["num_modes"]))
    elif kind == "PhysicsConv":
        y = oconv.physics_conv(xb, p["weight"], p.get("bias"), boundary_mode=kw["boundary_mode"], stride=kw["stride"],
                               dilation=kw["dilation"], groups=kw["groups"])
    elif kind == "PhysicsConvTranspose":
        y = oconv.physics_conv_transpose(xb, p["weight"], p.get("bias"), boundary_mode=kw["boundary_mode"],
                                         stride=kw["stride"], output_padding=kw["output_padding"],
                                         dilation=kw.get("dilation", 1))
    elif kind == "PointwiseLinearConv":
        y = oconv.pointwise_conv(xb, p["weight"], p.get("bias"))
    elif kind == "ClassicSpectralBlock":
        q = {"b." + k: v for k, v in p.items()}
        modes = (kw["num_modes"],) * kw["num_spatial_dims"]
        y = onn.spectral_block(xb, q, "b.", modes, case["activation"])
    elif kind == "ClassicResBlock":
        y = onn.res_block_any(xb, p, "", kw["boundary_mode"], case["activation"], kw["num_spatial_dims"])
    elif kind == "DilatedResBlock":
        y = onn.dilated_block_any(xb, p, "", kw["boundary_mode"], case["activation"], kw["num_spatial_dims"])
    elif kind == "ClassicFNO":
        D = kw["num_spatial_dims"]
        nb = kw.get("num_blocks", 4)
        y = onn.fno_forward(p, xb, (kw.get("num_modes", 12),) * D, nb)
    elif kind == "ClassicResNet":
        D = kw["num_spatial_dims"]
        y = onn.sequential_forward(p, xb, lambda h, pre: onn.res_block_any(h, p, pre, kw["boundary_mode"], "relu", D),
                                   kw.get("num_blocks", 6))
    elif kind == "DilatedResNet":
        D = kw["num_spatial_dims"]
        y = onn.sequential_forward(p, xb, lambda h, pre: onn.dilated_block_any(h, p, pre, kw["boundary_mode"], "relu", D),
                                   kw.get("num_blocks", 2))
    elif kind == "ModernResBlock":
        y = onn.modern_block_any(xb, p, "", kw["boundary_mode"], case["activation"], kw["num_spatial_dims"])
    elif kind == "ModernResNet":
        D = kw["num_spatial_dims"]
        y = onn.sequential_forward(p, xb, lambda h, pre: onn.modern_block_any(h, p, pre, kw["boundary_mode"], "relu", D),
                                   kw.get("num_blocks", 6))
    elif kind == "ModernUNet":
        y = onn.modern_unet_forward(p, xb, kw["boundary_mode"], kw.get("num_levels", 4), kw.get("num_blocks", 2))
    elif kind == "LinearConvBlock":
        y = oconv.physics_conv(xb, p["weight"], p.get("bias"), boundary_mode=kw["boundary_mode"])
    elif kind == "ConvNet":
        y = onn.conv_net_forward(p, xb, kw["boundary_mode"])
    elif kind == "ClassicUNet":
        y = onn.unet_forward(p, xb, kw["boundary_mode"], kw.get("num_levels", 4))
    elif kind == "ClassicDoubleConvBlock":
        y = onn.double_conv_block(xb, p, "", kw["boundary_mode"], case["activation"], kw["num_spatial_dims"])
    elif kind == "LinearConvDownBlock":
        y = oconv.physics_conv(xb, p["weight"], p.get("bias"), boundary_mode=kw["boundary_mode"], stride=2)
    elif kind == "LinearConvUpBlock":
        y = oconv.physics_conv_transpose(xb, p["weight"], p.get("bias"), boundary_mode=kw["boundary_mode"], stride=2,
                                         output_padding=1)
    else:
        raise KeyError(kind)
    return y[0]
