"""jax.lax stand-in: conv_general_dilated through torch's CPU convolution (an implementation that is
independent of the oracle's tap-loop restatement)."""
import numpy as _np


def stop_gradient(x):
    return x


def conv_general_dilated(lhs, rhs, window_strides, padding, lhs_dilation=None, rhs_dilation=None,
                         dimension_numbers=None, feature_group_count=1, **kw):
    import torch
    import torch.nn.functional as F
    x = torch.from_numpy(_np.ascontiguousarray(_np.asarray(lhs))).double()
    w = torch.from_numpy(_np.ascontiguousarray(_np.asarray(rhs))).double()
    D = w.ndim - 2
    stride = tuple(int(s) for s in window_strides)
    rd = tuple(int(d) for d in (rhs_dilation or (1,) * D))
    ld = tuple(int(d) for d in (lhs_dilation or (1,) * D))
    if any(f != 1 for f in ld):  # zero insertion between input samples
        shape = list(x.shape)
        for a, f in enumerate(ld):
            shape[2 + a] = (x.shape[2 + a] - 1) * f + 1
        xd = torch.zeros(shape, dtype=x.dtype)
        xd[(slice(None), slice(None)) + tuple(slice(None, None, f) for f in ld)] = x
        x = xd
    pads = [(int(lo), int(hi)) for lo, hi in padding]
    # negative padding = crop; positive = zero pad (F.pad takes the LAST axis first)
    crop = [slice(None), slice(None)]
    for a, (lo, hi) in enumerate(pads):
        n = x.shape[2 + a]
        crop.append(slice(-lo if lo < 0 else 0, n + hi if hi < 0 else n))
    x = x[tuple(crop)]
    flat = []
    for lo, hi in reversed(pads):
        flat += [max(lo, 0), max(hi, 0)]
    x = F.pad(x, flat)
    conv = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[D]
    y = conv(x, w, None, stride, 0, rd, int(feature_group_count))
    return y.numpy().astype(_np.asarray(lhs).dtype)
