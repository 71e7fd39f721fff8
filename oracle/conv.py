"""Oracle: boundary-aware spatial convolution (TEST INFRASTRUCTURE ONLY).

Follows
  * /root/reference/pdequinox/conv/_physics_conv.py:11-26  (SAME padding)
  * /root/reference/pdequinox/conv/_physics_conv.py:92-101 (boundary -> padding mode)
  * /root/reference/pdequinox/conv/_conv.py:169-219        (pad, cross-correlate, bias)
  * /root/reference/pdequinox/conv/_conv.py:378-471        (transposed conv, reference-specific)
  * /root/reference/pdequinox/conv/_pointwise_linear_conv.py:37-49 (1x1 conv)
`lax.conv_general_dilated` itself is third-party (jax); it is restated here from
its documented definition (cross-correlation, NC*/OI* layouts, window_strides,
lhs_dilation, rhs_dilation, feature_group_count).

Arrays carry arbitrary leading batch axes: x is (..., C, *S).
"""
from __future__ import annotations

import math

import numpy as np

BOUNDARY_TO_PADDING = {"periodic": "circular", "dirichlet": "zeros", "neumann": "reflect"}
_NP_PAD_MODE = {"circular": "wrap", "reflect": "reflect", "replicate": "edge"}


def _ntuple(v, n):
    if isinstance(v, (tuple, list)):
        if len(v) != n:
            raise ValueError(f"Length of {v} (length = {len(v)}) is not equal to {n}")
        return tuple(int(i) for i in v)
    return (int(v),) * n


def compute_same_padding(num_spatial_dims, kernel_size, dilation):
    """_physics_conv.py:11-26."""
    k = _ntuple(kernel_size, num_spatial_dims)
    d = _ntuple(dilation, num_spatial_dims)
    return tuple((((ki - 1) // 2) * di, (((ki - 1) // 2) + ((ki - 1) % 2)) * di) for ki, di in zip(k, d))


def pad_boundary(x, padding, padding_mode, D):
    """jnp.pad(x, ((0,0),)+padding, mode=wrap|reflect|edge) (_conv.py:189-200); zeros pads with 0."""
    lead = x.ndim - D
    pw = ((0, 0),) * lead + tuple(tuple(p) for p in padding)
    if padding_mode == "zeros":
        return np.pad(x, pw, mode="constant")
    return np.pad(x, pw, mode=_NP_PAD_MODE[padding_mode])


def _dilate_input(x, lhs_dilation, D):
    """Insert (f-1) zeros between neighbouring input samples on each spatial axis."""
    if all(f == 1 for f in lhs_dilation):
        return x
    lead = x.ndim - D
    shp = list(x.shape)
    for a, f in enumerate(lhs_dilation):
        shp[lead + a] = (x.shape[lead + a] - 1) * f + 1
    out = np.zeros(shp, dtype=x.dtype)
    idx = (slice(None),) * lead + tuple(slice(None, None, f) for f in lhs_dilation)
    out[idx] = x
    return out


def conv_general_dilated(x, w, window_strides, padding, lhs_dilation=None, rhs_dilation=None, groups=1):
    """Documented semantics of jax.lax.conv_general_dilated for NC*/OI* layouts.

    x (..., C_i, *S); w (C_o, C_i/groups, *k). Cross-correlation (no kernel flip).
    padding: per-axis (lo, hi), may be negative (= crop).
    """
    D = w.ndim - 2
    lead = x.ndim - D - 1
    stride = _ntuple(window_strides, D)
    ld = _ntuple(lhs_dilation or 1, D)
    rd = _ntuple(rhs_dilation or 1, D)
    x = _dilate_input(x, ld, D)
    # zero pad / crop
    pos = tuple((max(lo, 0), max(hi, 0)) for lo, hi in padding)
    x = np.pad(x, ((0, 0),) * (lead + 1) + pos, mode="constant")
    crop = [slice(None)] * (lead + 1)
    for a, (lo, hi) in enumerate(padding):
        n = x.shape[lead + 1 + a]
        crop.append(slice(-lo if lo < 0 else 0, n + hi if hi < 0 else n))
    x = x[tuple(crop)]
    co, cig = w.shape[:2]
    k = w.shape[2:]
    ci = x.shape[lead]
    if ci != cig * groups or co % groups:
        raise ValueError("channel/group mismatch")
    spatial = x.shape[lead + 1:]
    out_sp = tuple((spatial[a] - rd[a] * (k[a] - 1) - 1) // stride[a] + 1 for a in range(D))
    acc_dtype = np.float64 if x.dtype == np.float64 else np.float32
    y = np.zeros(x.shape[:lead] + (co,) + out_sp, dtype=acc_dtype)
    cog = co // groups
    for tap in np.ndindex(*k):
        sl = tuple(slice(tap[a] * rd[a], tap[a] * rd[a] + (out_sp[a] - 1) * stride[a] + 1, stride[a]) for a in range(D))
        xs = x[(Ellipsis,) + sl]  # (..., C_i, *out)
        wt = w[(slice(None), slice(None)) + tap]  # (C_o, C_i/g)
        for g in range(groups):
            xg = xs[(slice(None),) * lead + (slice(g * cig, (g + 1) * cig),)]
            contrib = np.tensordot(wt[g * cog:(g + 1) * cog], xg, axes=([1], [lead]))  # (cog, ..., *out)
            contrib = np.moveaxis(contrib, 0, lead)
            y[(slice(None),) * lead + (slice(g * cog, (g + 1) * cog),)] += contrib
    return y.astype(x.dtype)


def more_padding_conv(x, weight, bias, *, stride, padding, padding_mode, dilation, groups=1):
    """MorePaddingConv.__call__ (_conv.py:169-219)."""
    D = weight.ndim - 2
    if padding_mode not in ("zeros", "reflect", "replicate", "circular"):
        raise ValueError("`padding_mode` must be one of ['zeros', 'reflect', 'replicate', 'circular']")
    if padding_mode == "zeros":
        padding_lax = tuple(tuple(p) for p in padding)
    else:
        x = pad_boundary(x, padding, padding_mode, D)
        padding_lax = ((0, 0),) * D
    y = conv_general_dilated(x, weight, _ntuple(stride, D), padding_lax, None, _ntuple(dilation, D), groups)
    if bias is not None:
        y = y + bias.reshape((-1,) + (1,) * D).astype(y.dtype)
    return y


def physics_conv(x, weight, bias, *, boundary_mode, stride=1, dilation=1, groups=1):
    """PhysicsConv (_physics_conv.py:29-118) = MorePaddingConv with SAME padding."""
    D = weight.ndim - 2
    mode = boundary_mode.lower()
    if mode not in BOUNDARY_TO_PADDING:
        raise ValueError(f"Only 'periodic', 'dirichlet', 'neumann' boundary modes are supported, got {boundary_mode}")
    k = weight.shape[2:]
    pad = compute_same_padding(D, k, dilation)
    return more_padding_conv(x, weight, bias, stride=stride, padding=pad,
                             padding_mode=BOUNDARY_TO_PADDING[mode], dilation=dilation, groups=groups)


def pointwise_conv(x, weight, bias):
    """PointwiseLinearConv = eqx.nn.Conv(kernel_size=1, padding=0) (_pointwise_linear_conv.py:37-49)."""
    D = weight.ndim - 2
    lead = x.ndim - D - 1
    w2 = weight.reshape(weight.shape[0], weight.shape[1])
    y = np.moveaxis(np.tensordot(w2, x, axes=([1], [lead])), 0, lead)
    if bias is not None:
        y = y + bias.reshape((-1,) + (1,) * D)
    return y.astype(x.dtype)


def pointwise_conv_vjp(x, weight, gy, use_bias=True):
    D = weight.ndim - 2
    lead = x.ndim - D - 1
    w2 = weight.reshape(weight.shape[0], weight.shape[1])
    gx = np.moveaxis(np.tensordot(w2.T, gy, axes=([1], [lead])), 0, lead)
    red_axes = tuple(a for a in range(x.ndim) if a != lead)
    gw = np.tensordot(np.moveaxis(gy, lead, 0).reshape(gy.shape[lead], -1),
                      np.moveaxis(x, lead, 0).reshape(x.shape[lead], -1), axes=([1], [1]))
    gb = gy.sum(axis=red_axes).reshape((-1,) + (1,) * D) if use_bias else None
    return gx.astype(x.dtype), gw.reshape(weight.shape).astype(x.dtype), gb


# ---------------------------------------------------------------------------
# VJP of PhysicsConv (stride/dilation/groups general); the reference relies on
# JAX autodiff -- checked against finite differences in float64.
# ---------------------------------------------------------------------------

def _pad_adjoint(gpad, padding, padding_mode, D):
    """Adjoint of pad_boundary: wrap -> add halo back modulo s; zeros -> drop; reflect -> fold."""
    lead = gpad.ndim - D
    g = gpad
    for a, (lo, hi) in enumerate(padding):
        ax = lead + a
        n = g.shape[ax] - lo - hi
        g = np.moveaxis(g, ax, -1)
        core = g[..., lo:lo + n].copy()
        if padding_mode == "zeros":
            pass
        elif padding_mode == "circular":
            for j in range(lo):       # padded index j  <-> source (j - lo) mod n
                core[..., (j - lo) % n] += g[..., j]
            for j in range(hi):       # padded index lo+n+j <-> source j mod n
                core[..., j % n] += g[..., lo + n + j]
        elif padding_mode == "reflect":
            for j in range(lo):       # padded index j <-> source lo - j
                core[..., lo - j] += g[..., j]
            for j in range(hi):       # padded index lo+n+j <-> source n-2-j
                core[..., n - 2 - j] += g[..., lo + n + j]
        elif padding_mode == "replicate":
            core[..., 0] += g[..., :lo].sum(-1)
            core[..., n - 1] += g[..., lo + n:].sum(-1)
        else:
            raise ValueError(padding_mode)
        g = np.moveaxis(core, -1, ax)
    return g


def physics_conv_vjp(x, weight, gy, *, boundary_mode, stride=1, dilation=1, groups=1, use_bias=True):
    """Returns (gx, gw, gb) for y = physics_conv(x, weight, bias)."""
    D = weight.ndim - 2
    lead = x.ndim - D - 1
    mode = BOUNDARY_TO_PADDING[boundary_mode.lower()]
    k = weight.shape[2:]
    st = _ntuple(stride, D)
    rd = _ntuple(dilation, D)
    pad = compute_same_padding(D, k, dilation)
    xpad = pad_boundary(x, pad, mode, D)
    co, cig = weight.shape[:2]
    cog = co // groups
    out_sp = gy.shape[lead + 1:]
    gxpad = np.zeros_like(xpad)
    gw = np.zeros_like(weight)
    for tap in np.ndindex(*k):
        sl = tuple(slice(tap[a] * rd[a], tap[a] * rd[a] + (out_sp[a] - 1) * st[a] + 1, st[a]) for a in range(D))
        for g in range(groups):
            chan_o = (slice(None),) * lead + (slice(g * cog, (g + 1) * cog),)
            chan_i = (slice(None),) * lead + (slice(g * cig, (g + 1) * cig),)
            gyg = gy[chan_o]
            xs = xpad[chan_i + sl]
            wt = weight[(slice(g * cog, (g + 1) * cog), slice(None)) + tap]  # (cog, cig)
            gxpad[chan_i + sl] += np.moveaxis(np.tensordot(wt.T, gyg, axes=([1], [lead])), 0, lead)
            gflat = np.moveaxis(gyg, lead, 0).reshape(cog, -1)
            xflat = np.moveaxis(xs, lead, 0).reshape(cig, -1)
            gw[(slice(g * cog, (g + 1) * cog), slice(None)) + tap] = gflat @ xflat.T
    gx = _pad_adjoint(gxpad, pad, mode, D)
    red_axes = tuple(a for a in range(gy.ndim) if a != lead)
    gb = gy.sum(axis=red_axes).reshape((-1,) + (1,) * D) if use_bias else None
    return gx, gw, gb


# ---------------------------------------------------------------------------
# Transposed convolution with the reference's specific semantics (App. A.3)
# ---------------------------------------------------------------------------

def more_padding_conv_transpose(x, weight, bias, *, stride, padding, output_padding, padding_mode, dilation, groups=1):
    """MorePaddingConvTranspose.__call__ (_conv.py:378-471): weight (C_out, C_in/groups, *k) is
    used as-is (no spatial flip, no in/out swap) in a stride-1 correlation over the
    stride-dilated input."""
    D = weight.ndim - 2
    k = weight.shape[2:]
    st = _ntuple(stride, D)
    rd = _ntuple(dilation, D)
    op = _ntuple(output_padding, D)
    padding = tuple(tuple(pp) for pp in padding)
    padding_t = tuple(  # _conv.py:411-416
        (d * (kk - 1) - p0, d * (kk - 1) - p1 + o)
        for (p0, p1), kk, o, d in zip(padding, k, op, rd)
    )
    if padding_mode != "zeros":
        pre_dilation_padding = tuple(  # _conv.py:420-426: ceil(transpose_padding / stride)
            ((p_l + (s - 1)) // s, (p_r + (s - 1)) // s) for (p_l, p_r), s in zip(padding_t, st)
        )
        post = tuple(  # _conv.py:428-436: from the module's own `padding`, NOT the transpose padding
            (p_l - pd_l * s, p_r - pd_r * s + o)
            for (p_l, p_r), (pd_l, pd_r), s, o in zip(padding, pre_dilation_padding, st, op)
        )
        x = pad_boundary(x, pre_dilation_padding, padding_mode, D)
    else:
        post = tuple((p_l, p_r + o) for (p_l, p_r), o in zip(padding, op))  # _conv.py:448-455
    y = conv_general_dilated(x, weight, (1,) * D, post, st, rd, groups)
    if bias is not None:
        y = y + bias.reshape((-1,) + (1,) * D).astype(y.dtype)
    return y


def physics_conv_transpose(x, weight, bias, *, boundary_mode, stride=1, output_padding=0, dilation=1, groups=1):
    """PhysicsConvTranspose (_physics_conv.py:121-236)."""
    D = weight.ndim - 2
    mode = boundary_mode.lower()
    k = weight.shape[2:]
    pad = compute_same_padding(D, k, dilation)
    return more_padding_conv_transpose(x, weight, bias, stride=stride, padding=pad, output_padding=output_padding,
                                       padding_mode=BOUNDARY_TO_PADDING[mode], dilation=dilation, groups=groups)


def init_conv_weights(rng, num_spatial_dims, in_channels, out_channels, kernel_size, groups=1, use_bias=True,
                      zero_bias_init=False, dtype=np.float32):
    """U(-lim, lim), lim = 1/sqrt(C_i/groups * prod(k)) for weight and bias (_conv.py:129-145)."""
    k = _ntuple(kernel_size, num_spatial_dims)
    cig = in_channels // groups
    lim = 1.0 / np.sqrt(cig * math.prod(k))
    w = rng.uniform(-lim, lim, (out_channels, cig) + k).astype(dtype)
    b = None
    if use_bias:
        b = rng.uniform(-lim, lim, (out_channels,) + (1,) * num_spatial_dims).astype(dtype)
        if zero_bias_init:
            b = np.zeros_like(b)
    return w, b
