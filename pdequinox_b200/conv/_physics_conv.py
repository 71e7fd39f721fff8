"""PhysicsConv / MorePaddingConv: boundary-aware convolution with "SAME" padding.

reference: pdequinox/conv/_conv.py:40-230 (MorePaddingConv), pdequinox/conv/_physics_conv.py:11-118.
The halo (periodic / dirichlet / neumann) is built inside the kernel instead of a padded copy.
"""
from __future__ import annotations

import math

import numpy as np

from .. import _lib, random as jr
from .._module import Module, conv_workspace, new_buf, precision_code, to_param


def _ntuple(v, n, name):
    if isinstance(v, (tuple, list)):
        if len(v) != n:
            raise ValueError(f"Length of {name} (length = {len(v)}) is not equal to {n}")
        return tuple(int(i) for i in v)
    return (int(v),) * n


def compute_same_padding(num_spatial_dims, kernel_size, dilation):
    """_physics_conv.py:11-26."""
    k = _ntuple(kernel_size, num_spatial_dims, "kernel_size")
    d = _ntuple(dilation, num_spatial_dims, "dilation")
    return tuple((((ki - 1) // 2) * di, (((ki - 1) // 2) + ((ki - 1) % 2)) * di) for ki, di in zip(k, d))


_PADDING_MODES = {"zeros": _lib.PAD_ZEROS, "circular": _lib.PAD_WRAP, "reflect": _lib.PAD_REFLECT,
                  "replicate": _lib.PAD_EDGE}


class MorePaddingConv(Module):
    """General N-D convolution with zeros / reflect / replicate / circular padding (_conv.py:40-230)."""
    _fields = ("weight", "bias")

    def __init__(self, num_spatial_dims, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 padding_mode="zeros", dilation=1, groups=1, use_bias=True, *, key):
        self.num_spatial_dims = num_spatial_dims
        self.kernel_size = _ntuple(kernel_size, num_spatial_dims, "kernel_size")
        self.stride = _ntuple(stride, num_spatial_dims, "stride")
        self.dilation = _ntuple(dilation, num_spatial_dims, "dilation")
        if in_channels % groups != 0:  # _conv.py:110-114
            raise ValueError(f"`in_channels` (={in_channels}) must be divisible by `groups` (={groups}).")
        if isinstance(padding, int):
            self.padding = tuple((padding, padding) for _ in range(num_spatial_dims))
        elif isinstance(padding, (tuple, list)) and len(padding) == num_spatial_dims:
            self.padding = tuple((p, p) if isinstance(p, int) else tuple(p) for p in padding)
        else:  # _conv.py:123-127
            raise ValueError(f"`padding` must either be an int or tuple of length {num_spatial_dims} containing ints or tuples of length 2.")
        if padding_mode not in _PADDING_MODES:
            raise ValueError("`padding_mode` must be one of ['zeros', 'reflect', 'replicate', 'circular']")
        self.padding_mode = padding_mode
        self.in_channels, self.out_channels, self.groups, self.use_bias = in_channels, out_channels, groups, use_bias
        wkey, bkey = jr.split(key, 2)  # _conv.py:116
        cig = in_channels // groups
        lim = 1.0 / math.sqrt(cig * math.prod(self.kernel_size))  # _conv.py:129-136
        self.weight = to_param(jr.uniform(wkey, (out_channels, cig) + self.kernel_size, -lim, lim))
        self.bias = to_param(jr.uniform(bkey, (out_channels,) + (1,) * num_spatial_dims, -lim, lim)) if use_bias else None

    @property
    def receptive_field(self):  # _conv.py:221-230
        return tuple((float(((k - 1) // 2) * d), float((k // 2) * d)) for k, d in zip(self.kernel_size, self.dilation))

    def _desc(self, x):
        D = self.num_spatial_dims
        if x.ndim != D + 2:
            raise ValueError(f"Input to `Conv` needs to have rank {D + 1}, but input has shape {tuple(x.shape[1:])}.")
        if x.shape[1] != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {x.shape[1]}")
        d = _lib.ConvDesc()
        d.ndim, d.batch, d.c_in, d.c_out, d.groups = D, x.shape[0], self.in_channels, self.out_channels, self.groups
        d.spatial = _lib.ivec(x.shape[2:], 1)
        d.kernel = _lib.ivec(self.kernel_size, 1)
        d.stride = _lib.ivec(self.stride, 1)
        d.dilation = _lib.ivec(self.dilation, 1)
        d.pad_lo = _lib.ivec([p[0] for p in self.padding], 0)
        d.pad_hi = _lib.ivec([p[1] for p in self.padding], 0)
        d.boundary = _PADDING_MODES[self.padding_mode]
        d.precision = precision_code()
        out = (_lib.C.c_int32 * _lib.MAX_DIMS)()
        _lib.check(_lib.lib().pdeqx_conv_out_spatial(d, out))
        return d, tuple(out[a] for a in range(D))

    def apply(self, x, tape, *, residual=None, act=_lib.ACT_IDENTITY, tag="y"):
        d, out_sp = self._desc(x)
        y = new_buf(tape, self, tag, (x.shape[0], self.out_channels) + out_sp)
        nws = _lib.lib().pdeqx_conv_workspace_bytes(d)
        _lib.check(_lib.lib().pdeqx_conv_fwd(d, _lib.ptr(x), _lib.ptr(self.weight), _lib.ptr(self.bias),
                                             _lib.ptr(residual), act, _lib.ptr(y), _lib.ptr(conv_workspace(tape, nws)), nws,
                                             _lib.stream()))
        return y

    def apply_bwd(self, x, gy, tape, need_gx=True, tag="gx", add=None, relu_mask=None):
        """Filter / bias gradients into the bound gradient buffers; returns the input gradient (+ `add`, the cotangent of a
        residual connection joining here), zeroed where `relu_mask` <= 0 (the output of the relu layer that produced x: its
        derivative is applied by the kernel that forms the cotangent)."""
        d, _ = self._desc(x)
        L = _lib.lib()
        gb = self.grad("bias") if self.bias is not None else None
        nws = L.pdeqx_conv_workspace_bytes(d)
        ws = _lib.ptr(conv_workspace(tape, nws))
        _lib.check(L.pdeqx_conv_bwd_filter(d, _lib.ptr(x), _lib.ptr(gy), _lib.ptr(self.grad("weight")), _lib.ptr(gb),
                                           ws, nws, _lib.stream()))
        if not need_gx:
            return None
        gx = new_buf(tape, self, tag, x.shape)
        _lib.check(L.pdeqx_conv_bwd_data(d, _lib.ptr(gy), _lib.ptr(self.weight), _lib.ptr(add), _lib.ptr(relu_mask), _lib.ptr(gx),
                                         ws, nws, _lib.stream()))
        return gx

    # ---- GroupNorm (groups = 1) in front of this convolution folded into its passes (pdeqx_conv_gn_*) ----
    def gn_fusable(self, x):
        """True when the norm -> conv -> act chain of _dilated_res_block.py:141-151 runs without the normalised tensor:
        the tcgen05 CTA-pair shapes with periodic boundary, TF32 mode."""
        d, _ = self._desc(x)
        return bool(_lib.lib().pdeqx_conv_gn_fusable(d))

    def apply_gn(self, x, norm, stats, tape, *, act=_lib.ACT_IDENTITY, tag="y", want_parts=True):
        """act(conv(norm(x)) + bias) from x itself; returns (y, parts, slots): per-sample partial sums (y, y^2) for the norm
        behind this layer."""
        d, out_sp = self._desc(x)
        L = _lib.lib()
        y = new_buf(tape, self, tag, (x.shape[0], self.out_channels) + out_sp)
        slots = L.pdeqx_conv_gn_slots(d)
        parts = new_buf(tape, self, "gn_parts", (x.shape[0] * slots * 2,)) if want_parts else None
        nws = L.pdeqx_conv_workspace_bytes(d)
        _lib.check(L.pdeqx_conv_gn_fwd(d, _lib.ptr(x), _lib.ptr(self.weight), _lib.ptr(self.bias), _lib.ptr(stats),
                                       _lib.ptr(norm.weight), _lib.ptr(norm.bias), act, _lib.ptr(y), _lib.ptr(parts),
                                       _lib.ptr(conv_workspace(tape, nws)), nws, _lib.stream()))
        return y, parts, slots

    def apply_bwd_gn(self, h, gy, norm, stats, tape, tag="gv"):
        """Filter / bias gradients against the (virtual) normalised input, and v = the cotangent of norm(h) with its
        per-sample sums (gamma v, gamma v h) for pdeqx_groupnorm_bwd_fused; returns (v, parts, slots)."""
        d, _ = self._desc(h)
        L = _lib.lib()
        gb = self.grad("bias") if self.bias is not None else None
        nws = L.pdeqx_conv_workspace_bytes(d)
        ws = _lib.ptr(conv_workspace(tape, nws))
        _lib.check(L.pdeqx_conv_gn_bwd_filter(d, _lib.ptr(h), _lib.ptr(gy), _lib.ptr(stats), _lib.ptr(norm.weight),
                                              _lib.ptr(norm.bias), _lib.ptr(self.grad("weight")), _lib.ptr(gb), ws, nws,
                                              _lib.stream()))
        slots = L.pdeqx_conv_gn_slots(d)
        v = new_buf(tape, self, tag, h.shape)
        parts = new_buf(tape, self, "gn_bparts", (h.shape[0] * slots * 2,))
        _lib.check(L.pdeqx_conv_gn_bwd_data(d, _lib.ptr(gy), _lib.ptr(self.weight), _lib.ptr(norm.weight), _lib.ptr(h),
                                            _lib.ptr(v), _lib.ptr(parts), ws, nws, _lib.stream()))
        return v, parts, slots

    def _fwd(self, x, tape):
        y = self.apply(x, tape)
        if tape is not None:
            tape.push(x)
        return y

    def _bwd(self, gy, tape, need_gx=True):
        return self.apply_bwd(tape.pop(), gy, tape, need_gx)


class PhysicsConv(MorePaddingConv):
    """_physics_conv.py:29-118."""

    def __init__(self, num_spatial_dims, in_channels, out_channels, kernel_size, stride=1, dilation=1, groups=1,
                 use_bias=True, *, key, boundary_mode, zero_bias_init=False):
        self.boundary_mode = boundary_mode.lower()
        if self.boundary_mode == "periodic":
            padding_mode = "circular"
        elif self.boundary_mode == "dirichlet":
            padding_mode = "zeros"
        elif self.boundary_mode == "neumann":
            padding_mode = "reflect"
        else:  # _physics_conv.py:98-101
            raise ValueError(f"boundary_mode={boundary_mode} not implemented")
        super().__init__(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels,
                         kernel_size=kernel_size, stride=stride,
                         padding=compute_same_padding(num_spatial_dims, kernel_size, dilation), dilation=dilation,
                         groups=groups, use_bias=use_bias, padding_mode=padding_mode, key=key)
        if use_bias and zero_bias_init:  # _physics_conv.py:117-118
            self.bias = to_param(np.zeros(tuple(self.bias.shape), np.float32))


class MorePaddingConvTranspose(Module):
    """General N-D transposed convolution with the reference's semantics (_conv.py:233-471): the weight
    (C_out, C_in/groups, *k) is used as is in a stride-1 correlation over the boundary-padded, stride-dilated input."""
    _fields = ("weight", "bias")

    def __init__(self, num_spatial_dims, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 padding_mode="zeros", output_padding=0, dilation=1, groups=1, use_bias=True, *, key):
        self.num_spatial_dims = num_spatial_dims
        self.kernel_size = _ntuple(kernel_size, num_spatial_dims, "kernel_size")
        self.stride = _ntuple(stride, num_spatial_dims, "stride")
        self.output_padding = _ntuple(output_padding, num_spatial_dims, "output_padding")
        self.dilation = _ntuple(dilation, num_spatial_dims, "dilation")
        if self.output_padding >= self.stride and any(o >= s for o, s in zip(self.output_padding, self.stride)):
            raise ValueError("Must have `output_padding < stride` (elementwise).")  # _conv.py:333-335
        if isinstance(padding, int):
            self.padding = tuple((padding, padding) for _ in range(num_spatial_dims))
        elif isinstance(padding, (tuple, list)) and len(padding) == num_spatial_dims:
            self.padding = tuple((p, p) if isinstance(p, int) else tuple(p) for p in padding)
        else:  # _conv.py:365-369
            raise ValueError(f"`padding` must either be an int or tuple of length {num_spatial_dims} containing ints or tuples of length 2.")
        if padding_mode not in _PADDING_MODES:
            raise ValueError("`padding_mode` must be one of ['zeros', 'reflect', 'replicate', 'circular']")
        self.padding_mode = padding_mode
        self.in_channels, self.out_channels, self.groups, self.use_bias = in_channels, out_channels, groups, use_bias
        wkey, bkey = jr.split(key, 2)  # _conv.py:321
        cig = in_channels // groups
        lim = 1.0 / math.sqrt(cig * math.prod(self.kernel_size))  # _conv.py:337-352
        self.weight = to_param(jr.uniform(wkey, (out_channels, cig) + self.kernel_size, -lim, lim))
        self.bias = to_param(jr.uniform(bkey, (out_channels,) + (1,) * num_spatial_dims, -lim, lim)) if use_bias else None

    @property
    def receptive_field(self):  # _conv.py:473-482
        return tuple((float(((k - 1) // 2) * d), float((k // 2) * d)) for k, d in zip(self.kernel_size, self.dilation))

    def _desc(self, x):
        D = self.num_spatial_dims
        if x.ndim != D + 2:
            raise ValueError(f"Input to `ConvTranspose` needs to have rank {D + 1}, but input has shape {tuple(x.shape[1:])}.")
        if x.shape[1] != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {x.shape[1]}")
        d = _lib.ConvTDesc()
        d.ndim, d.batch, d.c_in, d.c_out, d.groups = D, x.shape[0], self.in_channels, self.out_channels, self.groups
        d.spatial = _lib.ivec(x.shape[2:], 1)
        d.kernel = _lib.ivec(self.kernel_size, 1)
        d.stride = _lib.ivec(self.stride, 1)
        d.dilation = _lib.ivec(self.dilation, 1)
        d.pad_lo = _lib.ivec([p[0] for p in self.padding], 0)
        d.pad_hi = _lib.ivec([p[1] for p in self.padding], 0)
        d.output_padding = _lib.ivec(self.output_padding, 0)
        d.boundary = _PADDING_MODES[self.padding_mode]
        out = (_lib.C.c_int32 * _lib.MAX_DIMS)()
        _lib.check(_lib.lib().pdeqx_convT_out_spatial(d, out))
        return d, tuple(out[a] for a in range(D))

    def _fwd(self, x, tape):
        d, out_sp = self._desc(x)
        y = new_buf(tape, self, "y", (x.shape[0], self.out_channels) + out_sp)
        _lib.check(_lib.lib().pdeqx_convT_fwd(d, _lib.ptr(x), _lib.ptr(self.weight), _lib.ptr(self.bias), _lib.ptr(y),
                                              _lib.stream()))
        if tape is not None:
            tape.push(x)
        return y

    def _bwd(self, gy, tape, need_gx=True):
        x = tape.pop()
        d, _ = self._desc(x)
        L = _lib.lib()
        gb = self.grad("bias") if self.bias is not None else None
        _lib.check(L.pdeqx_convT_bwd_filter(d, _lib.ptr(x), _lib.ptr(gy), _lib.ptr(self.grad("weight")), _lib.ptr(gb),
                                            _lib.stream()))
        if not need_gx:
            return None
        gx = new_buf(tape, self, "gx", x.shape)
        _lib.check(L.pdeqx_convT_bwd_data(d, _lib.ptr(gy), _lib.ptr(self.weight), _lib.ptr(gx), _lib.stream()))
        return gx


class PhysicsConvTranspose(MorePaddingConvTranspose):
    """_physics_conv.py:121-236."""

    def __init__(self, num_spatial_dims, in_channels, out_channels, kernel_size, stride=1, output_padding=0, dilation=1,
                 groups=1, use_bias=True, *, key, boundary_mode, zero_bias_init=False):
        self.boundary_mode = boundary_mode.lower()
        if self.boundary_mode == "periodic":
            padding_mode = "circular"
        elif self.boundary_mode == "dirichlet":
            padding_mode = "zeros"
        elif self.boundary_mode == "neumann":
            padding_mode = "reflect"
        else:  # _physics_conv.py:211-215
            raise ValueError(f"Only 'periodic', 'dirichlet', 'neumann' boundary modes are supported, got {boundary_mode}")
        super().__init__(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels,
                         kernel_size=kernel_size, stride=stride,
                         padding=compute_same_padding(num_spatial_dims, kernel_size, dilation),
                         padding_mode=padding_mode, output_padding=output_padding, dilation=dilation, groups=groups,
                         use_bias=use_bias, key=key)
        if use_bias and zero_bias_init:  # _physics_conv.py:233-234
            self.bias = to_param(np.zeros(tuple(self.bias.shape), np.float32))
