"""PointwiseLinearConv: 1x1 convolution (reference: pdequinox/conv/_pointwise_linear_conv.py:6-53,
a subclass of eqx.nn.Conv with kernel_size=1, stride=1, padding=0)."""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _lib, random as jr
from .._module import Module, new_buf, to_param


class PointwiseLinearConv(Module):
    _fields = ("weight", "bias")

    def __init__(self, num_spatial_dims: int, in_channels: int, out_channels: int, use_bias: bool = True, *,
                 zero_bias_init: bool = False, key):
        self.num_spatial_dims = num_spatial_dims
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.use_bias = use_bias
        # eqx.nn.Conv init: wkey, bkey = split(key, 2); U(-lim, lim), lim = 1/sqrt(C_i * prod(k))
        wkey, bkey = jr.split(key, 2)
        lim = 1.0 / math.sqrt(in_channels)
        ones = (1,) * num_spatial_dims
        self.weight = to_param(jr.uniform(wkey, (out_channels, in_channels) + ones, -lim, lim))
        if use_bias:
            b = jr.uniform(bkey, (out_channels,) + ones, -lim, lim)
            if zero_bias_init:  # _pointwise_linear_conv.py:48-49
                b = np.zeros_like(b)
            self.bias = to_param(b)
        else:
            self.bias = None

    @property
    def receptive_field(self):
        return ((0.0, 0.0),) * self.num_spatial_dims

    def apply(self, x, tape, *, add=None, act=_lib.ACT_IDENTITY, tag="y"):
        B, C = x.shape[:2]
        if C != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {C}")
        n = int(np.prod(x.shape[2:]))
        y = new_buf(tape, self, tag, (B, self.out_channels) + tuple(x.shape[2:]))
        _lib.check(_lib.lib().pdeqx_pointwise_fwd(B, self.in_channels, self.out_channels, n, _lib.ptr(x),
                                                  _lib.ptr(self.weight), _lib.ptr(self.bias), _lib.ptr(add), act,
                                                  _lib.ptr(y), _lib.stream()))
        return y

    def apply_bwd(self, x, gy, tape, need_gx=True, tag="gx"):
        B = x.shape[0]
        n = int(np.prod(x.shape[2:]))
        gx = new_buf(tape, self, tag, x.shape) if need_gx else None
        gb = self.grad("bias") if self.bias is not None else None
        _lib.check(_lib.lib().pdeqx_pointwise_bwd(B, self.in_channels, self.out_channels, n, _lib.ptr(x),
                                                  _lib.ptr(self.weight), _lib.ptr(gy), _lib.ptr(gx),
                                                  _lib.ptr(self.grad("weight")), _lib.ptr(gb), _lib.stream()))
        return gx

    def bias_grad_only(self, gy, tape):
        """Reverse pass of a C -> 1 projection whose input was never stored (folded into the producing spectral block, which
        also accumulates the weight gradient): only the bias gradient sum_{b,pixel} gy is formed here."""
        if self.out_channels != 1:
            raise ValueError("bias_grad_only is the reverse pass of a C -> 1 projection")
        if self.bias is None:
            return
        B = gy.shape[0]
        n = int(np.prod(gy.shape[2:]))
        junk = new_buf(tape, self, "gw1", (1,))
        _lib.check(_lib.lib().pdeqx_pointwise_bwd(B, 1, 1, n, _lib.ptr(gy), _lib.ptr(self.weight), _lib.ptr(gy), None,
                                                  _lib.ptr(junk), _lib.ptr(self.grad("bias")), _lib.stream()))

    def _fwd(self, x, tape):
        y = self.apply(x, tape)
        if tape is not None:
            tape.push(x)
        return y

    def _bwd(self, gy, tape, need_gx=True):
        x = tape.pop()
        return self.apply_bwd(x, gy, tape, need_gx)
