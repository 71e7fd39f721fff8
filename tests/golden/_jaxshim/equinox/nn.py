import math

import numpy as _np

import jax.lax as lax
import jax.random as jrandom


def field(*, static=False, default=None, **kw):
    return default


class Module:
    """Plain class: the reference only relies on attribute assignment in __init__ and __call__."""


class Identity(Module):
    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, x, *, key=None):
        return x


def _ntuple(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class Conv(Module):
    """eqx.nn.Conv: U(-lim, lim) init with lim = 1/sqrt(C_in/groups * prod(k)), wkey/bkey = split(key, 2);
    zero padding inside lax.conv_general_dilated; bias shape (C_out, 1, ..., 1)."""

    def __init__(self, num_spatial_dims, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, use_bias=True, padding_mode="ZEROS", dtype=None, *, key):
        wkey, bkey = jrandom.split(key, 2)
        self.num_spatial_dims = num_spatial_dims
        self.kernel_size = _ntuple(kernel_size, num_spatial_dims)
        self.stride = _ntuple(stride, num_spatial_dims)
        self.dilation = _ntuple(dilation, num_spatial_dims)
        self.padding = tuple((p, p) if isinstance(p, int) else tuple(p) for p in _ntuple(padding, num_spatial_dims))
        self.groups, self.use_bias = groups, use_bias
        self.in_channels, self.out_channels = in_channels, out_channels
        cig = in_channels // groups
        lim = 1 / math.sqrt(cig * math.prod(self.kernel_size))
        self.weight = jrandom.uniform(wkey, (out_channels, cig) + self.kernel_size, minval=-lim, maxval=lim)
        self.bias = jrandom.uniform(bkey, (out_channels,) + (1,) * num_spatial_dims, minval=-lim, maxval=lim) if use_bias else None

    def __call__(self, x, *, key=None):
        if x.ndim != self.num_spatial_dims + 1:
            raise ValueError(f"Input to `Conv` needs to have rank {self.num_spatial_dims + 1}, but input has shape {x.shape}.")
        y = lax.conv_general_dilated(lhs=x[None], rhs=self.weight, window_strides=self.stride, padding=self.padding,
                                     rhs_dilation=self.dilation, feature_group_count=self.groups)[0]
        if self.use_bias:
            y = y + self.bias
        return y


class ConvTranspose(Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("eqx.nn.ConvTranspose is not on the hot path")


class GroupNorm(Module):
    """eqx.nn.GroupNorm(groups, channels, eps=1e-5, channelwise_affine=True)."""

    def __init__(self, groups, channels=None, eps=1e-5, channelwise_affine=True, **kw):
        self.groups, self.channels, self.eps = groups, channels, eps
        self.weight = _np.ones(channels, _np.float32) if channelwise_affine else None
        self.bias = _np.zeros(channels, _np.float32) if channelwise_affine else None

    def __call__(self, x, state=None, *, key=None):
        y = x.reshape(self.groups, -1)
        mean = y.mean(axis=1, keepdims=True)
        var = y.var(axis=1, keepdims=True)
        out = ((y - mean) / _np.sqrt(var + self.eps)).reshape(x.shape)
        if self.weight is not None:
            shp = (-1,) + (1,) * (x.ndim - 1)
            out = self.weight.reshape(shp) * out + self.bias.reshape(shp)
        return out.astype(x.dtype)


class MLP(Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("eqx.nn.MLP is not on the hot path")
