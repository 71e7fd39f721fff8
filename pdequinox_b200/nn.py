"""Activation handles and the small third-party layers the reference takes from jax.nn / eqx.nn.

The hot-path kernels fuse the activation into their epilogues, so an activation is a named handle
(`gelu`, `relu`, `identity`) rather than an arbitrary callable; calling a handle on a tensor runs the
standalone kernel.
"""
from __future__ import annotations

from . import _lib
from ._module import Module, _as_device, new_buf


class Activation:
    def __init__(self, name, code):
        self.name, self.code = name, code

    def __call__(self, x):
        import torch
        x = _as_device(x)
        y = torch.empty_like(x)
        _lib.check(_lib.lib().pdeqx_activation_fwd(x.numel(), _lib.ptr(x), self.code, _lib.ptr(y), _lib.stream()))
        return y

    def __repr__(self):
        return f"pdequinox_b200.nn.{self.name}"


#: jax.nn.gelu (approximate=True: tanh form) -- ClassicFNO default (_classic_fno.py:20)
gelu = Activation("gelu", _lib.ACT_GELU_TANH)
#: jax.nn.relu -- ClassicResNet / DilatedResNet / ClassicUNet default
relu = Activation("relu", _lib.ACT_RELU)
identity = Activation("identity", _lib.ACT_IDENTITY)


def activation_code(fn) -> int:
    if isinstance(fn, Activation):
        return fn.code
    name = getattr(fn, "__name__", None)
    if name in ("gelu", "relu"):
        return {"gelu": _lib.ACT_GELU_TANH, "relu": _lib.ACT_RELU}[name]
    raise ValueError("activation must be pdequinox_b200.nn.gelu / relu / identity (fused into the kernels' epilogues)")


class Identity(Module):
    """eqx.nn.Identity."""
    _fields = ()

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, x, *, key=None):
        return x

    def _fwd(self, x, tape):
        return x

    def _bwd(self, gy, tape, need_gx=True):
        return gy


def apply_activation(x, act, tape, owner, tag):
    if act == _lib.ACT_IDENTITY:
        return x
    y = new_buf(tape, owner, tag, x.shape)
    _lib.check(_lib.lib().pdeqx_activation_fwd(x.numel(), _lib.ptr(x), act, _lib.ptr(y), _lib.stream()))
    return y


def activation_bwd(pre, gy, act, tape, owner, tag):
    """gy * act'(pre). For relu `pre` may be the activation OUTPUT (same sign pattern)."""
    if act == _lib.ACT_IDENTITY:
        return gy
    gx = new_buf(tape, owner, tag, gy.shape)
    _lib.check(_lib.lib().pdeqx_activation_bwd(gy.numel(), _lib.ptr(pre), _lib.ptr(gy), act, _lib.ptr(gx), _lib.stream()))
    return gx


ADD_PARTS_SLOTS = 64


def add_parts(a, b, tape, owner, tag):
    """a + b with per-sample partial sums (sum, sum of squares) of the result: (out, parts, slots) or (out, None, 0) when the
    per-sample size does not split evenly."""
    per_sample = a[0].numel()
    if per_sample % (4 * ADD_PARTS_SLOTS) != 0:
        return add(a, b, tape, owner, tag), None, 0
    out = new_buf(tape, owner, tag, a.shape)
    parts = new_buf(tape, owner, tag + "_parts", (a.shape[0] * ADD_PARTS_SLOTS * 2,))
    _lib.check(_lib.lib().pdeqx_add_parts(a.shape[0], per_sample, ADD_PARTS_SLOTS, _lib.ptr(a), _lib.ptr(b), _lib.ptr(out),
                                          _lib.ptr(parts), _lib.stream()))
    return out, parts, ADD_PARTS_SLOTS


def add(a, b, tape, owner, tag):
    out = new_buf(tape, owner, tag, a.shape)
    _lib.check(_lib.lib().pdeqx_add(a.numel(), _lib.ptr(a), _lib.ptr(b), _lib.ptr(out), _lib.stream()))
    return out


class GroupNorm(Module):
    """eqx.nn.GroupNorm(groups, channels, eps=1e-5, channelwise_affine=True): per-sample, per-group mean and
    biased variance over (channels/groups x spatial); per-channel weight (init 1) and bias (init 0)."""
    _fields = ("weight", "bias")

    def __init__(self, groups, channels, eps=1e-5, channelwise_affine=True):
        import numpy as np
        from ._module import to_param
        if channels % groups != 0:
            raise ValueError("The number of groups must divide the number of channels.")
        self.groups, self.channels, self.eps = groups, channels, eps
        self.weight = to_param(np.ones(channels, np.float32)) if channelwise_affine else None
        self.bias = to_param(np.zeros(channels, np.float32)) if channelwise_affine else None

    def _fwd(self, x, tape):
        B, Cc = x.shape[:2]
        n = x[0, 0].numel()
        y = new_buf(tape, self, "y", x.shape)
        stats = new_buf(tape, self, "stats", (B * self.groups * 2,))
        ws = new_buf(tape, self, "ws", (B * Cc * 4,))  # 2 doubles per (sample, channel)
        _lib.check(_lib.lib().pdeqx_groupnorm_fwd(B, Cc, self.groups, n, _lib.ptr(x), _lib.ptr(self.weight),
                                                  _lib.ptr(self.bias), self.eps, _lib.ptr(y), _lib.ptr(stats),
                                                  _lib.ptr(ws), _lib.stream()))
        if tape is not None:
            tape.push((x, stats))
        return y

    # ---- the pieces left of a norm that is folded into the convolution behind it (groups = 1; pdeqx_conv_gn_*) ----
    def stats_of(self, x, tape, tag="stats"):
        """(mean, rstd) per sample and group in a pass of its own -- for an input whose producer did not sum it."""
        B, Cc = x.shape[:2]
        stats = new_buf(tape, self, tag, (B * self.groups * 2,))
        ws = new_buf(tape, self, "ws", (B * Cc * 4,))
        _lib.check(_lib.lib().pdeqx_groupnorm_stats(B, Cc, self.groups, x[0, 0].numel(), _lib.ptr(x), self.eps,
                                                    _lib.ptr(stats), _lib.ptr(ws), _lib.stream()))
        return stats

    def stats_from_parts(self, parts, slots, batch, count, tape, tag="stats"):
        """(mean, rstd) per sample from the partial sums (x, x^2) its producer wrote in its epilogue."""
        stats = new_buf(tape, self, tag, (batch * 2,))
        _lib.check(_lib.lib().pdeqx_groupnorm_stats_from_parts(batch, slots, count, _lib.ptr(parts), self.eps,
                                                               _lib.ptr(stats), _lib.stream()))
        return stats

    def bwd_fused(self, h, v, stats, parts, slots, tape, *, mask_h=False, add=None):
        """Input gradient of the folded norm from v = cotangent of norm(h) (+ relu' of the layer whose output h is, + the
        cotangent `add` of a residual connection joining here) and the norm's parameter gradients, in one pass."""
        B, Cc = h.shape[:2]
        gx = new_buf(tape, self, "gx", h.shape)
        scratch = new_buf(tape, self, "scr", (B * Cc * 2,))
        gw = self.grad("weight") if self.weight is not None else None
        gb = self.grad("bias") if self.bias is not None else None
        _lib.check(_lib.lib().pdeqx_groupnorm_bwd_fused(B, Cc, h[0, 0].numel(), _lib.ptr(h), _lib.ptr(v), _lib.ptr(self.weight),
                                                        _lib.ptr(stats), _lib.ptr(parts), slots, 1 if mask_h else 0,
                                                        _lib.ptr(add), _lib.ptr(gx), _lib.ptr(gw), _lib.ptr(gb),
                                                        _lib.ptr(scratch), _lib.stream()))
        return gx

    def _bwd(self, gy, tape, need_gx=True, relu_mask=None):
        """relu_mask: the input gradient is zeroed where it is <= 0 (relu' of the layer whose output this norm read)."""
        x, stats = tape.pop()
        B, Cc = x.shape[:2]
        n = x[0, 0].numel()
        gx = new_buf(tape, self, "gx", x.shape)
        scratch = new_buf(tape, self, "scr", (B * Cc * 2,))
        gw = self.grad("weight") if self.weight is not None else None
        gb = self.grad("bias") if self.bias is not None else None
        _lib.check(_lib.lib().pdeqx_groupnorm_bwd(B, Cc, self.groups, n, _lib.ptr(x), _lib.ptr(self.weight),
                                                  _lib.ptr(stats), _lib.ptr(gy), _lib.ptr(relu_mask), _lib.ptr(gx), _lib.ptr(gw),
                                                  _lib.ptr(gb), _lib.ptr(scratch), _lib.stream()))
        return gx
