"""The three hot-path ops as differentiable, vmappable JAX functions over the XLA FFI targets.

Structure of every op (SURVEY.md 7.2 #5 -- the integration risk this layout answers):

    op(x_one_sample, params..., static...)              jax.custom_vjp  (the reverse pass is the library's own)
        fwd  -> _fwd_call(static)(x, params...)         jax.custom_batching.custom_vmap around jax.ffi.ffi_call
        bwd  -> _bwd_call(static)(residuals, gy)        the same

  * un-batched call = the reference's convention (`model(x)`, x: (C, *S)): a batch of one;
  * under `jax.vmap` with only the DATA batched (what `jax.vmap(model)(x)` does: the parameters are closed over) the vmap
    rule hands the whole batch to ONE kernel launch -- the batch axis is a kernel dimension;
  * the reverse-pass call returns the parameter cotangents ALREADY REDUCED over the batch and declares them un-batched.
    jax's batching rule for custom_vjp only sums a cotangent over the batch axis if it comes back batched for an un-batched
    primal (`_match_axes_and_sum`), so the reduced arrays pass through unchanged: the 16.8 MB spectral weight gradient is
    never materialised per sample.
  * vmapping over PARAMETERS (ensembles, `eqx.filter_vmap` over models) is not a hot-path case: it raises
    NotImplementedError -- use the reference module there.
"""
from __future__ import annotations

import functools
from typing import NamedTuple

import jax
import jax.numpy as jnp
import numpy as np
from jax.custom_batching import custom_vmap

from . import _ffi

F32 = jnp.float32


def _sds(shape):
    return jax.ShapeDtypeStruct(tuple(int(s) for s in shape), F32)


def _absent():
    return jnp.zeros((0,), F32)  # zero-element operand = "not given" (the handlers' convention)


def _opt(a):
    return _absent() if a is None else a


def _require_params_unbatched(in_batched, first_param, name):
    if any(in_batched[first_param:]):
        raise NotImplementedError(f"{name}: vmap over the parameters is not supported by the B200 kernels (only the data may "
                                  "be batched, as in jax.vmap(model)(x)); use the reference module for ensembles")


def _lead(x, batched, axis_size):
    return x if batched else jnp.broadcast_to(x, (axis_size,) + x.shape)


# ------------------------------------------------------------------------------------------
# SpectralConv / ClassicSpectralBlock          (pdequinox/conv/_spectral_conv.py:107-136, blocks/_classic_spectral_block.py:72-75)
# ------------------------------------------------------------------------------------------
def _spectral_fwd_batched(x, w_re, w_im, byp_w, byp_b, activation, precision):
    _ffi.register()
    B, ci, *S = x.shape
    co, modes = w_re.shape[1], w_re.shape[3:]
    nx, nz, nws = _ffi.spectral_sizes(len(S), B, ci, co, S, modes, precision)
    out = (_sds((B, co, *S)), _sds((nx,)), _sds((nz,)), _sds((nws // 4 + 1,)))
    y, xhat, z, _ = jax.ffi.ffi_call("pdeqx_spectral_block_fwd", out)(
        x, w_re, w_im, byp_w, byp_b, activation=_ffi.i32(activation), precision=_ffi.i32(precision))
    return y, xhat.reshape(B, -1), z.reshape(B, -1)  # saved tensors are batch-major: expose the batch axis


def _spectral_bwd_batched(x, gy, w_re, w_im, byp_w, byp_b, xhat, z, activation, precision):
    _ffi.register()
    B, ci, *S = x.shape
    co, modes = w_re.shape[1], w_re.shape[3:]
    _, _, nws = _ffi.spectral_sizes(len(S), B, ci, co, S, modes, precision)
    out = (_sds(x.shape), _sds(w_re.shape), _sds(w_im.shape), _sds(byp_w.shape), _sds(byp_b.shape), _sds(gy.shape),
           _sds((nws // 4 + 1,)))
    gx, gwr, gwi, gbw, gbb, _, _ = jax.ffi.ffi_call("pdeqx_spectral_block_bwd", out)(
        x, gy, w_re, w_im, byp_w, byp_b, xhat.reshape(-1), z.reshape(-1), activation=_ffi.i32(activation),
        precision=_ffi.i32(precision))
    return gx, gwr, gwi, gbw, gbb


@functools.lru_cache(maxsize=None)
def _spectral_fwd_call(activation, precision):
    @custom_vmap
    def call(x, w_re, w_im, byp_w, byp_b):
        y, xhat, z = _spectral_fwd_batched(x[None], w_re, w_im, byp_w, byp_b, activation, precision)
        return y[0], xhat[0], z[0]

    @call.def_vmap
    def rule(axis_size, in_batched, x, w_re, w_im, byp_w, byp_b):
        _require_params_unbatched(in_batched, 1, "spectral_block")
        out = _spectral_fwd_batched(_lead(x, in_batched[0], axis_size), w_re, w_im, byp_w, byp_b, activation, precision)
        return out, (True, True, True)

    return call


@functools.lru_cache(maxsize=None)
def _spectral_bwd_call(activation, precision):
    @custom_vmap
    def call(x, gy, w_re, w_im, byp_w, byp_b, xhat, z):
        gx, *gp = _spectral_bwd_batched(x[None], gy[None], w_re, w_im, byp_w, byp_b, xhat[None], z[None], activation, precision)
        return (gx[0], *gp)

    @call.def_vmap
    def rule(axis_size, in_batched, x, gy, w_re, w_im, byp_w, byp_b, xhat, z):
        _require_params_unbatched(in_batched[:6], 2, "spectral_block (reverse pass)")
        out = _spectral_bwd_batched(_lead(x, in_batched[0], axis_size), _lead(gy, in_batched[1], axis_size), w_re, w_im,
                                    byp_w, byp_b, _lead(xhat, in_batched[6], axis_size), _lead(z, in_batched[7], axis_size),
                                    activation, precision)
        return out, (True, False, False, False, False)  # parameter cotangents: reduced over the batch inside the kernel

    return call


@functools.partial(jax.custom_vjp, nondiff_argnums=(5, 6))
def _spectral_block(x, w_re, w_im, byp_w, byp_b, activation, precision):
    return _spectral_fwd_call(activation, precision)(x, w_re, w_im, byp_w, byp_b)[0]


def _spectral_block_fwd(x, w_re, w_im, byp_w, byp_b, activation, precision):
    y, xhat, z = _spectral_fwd_call(activation, precision)(x, w_re, w_im, byp_w, byp_b)
    return y, (x, w_re, w_im, byp_w, byp_b, xhat, z)


def _spectral_block_bwd(activation, precision, res, gy):
    x, w_re, w_im, byp_w, byp_b, xhat, z = res
    return _spectral_bwd_call(activation, precision)(x, gy, w_re, w_im, byp_w, byp_b, xhat, z)


_spectral_block.defvjp(_spectral_block_fwd, _spectral_block_bwd)


def spectral_block(x, weights_real, weights_imag, bypass_weight=None, bypass_bias=None, *, activation="identity"):
    """activation(spectral_conv(x) [+ bypass_weight . x + bypass_bias]) for ONE sample x: (C_in, *S) -> (C_out, *S).
    weights: (G, C_out, C_in, *modes) as the reference interprets them at call time (_spectral_conv.py:129-134)."""
    if x.ndim != weights_real.ndim - 2:
        raise ValueError(f"expected one sample of rank {weights_real.ndim - 2} (C, *S), got shape {x.shape}")
    return _spectral_block(x, weights_real, weights_imag, _opt(bypass_weight), _opt(bypass_bias), _ffi.ACT[activation],
                           _ffi.get_precision())


# ------------------------------------------------------------------------------------------
# PhysicsConv / MorePaddingConv                (pdequinox/conv/_conv.py:169-219, _physics_conv.py:11-26,92-101)
# ------------------------------------------------------------------------------------------
class ConvCfg(NamedTuple):
    stride: tuple
    dilation: tuple
    pad_lo: tuple
    pad_hi: tuple
    boundary: int
    groups: int
    precision: int


def _conv_attrs(cfg, activation=None):
    a = dict(stride=_ffi.ints(cfg.stride), dilation=_ffi.ints(cfg.dilation), pad_lo=_ffi.ints(cfg.pad_lo), pad_hi=_ffi.ints(cfg.pad_hi),
             boundary=_ffi.i32(cfg.boundary), groups=_ffi.i32(cfg.groups), precision=_ffi.i32(cfg.precision))
    if activation is not None:
        a["activation"] = _ffi.i32(activation)
    return a


def _conv_fwd_batched(x, w, b, cfg):
    _ffi.register()
    B, ci, *S = x.shape
    out_sp, nws = _ffi.conv_out_and_workspace(B, ci, w.shape[0], S, w.shape[2:], cfg)
    y, _ = jax.ffi.ffi_call("pdeqx_conv_fwd", (_sds((B, w.shape[0], *out_sp)), _sds((nws,))))(
        x, w, b, _absent(), **_conv_attrs(cfg, _ffi.ACT["identity"]))
    return y


def _conv_bwd_batched(x, gy, w, b, cfg):
    _ffi.register()
    B, ci, *S = x.shape
    _, nws = _ffi.conv_out_and_workspace(B, ci, w.shape[0], S, w.shape[2:], cfg)
    gx, gw, gb, _ = jax.ffi.ffi_call("pdeqx_conv_bwd", (_sds(x.shape), _sds(w.shape), _sds(b.shape), _sds((nws,))))(
        x, gy, w, _absent(), **_conv_attrs(cfg))
    return gx, gw, gb


@functools.lru_cache(maxsize=None)
def _conv_fwd_call(cfg):
    @custom_vmap
    def call(x, w, b):
        return _conv_fwd_batched(x[None], w, b, cfg)[0]

    @call.def_vmap
    def rule(axis_size, in_batched, x, w, b):
        _require_params_unbatched(in_batched, 1, "physics_conv")
        return _conv_fwd_batched(_lead(x, in_batched[0], axis_size), w, b, cfg), True

    return call


@functools.lru_cache(maxsize=None)
def _conv_bwd_call(cfg):
    @custom_vmap
    def call(x, gy, w, b):
        gx, gw, gb = _conv_bwd_batched(x[None], gy[None], w, b, cfg)
        return gx[0], gw, gb

    @call.def_vmap
    def rule(axis_size, in_batched, x, gy, w, b):
        _require_params_unbatched(in_batched, 2, "physics_conv (reverse pass)")
        out = _conv_bwd_batched(_lead(x, in_batched[0], axis_size), _lead(gy, in_batched[1], axis_size), w, b, cfg)
        return out, (True, False, False)

    return call


@functools.partial(jax.custom_vjp, nondiff_argnums=(3,))
def _physics_conv(x, w, b, cfg):
    return _conv_fwd_call(cfg)(x, w, b)


def _physics_conv_fwd(x, w, b, cfg):
    return _conv_fwd_call(cfg)(x, w, b), (x, w, b)


def _physics_conv_bwd(cfg, res, gy):
    x, w, b = res
    return _conv_bwd_call(cfg)(x, gy, w, b)


_physics_conv.defvjp(_physics_conv_fwd, _physics_conv_bwd)


def physics_conv(x, weight, bias, *, stride, dilation, padding, padding_mode, groups=1):
    """corr(pad(x), weight) + bias for ONE sample x: (C_in, *S); `padding` = ((lo, hi), ...) per axis, `padding_mode` one of
    zeros / circular / reflect / replicate (MorePaddingConv's vocabulary, _conv.py:189-200). The halo is an index map inside
    the kernel: no padded copy of x is made."""
    cfg = ConvCfg(tuple(stride), tuple(dilation), tuple(p[0] for p in padding), tuple(p[1] for p in padding),
                  _ffi.BOUNDARY[padding_mode], int(groups), _ffi.get_precision())
    return _physics_conv(x, weight, _opt(bias), cfg)


# ------------------------------------------------------------------------------------------
# GroupNorm(1, C) -> PhysicsConv -> activation as one op   (pdequinox/blocks/_dilated_res_block.py:141-151)
#   the norm is folded into the convolution's passes (pdeqx_conv_gn_*): the normalised tensor never exists. Accepted shapes:
#   pdeqx_conv_gn_fusable (periodic, the tcgen05 CTA-pair shapes, TF32 precision); callers fall back to
#   group_norm + physics_conv elsewhere (`norm_conv_act_fusable`).
# ------------------------------------------------------------------------------------------
def norm_conv_act_fusable(batch, c_in, c_out, spatial, kernel, cfg):
    return bool(_ffi._core.lib().pdeqx_conv_gn_fusable(_ffi.conv_desc(batch, c_in, c_out, spatial, kernel, cfg)))


def _norm_conv_sizes(x, w, cfg):
    B, ci, *S = x.shape
    d = _ffi.conv_desc(B, ci, w.shape[0], S, w.shape[2:], cfg)
    L = _ffi._core.lib()
    slots = L.pdeqx_conv_gn_slots(d)
    if slots == 0:
        raise NotImplementedError("norm_conv_act: this convolution does not take the folded GroupNorm passes "
                                  "(pdeqx_conv_gn_fusable): use group_norm + physics_conv")
    nws = max(L.pdeqx_conv_workspace_bytes(d), L.pdeqx_groupnorm_workspace_bytes(B, ci)) // 4 + 1
    return slots, nws


def _norm_conv_fwd_batched(x, gn_w, gn_b, w, b, cfg, activation, eps):
    _ffi.register()
    B = x.shape[0]
    slots, nws = _norm_conv_sizes(x, w, cfg)
    y, stats, _, _ = jax.ffi.ffi_call("pdeqx_norm_conv_fwd", (_sds((B, w.shape[0], *x.shape[2:])), _sds((B, 2)), _sds((B, slots, 2)),
                                                              _sds((nws,))))(
        x, gn_w, gn_b, w, b, _absent(), eps=np.float32(eps), **_conv_attrs(cfg, activation))
    return y, stats


def _norm_conv_bwd_batched(x, stats, gn_w, gn_b, w, b, ga, cfg):
    _ffi.register()
    B, ci = x.shape[:2]
    slots, nws = _norm_conv_sizes(x, w, cfg)
    gx, ggw, ggb, gw, gb, *_ = jax.ffi.ffi_call(
        "pdeqx_norm_conv_bwd", (_sds(x.shape), _sds(gn_w.shape), _sds(gn_b.shape), _sds(w.shape), _sds(b.shape), _sds(x.shape),
                                _sds((B, slots, 2)), _sds((2 * B * ci,)), _sds((nws,))))(
        x, stats, gn_w, gn_b, w, ga, _absent(), relu_mask_x=_ffi.i32(0), **_conv_attrs(cfg))
    return gx, ggw, ggb, gw, gb


@functools.lru_cache(maxsize=None)
def _norm_conv_fwd_call(cfg, activation, eps):
    @custom_vmap
    def call(x, gn_w, gn_b, w, b):
        y, stats = _norm_conv_fwd_batched(x[None], gn_w, gn_b, w, b, cfg, activation, eps)
        return y[0], stats[0]

    @call.def_vmap
    def rule(axis_size, in_batched, x, gn_w, gn_b, w, b):
        _require_params_unbatched(in_batched, 1, "norm_conv_act")
        return _norm_conv_fwd_batched(_lead(x, in_batched[0], axis_size), gn_w, gn_b, w, b, cfg, activation, eps), (True, True)

    return call


@functools.lru_cache(maxsize=None)
def _norm_conv_bwd_call(cfg):
    @custom_vmap
    def call(x, stats, ga, gn_w, gn_b, w, b):
        gx, *gp = _norm_conv_bwd_batched(x[None], stats[None], gn_w, gn_b, w, b, ga[None], cfg)
        return (gx[0], *gp)

    @call.def_vmap
    def rule(axis_size, in_batched, x, stats, ga, gn_w, gn_b, w, b):
        _require_params_unbatched(in_batched, 3, "norm_conv_act (reverse pass)")
        out = _norm_conv_bwd_batched(_lead(x, in_batched[0], axis_size), _lead(stats, in_batched[1], axis_size), gn_w, gn_b, w, b,
                                     _lead(ga, in_batched[2], axis_size), cfg)
        return out, (True, False, False, False, False)

    return call


@functools.partial(jax.custom_vjp, nondiff_argnums=(5, 6, 7))
def _norm_conv_act(x, gn_w, gn_b, w, b, cfg, activation, eps):
    return _norm_conv_fwd_call(cfg, activation, eps)(x, gn_w, gn_b, w, b)[0]


def _norm_conv_act_fwd(x, gn_w, gn_b, w, b, cfg, activation, eps):
    y, stats = _norm_conv_fwd_call(cfg, activation, eps)(x, gn_w, gn_b, w, b)
    return y, (x, stats, y, gn_w, gn_b, w, b)


def _norm_conv_act_bwd(cfg, activation, eps, res, gy):
    x, stats, y, gn_w, gn_b, w, b = res
    ga = jnp.where(y > 0, gy, 0) if activation == _ffi.ACT["relu"] else gy  # cotangent of the pre-activation
    return _norm_conv_bwd_call(cfg)(x, stats, ga, gn_w, gn_b, w, b)


_norm_conv_act.defvjp(_norm_conv_act_fwd, _norm_conv_act_bwd)


def norm_conv_act(x, norm_weight, norm_bias, weight, bias, *, stride, dilation, padding, padding_mode, activation="relu",
                  eps=1e-5):
    """activation(corr(pad(group_norm_1(x)), weight) + bias) for ONE sample x: (C, *S) -- the layer of DilatedResBlock
    (norm -> conv -> act) as one custom call each way. activation: relu or identity (the reverse pass needs act' from the
    output alone). In this functional binding every layer takes its statistics in a pass of its own; the torch host side
    (pdequinox_b200/blocks/_dilated_res_block.py) additionally chains the producer's partial sums from layer to layer."""
    if activation not in ("relu", "identity"):
        raise NotImplementedError("norm_conv_act: relu or identity")
    cfg = ConvCfg(tuple(stride), tuple(dilation), tuple(p[0] for p in padding), tuple(p[1] for p in padding),
                  _ffi.BOUNDARY[padding_mode], 1, _ffi.get_precision())
    return _norm_conv_act(x, norm_weight, norm_bias, weight, _opt(bias), cfg, _ffi.ACT[activation], float(eps))


# ------------------------------------------------------------------------------------------
# PointwiseLinearConv                          (pdequinox/conv/_pointwise_linear_conv.py:37-49)
# ------------------------------------------------------------------------------------------
def _pw_fwd_batched(x, w, b):
    _ffi.register()
    return jax.ffi.ffi_call("pdeqx_pointwise_fwd", _sds((x.shape[0], w.shape[0], *x.shape[2:])))(
        x, w, b, _absent(), activation=_ffi.i32(_ffi.ACT["identity"]))


def _pw_bwd_batched(x, w, b, gy):
    _ffi.register()
    return jax.ffi.ffi_call("pdeqx_pointwise_bwd", (_sds(x.shape), _sds(w.shape), _sds(b.shape)))(x, w, gy)


@custom_vmap
def _pw_fwd_call(x, w, b):
    return _pw_fwd_batched(x[None], w, b)[0]


@_pw_fwd_call.def_vmap
def _pw_fwd_rule(axis_size, in_batched, x, w, b):
    _require_params_unbatched(in_batched, 1, "pointwise_conv")
    return _pw_fwd_batched(_lead(x, in_batched[0], axis_size), w, b), True


@custom_vmap
def _pw_bwd_call(x, w, b, gy):
    gx, gw, gb = _pw_bwd_batched(x[None], w, b, gy[None])
    return gx[0], gw, gb


@_pw_bwd_call.def_vmap
def _pw_bwd_rule(axis_size, in_batched, x, w, b, gy):
    _require_params_unbatched(in_batched[:3], 1, "pointwise_conv (reverse pass)")
    out = _pw_bwd_batched(_lead(x, in_batched[0], axis_size), w, b, _lead(gy, in_batched[3], axis_size))
    return out, (True, False, False)


@jax.custom_vjp
def _pointwise(x, w, b):
    return _pw_fwd_call(x, w, b)


_pointwise.defvjp(lambda x, w, b: (_pw_fwd_call(x, w, b), (x, w, b)), lambda res, gy: _pw_bwd_call(*res, gy))


def pointwise_conv(x, weight, bias=None):
    """weight . x + bias (1x1 convolution) for ONE sample x: (C_in, *S)."""
    return _pointwise(x, weight, _opt(bias))
