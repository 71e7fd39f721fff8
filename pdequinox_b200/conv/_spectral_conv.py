"""SpectralConv: FNO spectral convolution as a fused truncated-DFT pipeline.

reference: pdequinox/conv/_spectral_conv.py:10-136 (rfftn -> per-corner complex einsum -> irfftn).
Only the retained modes are computed (dense twiddle GEMMs); the zero-filled spectrum never exists.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib, random as jr
from .._module import Module, default_device, new_buf, precision_code, to_param


class _Plan:
    def __init__(self, desc):
        self.handle = C.c_void_p()
        _lib.check(_lib.lib().pdeqx_spectral_plan_create(desc, C.byref(self.handle)))
        L = _lib.lib()
        self.ws_bytes = L.pdeqx_spectral_workspace_bytes(self.handle)
        self.xhat_floats = L.pdeqx_spectral_xhat_floats(self.handle)
        self.z_floats = L.pdeqx_spectral_z_floats(self.handle)
        self.workspace = torch.empty(self.ws_bytes // 4 + 1, dtype=torch.float32, device=default_device())

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().pdeqx_spectral_plan_destroy(self.handle)
        except Exception:
            pass


_PLANS = {}


def get_plan(D, batch, c_in, c_out, spatial, modes):
    key = (D, batch, c_in, c_out, tuple(spatial), tuple(modes), precision_code(), torch.cuda.current_device())
    p = _PLANS.get(key)
    if p is None:
        d = _lib.SpectralDesc()
        d.ndim, d.batch, d.c_in, d.c_out = D, batch, c_in, c_out
        d.spatial = _lib.ivec(spatial, 1)
        d.modes = _lib.ivec(modes, 1)
        d.precision = precision_code()
        p = _PLANS[key] = _Plan(d)
    return p


class SpectralConv(Module):
    _fields = ("num_spatial_dims", "num_modes", "weights_real", "weights_imag")

    def __init__(self, num_spatial_dims: int, in_channels: int, out_channels: int, num_modes, *, key):
        if isinstance(num_modes, int):
            num_modes = (num_modes,) * num_spatial_dims
        if len(num_modes) != num_spatial_dims:  # _spectral_conv.py:48-49
            raise ValueError("num_modes must have the same length as num_spatial_dims")
        self.num_spatial_dims = num_spatial_dims
        self.num_modes = tuple(int(m) for m in num_modes)
        # allocation shape (G, in, out, *modes) (_spectral_conv.py:54-58); the call interprets it as (G, o, i, *modes)
        weight_shape = (2 ** (num_spatial_dims - 1), in_channels, out_channels) + self.num_modes
        real_key, imag_key = jr.split(key)  # _spectral_conv.py:60
        scale = 1.0 / (in_channels * out_channels)
        self.weights_real = to_param(scale * jr.normal(real_key, weight_shape))
        self.weights_imag = to_param(scale * jr.normal(imag_key, weight_shape))

    @property
    def receptive_field(self):
        return ((float("inf"), float("inf")),) * self.num_spatial_dims

    # call-time channel interpretation: `_, o, *_ = weight.shape`, einsum "i...,oi...->o..." (_spectral_conv.py:129-134)
    @property
    def call_out_channels(self):
        return int(self.weights_real.shape[1])

    @property
    def call_in_channels(self):
        return int(self.weights_real.shape[2])

    def apply(self, x, tape, *, bypass=None, act=_lib.ACT_IDENTITY, save=False):
        """y = act(spectral(x) [+ bypass(x)]); returns (y, saved) with saved = (plan, xhat, z) when `save`."""
        D = self.num_spatial_dims
        if x.ndim != D + 2:
            raise ValueError(f"Input to `SpectralConv` needs to have rank {D + 1}, but input has shape {tuple(x.shape[1:])}.")
        B, ci = x.shape[:2]
        co = self.call_out_channels
        if ci != self.call_in_channels:
            raise ValueError(f"SpectralConv: input has {ci} channels but the weight contracts over {self.call_in_channels}")
        spatial = tuple(x.shape[2:])
        plan = get_plan(D, B, ci, co, spatial, self.num_modes)
        y = new_buf(tape, self, "y", (B, co) + spatial)
        xhat = z = None
        if save:
            xhat = new_buf(tape, self, "xhat", (plan.xhat_floats,))
            z = new_buf(tape, self, "z", (plan.z_floats,))
        bw = bypass.weight if bypass is not None else None
        bb = bypass.bias if bypass is not None else None
        _lib.check(_lib.lib().pdeqx_spectral_block_fwd(
            plan.handle, _lib.ptr(x), _lib.ptr(self.weights_real), _lib.ptr(self.weights_imag), _lib.ptr(bw),
            _lib.ptr(bb), act, _lib.ptr(y), _lib.ptr(xhat), _lib.ptr(z), _lib.ptr(plan.workspace), plan.ws_bytes,
            _lib.stream()))
        return y, (plan, xhat, z)

    def apply_ex(self, x, tape, *, bypass, act, save=False, lifting=None, proj=None):
        """The block's forward pass with the pointwise ends of the architecture folded in (pdeqx_spectral_block_fwd_ex).
        `lifting = (lift_w, lift_b)`: `x` is the single-channel field u [B, 1, *S] and the block's input is the 1 -> C
        lifting lift_w[i] * u + lift_b[i], never materialised. `proj = (proj_w, proj_b)`: the block's output only feeds a
        C -> 1 pointwise projection; the projected field [B, 1, *S] is returned and the block's output is never written.
        Returns (y or projected field, saved)."""
        D = self.num_spatial_dims
        B = x.shape[0]
        ci, co = self.call_in_channels, self.call_out_channels
        spatial = tuple(x.shape[2:])
        plan = get_plan(D, B, ci, co, spatial, self.num_modes)
        y = out = None
        if proj is None:
            y = new_buf(tape, self, "y", (B, co) + spatial)
        else:
            out = new_buf(tape, self, "proj", (B, 1) + spatial)
        xhat = z = None
        if save:
            xhat = new_buf(tape, self, "xhat", (plan.xhat_floats,))
            z = new_buf(tape, self, "z", (plan.z_floats,))
        lw, lb = lifting if lifting is not None else (None, None)
        pw, pb = proj if proj is not None else (None, None)
        _lib.check(_lib.lib().pdeqx_spectral_block_fwd_ex(
            plan.handle, _lib.ptr(None if lifting is not None else x), _lib.ptr(x if lifting is not None else None),
            _lib.ptr(lw), _lib.ptr(lb), _lib.ptr(pw), _lib.ptr(pb), _lib.ptr(out), _lib.ptr(self.weights_real),
            _lib.ptr(self.weights_imag), _lib.ptr(bypass.weight), _lib.ptr(bypass.bias), act, _lib.ptr(y), _lib.ptr(xhat),
            _lib.ptr(z), _lib.ptr(plan.workspace), plan.ws_bytes, _lib.stream()))
        return (out if proj is not None else y), (plan, xhat, z)

    def apply_lifted(self, u, lift_w, lift_b, tape, *, bypass, act, save=False):
        """`apply_ex` with only the lifting folded in."""
        return self.apply_ex(u, tape, bypass=bypass, act=act, save=save, lifting=(lift_w, lift_b))

    def projected_ok(self, x, act):
        """True if `apply_ex(..., proj=...)` can run for a batch / grid of this shape (tcgen05 slab path with a rank-one
        cotangent in the reverse pass, at least 2 spatial dims, at most 64 output channels)."""
        D = self.num_spatial_dims
        if D < 2 or not x.is_cuda or act == _lib.ACT_IDENTITY or self.call_out_channels > 64:
            return False
        plan = get_plan(D, x.shape[0], self.call_in_channels, self.call_out_channels, tuple(x.shape[2:]), self.num_modes)
        return bool(_lib.lib().pdeqx_spectral_block_rank1_ok(plan.handle, 1, act))

    def lifted_ok(self, u, act):
        """True if `apply_lifted` can run for a field of this shape (tcgen05 path, at least 2 spatial dims)."""
        D = self.num_spatial_dims
        if D < 2 or not u.is_cuda or act == _lib.ACT_IDENTITY:
            return False
        plan = get_plan(D, u.shape[0], self.call_in_channels, self.call_out_channels, tuple(u.shape[2:]), self.num_modes)
        L = _lib.lib()
        return bool(L.pdeqx_spectral_block_rank1_ok(plan.handle, 1, act)) and (L.pdeqx_spectral_plan_tcgen05_stages(plan.handle) & 1) == 1

    def apply_bwd(self, x, gy, saved, tape, *, bypass=None, act=_lib.ACT_IDENTITY, need_gx=True, gy_w=None, lift=None,
                  lifted=None, proj_gw=None, gy_is_gpre=False, chain=None):
        """VJP (pdeqx_spectral_block_bwd_args). With `gy_w` the cotangent is rank one: gy_w[o] * gy[b, 0, *S] (what a C -> 1
        pointwise projection sends back); it is then formed inside the kernel instead of being materialised. With
        `lift = (u, gw, gb)` the input gradient is not stored: the gradients gw, gb of the 1 -> C pointwise lifting that made
        x from the field u are accumulated instead. With `lifted = (lift_w, lift_b)` the forward pass was `apply_lifted`: x is
        None and the reverse pass works from u alone. With `proj_gw` (needs `gy_w`) the forward pass was
        `apply_ex(..., proj=...)`: the projection's weight gradient is accumulated into it from the recomputed activation.
        `gy_is_gpre`: gy already is the cotangent of this block's pre-activation (produced by the chained call of the block
        behind it). `chain = dict(conv, x, saved, bypass, act, lifted, u)` describes the spectral block IN FRONT of this one:
        instead of this block's input gradient, the cotangent of that block's pre-activation is formed in the same pass and
        returned (it lives in that block's own `gpre` buffer)."""
        plan, xhat, z = saved
        ref = x if x is not None else lift[0]
        gx = new_buf(tape, self, "gx", x.shape) if (need_gx and lift is None and chain is None) else None
        full = (ref.shape[0], self.call_out_channels) + tuple(ref.shape[2:])
        scratch = new_buf(tape, self, "gpre", full) if (act != _lib.ACT_IDENTITY and not gy_is_gpre) else None
        a = _lib.SpectralBwdArgs()
        P = _lib.ptr
        a.x, a.gy, a.gy_w, a.proj_gw = P(x), P(gy), P(gy_w), P(proj_gw)
        if lift is not None:
            a.lift_u, a.lift_gw, a.lift_gb = P(lift[0]), P(lift[1]), P(lift[2])
        if lifted is not None:
            a.lift_w, a.lift_b = P(lifted[0]), P(lifted[1])
        a.w_re, a.w_im = P(self.weights_real), P(self.weights_imag)
        if bypass is not None:
            a.bypass_w, a.bypass_b = P(bypass.weight), P(bypass.bias)
            a.g_bypass_w = P(bypass.grad("weight"))
            a.g_bypass_b = P(bypass.grad("bias")) if bypass.bias is not None else None
        a.activation = act
        a.saved_xhat, a.saved_z, a.gx = P(xhat), P(z), P(gx)
        a.gw_re, a.gw_im = P(self.grad("weights_real")), P(self.grad("weights_imag"))
        a.scratch_full = P(scratch)
        a.workspace, a.workspace_bytes, a.stream = P(plan.workspace), plan.ws_bytes, _lib.stream()
        a.gy_is_gpre = 1 if gy_is_gpre else 0
        out = gx
        if chain is not None:
            up = chain["conv"]
            up_plan, _, up_z = chain["saved"]
            up_ref = chain["x"] if chain["x"] is not None else chain["u"]
            up_gpre = new_buf(tape, up, "gpre", (up_ref.shape[0], up.call_out_channels) + tuple(up_ref.shape[2:]))
            a.up_plan, a.up_x, a.up_saved_z = up_plan.handle, P(chain["x"]), P(up_z)
            a.up_bypass_w, a.up_bypass_b = P(chain["bypass"].weight), P(chain["bypass"].bias)
            a.up_activation = chain["act"]
            if chain["x"] is None:
                a.up_lift_u, a.up_lift_w, a.up_lift_b = P(chain["u"]), P(chain["lifted"][0]), P(chain["lifted"][1])
            a.up_gpre = P(up_gpre)
            out = up_gpre
        _lib.check(_lib.lib().pdeqx_spectral_block_bwd_args(plan.handle, C.byref(a)))
        return out

    def _fwd(self, x, tape):
        y, saved = self.apply(x, tape, save=tape is not None)
        if tape is not None:
            tape.push((x, saved))
        return y

    def _bwd(self, gy, tape, need_gx=True):
        x, saved = tape.pop()
        return self.apply_bwd(x, gy, saved, tape, need_gx=need_gx)
