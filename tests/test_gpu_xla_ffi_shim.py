"""The XLA FFI handlers (libpdeqx_xla_ffi.so) driven through hand-built call frames on the GPU -- what XLA's runtime does
for a `jax.ffi.ffi_call` -- against the float64 oracle: operand / result / attribute decoding, the plan cache, the
workspace results and the stream hand-over of every handler a reference module binds (SpectralConv / ClassicSpectralBlock,
PhysicsConv, PhysicsConvTranspose, PointwiseLinearConv, GroupNorm, Adam)."""
import ctypes as C

import numpy as np
import pytest
import torch

import xla_ffi_frame as xf
from pdequinox_b200 import _lib
from oracle import conv as oconv, nn as onn, spectral as ospec
from conftest import rel_l2

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def B(t):
    """(device pointer, shape) of a tensor; None stands for an absent optional operand (zero-element buffer)."""
    return (None, (0,)) if t is None else (t.data_ptr(), tuple(t.shape))


@pytest.fixture(scope="module")
def ffi(cuda):
    lib, handlers = xf.load_shim()
    return lib, handlers, xf.Caller(stream=torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("precision,tol", [(0, 1e-5), (1, 2e-3)])
@pytest.mark.parametrize("D,spatial,modes", [(2, (32, 128), (8, 16)), (1, (64,), (12,))])
def test_spectral_block_handlers(ffi, precision, tol, D, spatial, modes):
    lib, h, call = ffi
    rng = np.random.default_rng(3)
    Bn, ci, co = 2, 32, 32
    x = rng.standard_normal((Bn, ci) + spatial)
    gy = rng.standard_normal((Bn, co) + spatial)
    wr = rng.standard_normal((2 ** (D - 1), co, ci) + modes) / (ci * co)
    wi = rng.standard_normal((2 ** (D - 1), co, ci) + modes) / (ci * co)
    bw = rng.uniform(-1, 1, (co, ci) + (1,) * D) / np.sqrt(ci)
    bb = rng.uniform(-1, 1, (co,) + (1,) * D)
    pre = ospec.spectral_conv_nd(x, wr, wi, modes) + oconv.pointwise_conv(x, bw, bb)
    ref_y = onn.gelu_tanh(pre)
    gpre = gy * onn.gelu_tanh_grad(pre)
    gx_s, gwr, gwi, _ = ospec.spectral_conv_vjp(x, wr, wi, modes, gpre)
    gx_b, gbw, gbb = oconv.pointwise_conv_vjp(x, bw, gpre)
    # result sizes as the JAX side obtains them at trace time
    d = _lib.SpectralDesc()
    d.ndim, d.batch, d.c_in, d.c_out = D, Bn, ci, co
    d.spatial, d.modes, d.precision = _lib.ivec(spatial, 1), _lib.ivec(modes, 1), precision
    sizes = (C.c_size_t * 3)()
    lib.pdeqx_ffi_spectral_sizes.argtypes = [C.POINTER(_lib.SpectralDesc), C.POINTER(C.c_size_t * 3)]
    assert lib.pdeqx_ffi_spectral_sizes(C.byref(d), C.byref(sizes)) == 0
    X, GY, WR, WI, BW, BB = dev(x), dev(gy), dev(wr), dev(wi), dev(bw), dev(bb)
    y = torch.empty((Bn, co) + spatial, device="cuda")
    xhat = torch.empty(sizes[0], device="cuda")
    z = torch.empty(sizes[1], device="cuda")
    ws = torch.empty(sizes[2] // 4 + 1, device="cuda")
    attrs = {"activation": 2, "precision": precision}
    call(h["pdeqx_ffi_spectral_block_fwd"], [B(X), B(WR), B(WI), B(BW), B(BB)], [B(y), B(xhat), B(z), B(ws)], attrs,
         stage=xf.STAGE_INITIALIZE)
    call(h["pdeqx_ffi_spectral_block_fwd"], [B(X), B(WR), B(WI), B(BW), B(BB)], [B(y), B(xhat), B(z), B(ws)], attrs)
    torch.cuda.synchronize()
    assert rel_l2(host(y), ref_y) <= tol
    gx, gwr_d, gwi_d = torch.empty_like(X), torch.empty_like(WR), torch.empty_like(WI)
    gbw_d, gbb_d, scratch = torch.empty_like(BW), torch.empty_like(BB), torch.empty_like(y)
    call(h["pdeqx_ffi_spectral_block_bwd"], [B(X), B(GY), B(WR), B(WI), B(BW), B(BB), B(xhat), B(z)],
         [B(gx), B(gwr_d), B(gwi_d), B(gbw_d), B(gbb_d), B(scratch), B(ws)], attrs)
    torch.cuda.synchronize()
    assert rel_l2(host(gx), gx_s + gx_b) <= tol
    assert rel_l2(host(gwr_d), gwr) <= tol and rel_l2(host(gwi_d), gwi) <= tol
    assert rel_l2(host(gbw_d), gbw) <= tol and rel_l2(host(gbb_d), gbb) <= tol
    # plain SpectralConv: no bypass operands, identity activation
    y2 = torch.empty_like(y)
    call(h["pdeqx_ffi_spectral_block_fwd"], [B(X), B(WR), B(WI), B(None), B(None)], [B(y2), B(xhat), B(z), B(ws)],
         {"activation": 0, "precision": precision})
    torch.cuda.synchronize()
    assert rel_l2(host(y2), ospec.spectral_conv_nd(x, wr, wi, modes)) <= tol


@pytest.mark.parametrize("mode,boundary", [("periodic", 1), ("dirichlet", 0), ("neumann", 2)])
@pytest.mark.parametrize("precision,C_,H,W,tol", [(0, 6, 12, 10, 1e-5), (1, 64, 16, 128, 2e-3)])
def test_conv_handlers(ffi, mode, boundary, precision, C_, H, W, tol):
    lib, h, call = ffi
    rng = np.random.default_rng(4)
    x = rng.standard_normal((2, C_, H, W))
    w = rng.standard_normal((C_, C_, 3, 3)) / np.sqrt(9 * C_)
    b = rng.standard_normal((C_, 1, 1))
    gy = rng.standard_normal((2, C_, H, W))
    attrs = {"activation": 1, "groups": 1, "boundary": boundary, "precision": precision, "stride": [1, 1], "dilation": [1, 1],
             "pad_lo": [1, 1], "pad_hi": [1, 1]}
    d = _lib.ConvDesc()
    d.ndim, d.batch, d.c_in, d.c_out, d.groups = 2, 2, C_, C_, 1
    d.spatial, d.kernel = _lib.ivec((H, W), 1), _lib.ivec((3, 3), 1)
    d.stride, d.dilation, d.pad_lo, d.pad_hi = _lib.ivec((1, 1), 1), _lib.ivec((1, 1), 1), _lib.ivec((1, 1), 0), _lib.ivec((1, 1), 0)
    d.boundary, d.precision = boundary, precision
    nws = _lib.lib().pdeqx_conv_workspace_bytes(d)
    assert (nws > 0) == (precision == 1)
    X, Wd, Bd, GY = dev(x), dev(w), dev(b), dev(gy)
    y = torch.empty_like(X)
    ws = torch.empty(nws // 4 + 1, device="cuda")
    call(h["pdeqx_ffi_conv_fwd"], [B(X), B(Wd), B(Bd), B(None)], [B(y), B(ws)], attrs)
    torch.cuda.synchronize()
    assert rel_l2(host(y), np.maximum(oconv.physics_conv(x, w, b, boundary_mode=mode), 0)) <= tol
    gx_r, gw_r, gb_r = oconv.physics_conv_vjp(x, w, gy, boundary_mode=mode)
    gx, gw, gb = torch.empty_like(X), torch.empty_like(Wd), torch.empty_like(Bd)
    call(h["pdeqx_ffi_conv_bwd"], [B(X), B(GY), B(Wd), B(None)], [B(gx), B(gw), B(gb), B(ws)], attrs)
    torch.cuda.synchronize()
    assert rel_l2(host(gx), gx_r) <= tol and rel_l2(host(gw), gw_r) <= tol and rel_l2(host(gb), gb_r) <= tol


def test_convT_pointwise_groupnorm_adam_handlers(ffi):
    lib, h, call = ffi
    rng = np.random.default_rng(5)
    # transposed conv: the UNet up-sampling block (k = 3, stride 2, output_padding 1; _linear_conv_up_block.py:16-36)
    x = rng.standard_normal((2, 4, 8))
    w = rng.standard_normal((3, 4, 3))
    b = rng.standard_normal((3, 1))
    attrs = {"groups": 1, "boundary": 0, "stride": [2], "dilation": [1], "pad_lo": [1], "pad_hi": [1], "output_padding": [1]}
    X, Wd, Bd = dev(x), dev(w), dev(b)
    y = torch.empty(2, 3, 16, device="cuda")
    call(h["pdeqx_ffi_convT_fwd"], [B(X), B(Wd), B(Bd)], [B(y)], attrs)
    torch.cuda.synchronize()
    ref = oconv.physics_conv_transpose(x, w, b, boundary_mode="dirichlet", stride=2, output_padding=1)
    assert rel_l2(host(y), ref) <= 1e-5
    gy = rng.standard_normal(ref.shape)
    gx, gw, gb = torch.empty_like(X), torch.empty_like(Wd), torch.empty_like(Bd)
    call(h["pdeqx_ffi_convT_bwd"], [B(X), B(dev(gy)), B(Wd)], [B(gx), B(gw), B(gb)], attrs)
    torch.cuda.synchronize()
    eps_dir = rng.standard_normal(x.shape)  # directional derivative of <convT(x), gy> along x against <gx, direction>
    num = np.sum((oconv.physics_conv_transpose(x + 1e-6 * eps_dir, w, b, boundary_mode="dirichlet", stride=2, output_padding=1)
                  - oconv.physics_conv_transpose(x - 1e-6 * eps_dir, w, b, boundary_mode="dirichlet", stride=2, output_padding=1)) * gy) / 2e-6
    assert abs(np.sum(host(gx) * eps_dir) - num) <= 1e-4 * abs(num)
    assert rel_l2(host(gb).ravel(), gy.sum(axis=(0, 2))) <= 1e-5

    # pointwise
    x = rng.standard_normal((3, 5, 6, 4))
    w = rng.standard_normal((2, 5, 1, 1))
    b = rng.standard_normal((2, 1, 1))
    X, Wd, Bd = dev(x), dev(w), dev(b)
    y = torch.empty(3, 2, 6, 4, device="cuda")
    call(h["pdeqx_ffi_pointwise_fwd"], [B(X), B(Wd), B(Bd), B(None)], [B(y)], {"activation": 0})
    torch.cuda.synchronize()
    assert rel_l2(host(y), oconv.pointwise_conv(x, w, b)) <= 1e-5
    gy = rng.standard_normal((3, 2, 6, 4))
    gx_r, gw_r, gb_r = oconv.pointwise_conv_vjp(x, w, gy)
    gx, gw, gb = torch.empty_like(X), torch.empty_like(Wd), torch.empty_like(Bd)
    call(h["pdeqx_ffi_pointwise_bwd"], [B(X), B(Wd), B(dev(gy))], [B(gx), B(gw), B(gb)], {})
    torch.cuda.synchronize()
    assert rel_l2(host(gx), gx_r) <= 1e-5 and rel_l2(host(gw), gw_r) <= 1e-5 and rel_l2(host(gb), gb_r) <= 1e-5

    # group norm
    x = rng.standard_normal((2, 6, 5, 4))
    wn, bn = rng.standard_normal(6), rng.standard_normal(6)
    gy = rng.standard_normal(x.shape)
    X, Wn, Bn_, GY = dev(x), dev(wn), dev(bn), dev(gy)
    y, stats = torch.empty_like(X), torch.empty(2 * 2 * 2, device="cuda")
    ws = torch.empty(_lib.lib().pdeqx_groupnorm_workspace_bytes(2, 6) // 4 + 1, device="cuda")
    call(h["pdeqx_ffi_groupnorm_fwd"], [B(X), B(Wn), B(Bn_)], [B(y), B(stats), B(ws)], {"groups": 2, "eps": 1e-5})
    torch.cuda.synchronize()
    assert rel_l2(host(y), onn.group_norm(x, wn, bn, 2, 2)) <= 1e-5
    gx_r, gw_r, gb_r = onn.group_norm_vjp(x, wn, 2, 2, gy)
    gx, gw, gb, scr = torch.empty_like(X), torch.empty_like(Wn), torch.empty_like(Bn_), torch.empty(2 * 2 * 6, device="cuda")
    call(h["pdeqx_ffi_groupnorm_bwd"], [B(X), B(Wn), B(stats), B(GY)], [B(gx), B(gw), B(gb), B(scr)], {"groups": 2})
    torch.cuda.synchronize()
    assert rel_l2(host(gx), gx_r) <= 1e-5 and rel_l2(host(gw), gw_r) <= 1e-5 and rel_l2(host(gb), gb_r) <= 1e-5

    # Adam, two updates in place (state on the device)
    p0 = rng.standard_normal(1000).astype(np.float32)
    g = rng.standard_normal(1000).astype(np.float32)
    P, G, M, V, S = dev(p0), dev(g), torch.zeros(1000, device="cuda"), torch.zeros(1000, device="cuda"), torch.zeros(4, device="cuda")
    ref_p, st = {"w": p0.copy()}, onn.adam_init({"w": p0})
    for _ in range(2):
        call(h["pdeqx_ffi_adam"], [B(P), B(G), B(M), B(V), B(S)], [B(P), B(M), B(V), B(S)],
             {"lr": 3e-4, "b1": 0.9, "b2": 0.999, "eps": 1e-8, "grad_scale": 1.0})
        ref_p, st = onn.adam_step(ref_p, {"w": g}, st)
    torch.cuda.synchronize()
    assert rel_l2(host(P), ref_p["w"]) <= 1e-6


@pytest.mark.parametrize("with_parts", [False, True])
def test_folded_norm_conv_handlers(ffi, with_parts):
    """pdeqx_ffi_norm_conv_fwd / _bwd: GroupNorm(1, C) -> periodic PhysicsConv -> relu as one op (the layer of
    pdequinox/blocks/_dilated_res_block.py:141-151), statistics from a pass of their own or from the producer's partial sums."""
    lib, h, call = ffi
    rng = np.random.default_rng(9)
    Bn, C_, H, W, dil = 2, 64, 16, 128, 2
    x = (0.5 + 1.5 * rng.standard_normal((Bn, C_, H, W))).astype(np.float32).astype(np.float64)
    gamma, beta = 1 + 0.3 * rng.standard_normal(C_), 0.3 * rng.standard_normal(C_)
    w = rng.standard_normal((C_, C_, 3, 3)) / np.sqrt(9 * C_)
    b = rng.standard_normal((C_, 1, 1))
    attrs = {"activation": 1, "eps": 1e-5, "groups": 1, "boundary": 1, "precision": 1, "stride": [1, 1], "dilation": [dil, dil],
             "pad_lo": [dil, dil], "pad_hi": [dil, dil]}
    d = _lib.ConvDesc()
    d.ndim, d.batch, d.c_in, d.c_out, d.groups = 2, Bn, C_, C_, 1
    d.spatial, d.kernel = _lib.ivec((H, W), 1), _lib.ivec((3, 3), 1)
    d.stride, d.dilation = _lib.ivec((1, 1), 1), _lib.ivec((dil, dil), 1)
    d.pad_lo, d.pad_hi = _lib.ivec((dil, dil), 0), _lib.ivec((dil, dil), 0)
    d.boundary, d.precision = 1, 1
    L = _lib.lib()
    assert L.pdeqx_conv_gn_fusable(d) == 1
    slots = L.pdeqx_conv_gn_slots(d)
    nws = max(L.pdeqx_conv_workspace_bytes(d), L.pdeqx_groupnorm_workspace_bytes(Bn, C_))
    X, G_, Be, Wd, Bd = dev(x), dev(gamma), dev(beta), dev(w), dev(b)
    y, stats = torch.empty_like(X), torch.empty(Bn, 2, device="cuda")
    parts, ws = torch.empty(Bn, slots, 2, device="cuda"), torch.empty(nws // 4 + 1, device="cuda")
    in_parts = None
    if with_parts:  # what a producing layer's epilogue would have written: here 4 slots per sample
        xs = x.reshape(Bn, 4, -1)
        in_parts = dev(np.stack([xs.sum(-1), (xs ** 2).sum(-1)], axis=-1))
    call(h["pdeqx_ffi_norm_conv_fwd"], [B(X), B(G_), B(Be), B(Wd), B(Bd), B(in_parts)], [B(y), B(stats), B(parts), B(ws)], attrs)
    torch.cuda.synchronize()
    hn = onn.group_norm(host(X).astype(np.float64), host(G_), host(Be), 1, 2)
    ref = np.maximum(oconv.physics_conv(hn, host(Wd), host(Bd), boundary_mode="periodic", dilation=dil), 0)
    assert rel_l2(host(y), ref) <= 2e-3
    p = host(parts).astype(np.float64).sum(1)
    np.testing.assert_allclose(p[:, 0], host(y).astype(np.float64).reshape(Bn, -1).sum(1), rtol=1e-4)
    # reverse pass: cotangent of the pre-activation in, everything else out
    ga = rng.standard_normal(ref.shape) * (host(y) > 0)
    ghn, gw_r, gb_r = oconv.physics_conv_vjp(hn, host(Wd).astype(np.float64), ga, boundary_mode="periodic", dilation=dil)
    gx_r, gg_r, gbe_r = onn.group_norm_vjp(host(X).astype(np.float64), host(G_).astype(np.float64), 1, 2, ghn)
    add = rng.standard_normal(x.shape)
    gx, gg, gbe, gw, gb = torch.empty_like(X), torch.empty_like(G_), torch.empty_like(Be), torch.empty_like(Wd), torch.empty_like(Bd)
    v, bparts, scr = torch.empty_like(X), torch.empty(Bn, slots, 2, device="cuda"), torch.empty(2 * Bn * C_, device="cuda")
    battrs = dict(attrs, relu_mask_x=1)
    GA, ADD = dev(ga), dev(add)  # (kept alive: a call frame holds raw device pointers)
    call(h["pdeqx_ffi_norm_conv_bwd"], [B(X), B(stats), B(G_), B(Be), B(Wd), B(GA), B(ADD)],
         [B(gx), B(gg), B(gbe), B(gw), B(gb), B(v), B(bparts), B(scr), B(ws)], battrs)
    torch.cuda.synchronize()
    assert rel_l2(host(gx), gx_r * (host(X) > 0) + add) <= 2e-3
    assert rel_l2(host(gw), gw_r) <= 2e-3 and rel_l2(host(gb).ravel(), gb_r.ravel()) <= 2e-3
    assert rel_l2(host(gg), gg_r) <= 2e-3 and rel_l2(host(gbe), gbe_r) <= 2e-3
