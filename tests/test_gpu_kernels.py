"""Parity of the CUDA kernels (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerance: rel-L2 <= 1e-5 in fp32 mode (BASELINE.json north_star); oracle evaluated in float64.
"""
import numpy as np
import pytest
import torch

import pdequinox_b200 as pdeqx
from pdequinox_b200 import _lib
from pdequinox_b200.train import ParameterArena
from oracle import conv as oconv, nn as onn, spectral as ospec
from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-5


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(autouse=True)
def _fp32(cuda):
    pdeqx.set_precision("fp32")


SPECTRAL_CASES = [
    # (D, B, ci, co, spatial, modes)
    (1, 3, 4, 5, (48,), (10,)),
    (1, 2, 32, 32, (64,), (12,)),
    (1, 2, 3, 3, (9,), (5,)),            # odd length, m = s//2+1
    (1, 2, 2, 3, (8,), (5,)),            # even length, Nyquist bin retained
    (2, 2, 3, 5, (20, 18), (6, 5)),
    (2, 2, 8, 8, (32, 32), (12, 12)),
    (2, 1, 2, 2, (12, 8), (3, 5)),       # Nyquist retained on the rfft axis
    (2, 1, 2, 2, (6, 16), (4, 3)),       # 2m > s: overlapping corner slices (later block overwrites)
    (3, 2, 3, 4, (12, 10, 8), (4, 3, 3)),
    (3, 1, 4, 4, (16, 16, 16), (6, 6, 6)),
]


@pytest.mark.parametrize("D,B,ci,co,spatial,modes", SPECTRAL_CASES)
def test_spectral_conv_forward_matches_oracle(D, B, ci, co, spatial, modes):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((B, ci) + spatial)
    wr = rng.standard_normal((2 ** (D - 1), co, ci) + modes) / (ci * co)
    wi = rng.standard_normal((2 ** (D - 1), co, ci) + modes) / (ci * co)
    ref = ospec.spectral_conv_nd(x, wr, wi, modes)
    conv = pdeqx.SpectralConv(D, co, ci, modes, key=0)  # allocation (G, in, out) == call-time (G, o, i)
    conv.weights_real, conv.weights_imag = dev(wr), dev(wi)
    y = host(pdeqx.vmap(conv)(dev(x)))
    assert y.shape == ref.shape
    assert rel_l2(y, ref) <= TOL
    y1 = host(conv(dev(x[0])))  # single-sample convention
    assert rel_l2(y1, ref[0]) <= TOL


@pytest.mark.parametrize("D,B,ci,co,spatial,modes", SPECTRAL_CASES)
@pytest.mark.parametrize("act", ["identity", "gelu"])
def test_spectral_block_fwd_bwd_matches_oracle(D, B, ci, co, spatial, modes, act):
    rng = np.random.default_rng(2)
    x = rng.standard_normal((B, ci) + spatial)
    gy = rng.standard_normal((B, co) + spatial)
    wr = rng.standard_normal((2 ** (D - 1), co, ci) + modes) / (ci * co)
    wi = rng.standard_normal((2 ** (D - 1), co, ci) + modes) / (ci * co)
    bw = rng.uniform(-1, 1, (co, ci) + (1,) * D) / np.sqrt(ci)
    bb = rng.uniform(-1, 1, (co,) + (1,) * D)
    f, fg = onn.ACTIVATIONS[act]
    pre = ospec.spectral_conv_nd(x, wr, wi, modes) + oconv.pointwise_conv(x, bw, bb)
    ref_y = f(pre)
    gpre = gy * fg(pre)
    gx_s, gwr, gwi, _ = ospec.spectral_conv_vjp(x, wr, wi, modes, gpre)
    gx_b, gbw, gbb = oconv.pointwise_conv_vjp(x, bw, gpre)

    blk = pdeqx.ClassicSpectralBlock(D, ci, co, activation=getattr(pdeqx.nn, act), num_modes=modes, key=0)
    blk.spectral_conv.weights_real, blk.spectral_conv.weights_imag = dev(wr), dev(wi)
    blk.by_pass_conv.weight, blk.by_pass_conv.bias = dev(bw), dev(bb)
    arena = ParameterArena(blk)
    tape = pdeqx._module.Tape()
    y = blk._fwd(dev(x), tape)
    assert rel_l2(host(y), ref_y) <= TOL
    gx = blk._bwd(dev(gy), tape, True)
    g = {k: host(v) for k, v in arena.named_grads().items()}
    assert rel_l2(host(gx), gx_s + gx_b) <= TOL
    assert rel_l2(g["spectral_conv.weights_real"], gwr) <= TOL
    assert rel_l2(g["spectral_conv.weights_imag"], gwi) <= TOL
    assert rel_l2(g["by_pass_conv.weight"], gbw) <= TOL
    assert rel_l2(g["by_pass_conv.bias"], gbb) <= TOL


@pytest.mark.parametrize("B,ci,co,n", [(2, 1, 64, 4096), (2, 64, 1, 4096), (3, 5, 7, 37), (2, 64, 64, 1000),
                                       (1, 32, 48, 333), (2, 3, 40, 64)])
def test_pointwise_fwd_bwd(B, ci, co, n):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((B, ci, n))
    w = rng.standard_normal((co, ci, 1))
    b = rng.standard_normal((co, 1))
    gy = rng.standard_normal((B, co, n))
    ref = oconv.pointwise_conv(x, w, b)
    gx_r, gw_r, gb_r = oconv.pointwise_conv_vjp(x, w, gy)
    m = pdeqx.PointwiseLinearConv(1, ci, co, key=0)
    m.weight, m.bias = dev(w), dev(b)
    arena = ParameterArena(m)
    tape = pdeqx._module.Tape()
    y = m._fwd(dev(x), tape)
    assert rel_l2(host(y), ref) <= TOL
    gx = m._bwd(dev(gy), tape, True)
    g = {k: host(v) for k, v in arena.named_grads().items()}
    assert rel_l2(host(gx), gx_r) <= TOL
    assert rel_l2(g["weight"], gw_r) <= TOL
    assert rel_l2(g["bias"], gb_r) <= TOL


def test_boundary_handling_known_answers():
    """reference tests/test_boundary_handling.py:35-52 through the CUDA path."""
    for mode, expect in [("periodic", [3] * 7), ("dirichlet", [2, 3, 3, 3, 3, 3, 2]), ("neumann", [3] * 7)]:
        conv = pdeqx.PhysicsConv(1, 1, 1, 3, use_bias=False, key=0, boundary_mode=mode)
        conv.weight = dev(np.ones((1, 1, 3)))
        y = host(conv(dev(np.ones((1, 7)))))
        assert y == pytest.approx(np.array([expect], dtype=np.float32))


CONV_CASES = [
    # (D, B, ci, co, spatial, k, stride, dilation, groups)
    (1, 2, 3, 4, (17,), 3, 1, 1, 1),
    (1, 2, 4, 6, (16,), 3, 2, 1, 2),
    (1, 1, 2, 2, (12,), 4, 1, 2, 1),      # even kernel: extra halo on the high side
    (1, 2, 2, 3, (9,), 5, 1, 3, 1),
    (2, 2, 3, 5, (10, 12), 3, 1, 1, 1),
    (2, 1, 4, 4, (16, 16), 3, 1, 4, 1),
    (2, 2, 2, 2, (9, 8), (3, 2), (2, 1), 1, 1),
    (2, 1, 16, 16, (12, 12), 3, 1, 8, 1),  # halo 8 > s/2: multiple wraps / reflections near the edges
    (3, 1, 2, 3, (6, 7, 8), 3, 1, 1, 1),
    (3, 1, 2, 2, (8, 8, 8), 3, 2, 1, 1),
]


@pytest.mark.parametrize("D,B,ci,co,spatial,k,stride,dilation,groups", CONV_CASES)
@pytest.mark.parametrize("mode", ["periodic", "dirichlet", "neumann"])
def test_physics_conv_fwd_bwd(D, B, ci, co, spatial, k, stride, dilation, groups, mode):
    rng = np.random.default_rng(4)
    conv = pdeqx.PhysicsConv(D, ci, co, k, stride=stride, dilation=dilation, groups=groups, key=0, boundary_mode=mode)
    w = rng.standard_normal(tuple(conv.weight.shape))
    b = rng.standard_normal(tuple(conv.bias.shape))
    x = rng.standard_normal((B, ci) + spatial)
    ref = oconv.physics_conv(x, w, b, boundary_mode=mode, stride=stride, dilation=dilation, groups=groups)
    gy = rng.standard_normal(ref.shape)
    gx_r, gw_r, gb_r = oconv.physics_conv_vjp(x, w, gy, boundary_mode=mode, stride=stride, dilation=dilation,
                                              groups=groups)
    conv.weight, conv.bias = dev(w), dev(b)
    arena = ParameterArena(conv)
    tape = pdeqx._module.Tape()
    y = conv._fwd(dev(x), tape)
    assert tuple(y.shape) == ref.shape
    assert rel_l2(host(y), ref) <= TOL
    gx = conv._bwd(dev(gy), tape, True)
    g = {kk: host(v) for kk, v in arena.named_grads().items()}
    assert rel_l2(host(gx), gx_r) <= TOL
    assert rel_l2(g["weight"], gw_r) <= TOL
    assert rel_l2(g["bias"], gb_r) <= TOL


@pytest.mark.parametrize("B,C,G,spatial", [(2, 8, 1, (16, 16)), (3, 6, 2, (37,)), (2, 4, 4, (5, 6, 7)), (2, 64, 1, (32, 32))])
def test_group_norm_fwd_bwd(B, C, G, spatial):
    rng = np.random.default_rng(5)
    D = len(spatial)
    x = rng.standard_normal((B, C) + spatial) * 2.0 + 0.5
    w = rng.standard_normal(C)
    b = rng.standard_normal(C)
    gy = rng.standard_normal(x.shape)
    ref = onn.group_norm(x, w, b, G, D)
    gx_r, gw_r, gb_r = onn.group_norm_vjp(x, w, G, D, gy)
    m = pdeqx.nn.GroupNorm(G, C)
    m.weight, m.bias = dev(w), dev(b)
    arena = ParameterArena(m)
    tape = pdeqx._module.Tape()
    y = m._fwd(dev(x), tape)
    assert rel_l2(host(y), ref) <= TOL
    gx = m._bwd(dev(gy), tape, True)
    g = {kk: host(v) for kk, v in arena.named_grads().items()}
    assert rel_l2(host(gx), gx_r) <= 2e-5
    assert rel_l2(g["weight"], gw_r) <= TOL
    assert rel_l2(g["bias"], gb_r) <= TOL


def test_errors_match_reference_behaviour():
    conv = pdeqx.PhysicsConv(2, 2, 2, 3, key=0, boundary_mode="periodic")
    with pytest.raises(ValueError):
        conv(torch.zeros(2, 8, device="cuda"))  # wrong rank (_conv.py:183-188)
    with pytest.raises(ValueError):
        pdeqx.SpectralConv(2, 2, 2, (4, 10), key=0)(torch.zeros(2, 16, 16, device="cuda"))  # modes > s//2+1 on last axis
