"""GroupNorm (groups = 1) folded into the periodic tcgen05 convolution behind it (pdeqx_conv_gn_*, pdeqx_groupnorm_bwd_fused,
pdeqx_add_parts): the norm -> conv -> act chain of /root/reference/pdequinox/blocks/_dilated_res_block.py:141-151 without the
normalised tensor, against the float64 oracle (oracle/nn.py group_norm / group_norm_vjp, oracle/conv.py physics_conv /
physics_conv_vjp). Tolerance: rel-L2 <= 2e-3 (TF32 mode, BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

import pdequinox_b200 as pdeqx
from pdequinox_b200 import _lib, nn as pnn
from pdequinox_b200._module import Tape
from pdequinox_b200.train import ParameterArena
from oracle import conv as oconv, nn as onn
from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 2e-3


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def host(t):
    return t.detach().cpu().numpy().astype(np.float64)


def tf32(a):
    """round to nearest tf32 (10 explicit mantissa bits): what every producer of a tensor-core operand stores"""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)


@pytest.fixture(autouse=True)
def _tf32(cuda):
    pdeqx.set_precision("tf32")
    yield
    pdeqx.set_precision("fp32")


class _Chain(pdeqx._module.Module):
    _fields = ("norm", "conv")

    def __init__(self, d):
        self.num_spatial_dims = 2
        self.norm = pnn.GroupNorm(1, 64)
        self.conv = pdeqx.PhysicsConv(2, 64, 64, 3, dilation=d, key=0, boundary_mode="periodic")


@pytest.mark.parametrize("H,d,B", [(32, 1, 2), (128, 1, 1), (24, 2, 2), (40, 4, 3), (32, 8, 2)])
@pytest.mark.parametrize("act", ["relu", "identity"])
def test_folded_norm_conv_act_chain(H, d, B, act):
    rng = np.random.default_rng(100 + H + d)
    m = _Chain(d)
    arena = ParameterArena(m)
    C = 64
    gamma, beta = 1.0 + 0.3 * rng.standard_normal(C), 0.3 * rng.standard_normal(C)
    w = rng.standard_normal((C, C, 3, 3)) / np.sqrt(9 * C)
    b = rng.standard_normal((C, 1, 1))
    with torch.no_grad():
        m.norm.weight.copy_(dev(gamma)); m.norm.bias.copy_(dev(beta)); m.conv.weight.copy_(dev(w)); m.conv.bias.copy_(dev(b))
    gamma, beta, w, b = host(m.norm.weight), host(m.norm.bias), host(m.conv.weight), host(m.conv.bias)
    # a mean and a scale well away from (0, 1), and values a tensor-core stage may read without rounding
    x = tf32(0.7 + 1.9 * rng.standard_normal((B, C, H, 128)))
    xd = dev(x)
    assert m.conv.gn_fusable(xd)
    code = {"relu": _lib.ACT_RELU, "identity": _lib.ACT_IDENTITY}[act]
    f = (lambda t: np.maximum(t, 0)) if act == "relu" else (lambda t: t)

    # ---- forward: statistics in a pass of their own (first layer of a block), conv with the norm folded in ----
    tape = Tape()
    stats = m.norm.stats_of(xd, tape)
    y, parts, slots = m.conv.apply_gn(xd, m.norm, stats, tape, act=code)
    torch.cuda.synchronize()
    hn = onn.group_norm(x, gamma, beta, 1, 2)
    pre = oconv.physics_conv(hn, w, b, boundary_mode="periodic", dilation=d)
    ref = f(pre)
    assert rel_l2(host(y), ref) <= TOL
    st = host(stats).reshape(B, 2)
    np.testing.assert_allclose(st[:, 0], x.reshape(B, -1).mean(1), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st[:, 1], 1 / np.sqrt(x.reshape(B, -1).var(1) + 1e-5), rtol=1e-5)
    # the epilogue's partial sums give the statistics of y (what the NEXT norm needs) without reading y again
    stats_y = m.norm.stats_from_parts(parts, slots, B, y[0].numel(), tape, tag="stats_y")
    sy, yh = host(stats_y).reshape(B, 2), host(y).reshape(B, -1)
    np.testing.assert_allclose(sy[:, 0], yh.mean(1), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(sy[:, 1], 1 / np.sqrt(yh.var(1) + 1e-5), rtol=1e-4)

    # ---- reverse: filter / bias gradient against the virtual normalised input, data gradient + the norm's reverse pass ----
    gy = rng.standard_normal(ref.shape)
    mask = (host(y) > 0) if act == "relu" else np.ones_like(ref, dtype=bool)
    ga = tf32(gy * mask)
    ghn_r, gw_r, gb_r = oconv.physics_conv_vjp(hn, w, ga, boundary_mode="periodic", dilation=d)
    gx_r, ggam_r, gbet_r = onn.group_norm_vjp(x, gamma, 1, 2, ghn_r)
    add = rng.standard_normal(x.shape)
    v, bparts, bslots = m.conv.apply_bwd_gn(xd, dev(ga), m.norm, stats, tape)
    assert rel_l2(host(v), ghn_r) <= TOL
    for mask_h in (False, True):
        gx = m.norm.bwd_fused(xd, v, stats, bparts, bslots, tape, mask_h=mask_h, add=dev(add))
        torch.cuda.synchronize()
        want = gx_r * ((x > 0) if mask_h else 1.0) + add
        assert rel_l2(host(gx), want) <= TOL, mask_h
    g = {k: host(t) for k, t in arena.named_grads().items()}
    assert rel_l2(g["conv.weight"], gw_r) <= TOL
    assert rel_l2(g["conv.bias"].ravel(), gb_r.ravel()) <= TOL
    assert rel_l2(g["norm.weight"], ggam_r) <= TOL
    assert rel_l2(g["norm.bias"], gbet_r) <= TOL


def test_add_parts_matches_add_and_numpy(cuda):
    rng = np.random.default_rng(7)
    a, b = rng.standard_normal((3, 64, 16, 128)), 2.0 + rng.standard_normal((3, 64, 16, 128))
    tape = Tape()
    owner = object()
    out, parts, slots = pnn.add_parts(dev(a), dev(b), tape, owner, "y")
    ref = tf32((host(dev(a)) + host(dev(b))).astype(np.float32))  # stored rounded: a tcgen05 stage reads it next
    np.testing.assert_allclose(host(out), ref, rtol=1e-6, atol=1e-6)
    p = host(parts).reshape(3, slots, 2).sum(1)
    np.testing.assert_allclose(p[:, 0], ref.reshape(3, -1).sum(1), rtol=1e-5)
    np.testing.assert_allclose(p[:, 1], (ref.reshape(3, -1) ** 2).sum(1), rtol=1e-5)


@pytest.mark.parametrize("act", ["identity", "relu"])
def test_dilated_block_takes_the_folded_path_and_matches_the_unfolded_one(cuda, act):
    """The same DilatedResBlock run folded (TF32, periodic, groups = 1) and unfolded (fast paths off -> fp32 kernels, separate
    GroupNorm passes). With the identity activation every output and gradient agrees to the TF32 tolerance; with relu the
    masks of the two runs differ at a few near-zero pre-activations and every flipped pixel moves a gradient by its full
    cotangent (error ~ sqrt(fraction flipped)), so that case is a wiring check -- the oracle comparison with the masks of
    the pass under test is tests/test_gpu_full_shapes.py::test_dilated_resnet_configs2_geometry."""
    rng = np.random.default_rng(11)
    blk = pdeqx.blocks.DilatedResBlock(2, 64, 64, boundary_mode="periodic", key=5, activation=getattr(pnn, act))
    arena = ParameterArena(blk)
    with torch.no_grad():
        for k, t in arena.named_params().items():
            if "norm_layers" in k:
                t.add_(dev(0.2 * rng.standard_normal(tuple(t.shape))))
    x, gy = dev(rng.standard_normal((2, 64, 32, 128))), dev(rng.standard_normal((2, 64, 32, 128)))
    res = {}
    for folded in (True, False):
        _lib.lib().pdeqx_set_fast_paths(1 if folded else 0)
        try:
            tape = Tape()
            y = blk._fwd(x, tape)
            assert (tape.stack[-2]["dilated_folded"] is not None) == folded
            gx = blk._bwd(gy, tape)
            torch.cuda.synchronize()
            res[folded] = (host(y), host(gx), {k: host(t) for k, t in arena.named_grads().items()})
        finally:
            _lib.lib().pdeqx_set_fast_paths(1)
    assert rel_l2(res[True][0], res[False][0]) <= TOL
    gtol = 2 * TOL if act == "identity" else 0.1  # (identity: 7 forward + 7 reverse tf32 stages in a row)
    assert rel_l2(res[True][1], res[False][1]) <= gtol
    for k in res[True][2]:
        assert rel_l2(res[True][2][k], res[False][2][k]) <= gtol, k
