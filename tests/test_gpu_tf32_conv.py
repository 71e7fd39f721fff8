"""TF32 mode of PhysicsConv: the tcgen05 implicit-GEMM kernel (2-D, k = 3, W = 128, C = 64 / 32) against the float64
oracle, tolerance rel-L2 <= 2e-3, for every boundary mode and the dilations of DilatedResNet."""
import numpy as np
import pytest
import torch

import pdequinox_b200 as pdeqx
from pdequinox_b200 import _lib
from pdequinox_b200.train import ParameterArena
from oracle import conv as oconv, nn as onn
from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 2e-3


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(autouse=True)
def _tf32(cuda):
    pdeqx.set_precision("tf32")
    yield
    pdeqx.set_precision("fp32")


@pytest.mark.parametrize("mode", ["periodic", "dirichlet", "neumann"])
@pytest.mark.parametrize("C,H,d,B", [(64, 32, 1, 2), (64, 128, 1, 1), (64, 24, 2, 2), (64, 40, 4, 1), (64, 32, 8, 2), (64, 37, 3, 1), (64, 70, 1, 1),
                                     (32, 16, 1, 3)])
def test_tf32_conv_fwd_bwd(mode, C, H, d, B):
    rng = np.random.default_rng(21)
    conv = pdeqx.PhysicsConv(2, C, C, 3, dilation=d, key=0, boundary_mode=mode)
    w = rng.standard_normal(tuple(conv.weight.shape)) / np.sqrt(9 * C)
    b = rng.standard_normal(tuple(conv.bias.shape))
    x = rng.standard_normal((B, C, H, 128))
    res = rng.standard_normal((B, C, H, 128))
    conv.weight, conv.bias = dev(w), dev(b)
    arena = ParameterArena(conv)
    desc, _ = conv._desc(dev(x))
    passes = _lib.lib().pdeqx_conv_tcgen05_passes(desc)
    if C == 64:
        assert passes == 7, f"tcgen05 passes {passes}"
    else:
        assert passes == 0  # C = 32 runs the fp32 path (also under set_precision("tf32"))
    # plain
    ref = oconv.physics_conv(x, w, b, boundary_mode=mode, dilation=d)
    pair0 = _lib.lib().pdeqx_conv_pair_launches()
    y = conv.apply(dev(x), None)
    # the CTA-pair kernel (cta_group::2) takes every C = 64 shape whose row chains split into equal segments
    assert (_lib.lib().pdeqx_conv_pair_launches() > pair0) == (C == 64 and H % d == 0)
    torch.cuda.synchronize()
    assert rel_l2(host(y), ref) <= TOL
    # fused residual + relu epilogue
    y2 = conv.apply(dev(x), None, residual=dev(res), act=_lib.ACT_RELU)
    torch.cuda.synchronize()
    assert rel_l2(host(y2), np.maximum(ref + res, 0)) <= TOL
    # reverse pass
    gy = rng.standard_normal(ref.shape)
    gx_r, gw_r, gb_r = oconv.physics_conv_vjp(x, w, gy, boundary_mode=mode, dilation=d)
    gx = conv.apply_bwd(dev(x), dev(gy), None, True, add=dev(res))
    torch.cuda.synchronize()
    g = {kk: host(v) for kk, v in arena.named_grads().items()}
    assert rel_l2(host(gx), gx_r + res) <= TOL
    assert rel_l2(g["weight"], gw_r) <= TOL
    assert rel_l2(g["bias"], gb_r) <= TOL


def test_tf32_resnet_train_step(cuda):
    from pdequinox_b200.train import TrainStep
    rng = np.random.default_rng(22)
    model = pdeqx.ClassicResNet(2, 1, 1, hidden_channels=64, num_blocks=2, boundary_mode="periodic", key=3)
    step = TrainStep(model)
    p = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in step.arena.named_params().items()}
    x = rng.standard_normal((2, 1, 32, 128))
    y = rng.standard_normal((2, 1, 32, 128))
    pred = host(pdeqx.vmap(model)(dev(x)))
    assert rel_l2(pred, onn.classic_resnet_forward(p, x, 2, "periodic")) <= TOL
    loss = step.loss_and_grad(x, y).item()
    # relu' is discontinuous: the oracle's reverse pass uses the relu masks of the forward pass under test
    from pdequinox_b200._module import Tape
    tape = Tape()
    model._fwd(dev(x), tape)
    masks = {}
    for i, (_, h, p1, yy, p2) in enumerate(r for r in tape.stack if isinstance(r, tuple) and len(r) == 5):
        masks[(i, 0)] = host(h if p1 is None else p1) > 0
        masks[(i, 1)] = host(yy if p2 is None else p2) > 0
    loss_ref, g_ref = onn.classic_resnet_loss_and_grad(p, x, y, 2, "periodic", relu_masks=masks)
    assert abs(loss - loss_ref) <= TOL * abs(loss_ref)
    g = {k: host(v) for k, v in step.arena.named_grads().items()}
    for k in g_ref:
        assert rel_l2(g[k], g_ref[k]) <= TOL, k


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
@pytest.mark.parametrize("mode,d,H", [("periodic", 1, 32), ("dirichlet", 2, 24), ("neumann", 1, 32), ("periodic", 3, 37)])
def test_conv_data_gradient_applies_the_relu_mask(precision, mode, d, H):
    """pdeqx_conv_bwd_data(relu_mask): (conv^T(gy) + add) * [mask > 0] in ONE pass equals the unmasked data gradient followed by
    the mask -- on the CTA-pair kernel, the one-CTA kernel (H % d != 0), the two-pass neumann path and the fp32 kernels."""
    pdeqx.set_precision(precision)
    rng = np.random.default_rng(31)
    conv = pdeqx.PhysicsConv(2, 64, 64, 3, dilation=d, key=0, boundary_mode=mode)
    ParameterArena(conv)
    x = dev(rng.standard_normal((2, 64, H, 128)))
    gy = dev(rng.standard_normal((2, 64, H, 128)))
    add = dev(rng.standard_normal((2, 64, H, 128)))
    mask = dev(rng.standard_normal((2, 64, H, 128)))
    plain = host(conv.apply_bwd(x, gy, None, True, add=add))
    masked = host(conv.apply_bwd(x, gy, None, True, add=add, relu_mask=mask))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(masked, plain * (host(mask) > 0))


def test_groupnorm_reverse_pass_applies_the_relu_mask(cuda):
    from pdequinox_b200 import nn as pnn
    from pdequinox_b200._module import Tape
    rng = np.random.default_rng(32)
    norm = pnn.GroupNorm(groups=2, channels=8)
    ParameterArena(norm)
    x = dev(rng.standard_normal((3, 8, 10, 12)))
    gy = dev(rng.standard_normal((3, 8, 10, 12)))
    out = {}
    for masked in (False, True):
        tape = Tape()
        norm._fwd(x, tape)
        out[masked] = host(norm._bwd(gy, tape, relu_mask=x if masked else None)).copy()
    np.testing.assert_array_equal(out[True], out[False] * (host(x) > 0))
