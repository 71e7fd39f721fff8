"""Parity at the REAL geometry of every BASELINE.json config that bench.py / scripts/bench_configs.py time (small batch, full
grid / width / depth): unit counts per CTA, 3-D row grouping, the four-block TF32 error chain and the dilated row taps are
those of the benchmarked shapes. Oracle = float64 NumPy restatement (oracle/nn.py), tolerances = BASELINE.json north_star:
rel-L2 <= 2e-3 in the stated TF32 mode, <= 1e-5 in fp32 mode.

Reference: /root/reference/pdequinox/arch/_classic_fno.py:10-84, _dilated_res_net.py:10-75, _classic_res_net.py:10-77."""
import numpy as np
import pytest
import torch

import pdequinox_b200 as pdeqx
from pdequinox_b200.train import TrainStep
from oracle import nn as onn
from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = {"tf32": 2e-3, "fp32": 1e-5}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def oracle_params(step):
    return {k: host(v).astype(np.float64) for k, v in step.arena.named_params().items()}


def check_grads(step, g_ref, tol, label):
    g = {k: host(v) for k, v in step.arena.named_grads().items()}
    assert list(g) == list(g_ref)
    worst = max((rel_l2(g[k], g_ref[k]), k) for k in g_ref)
    assert worst[0] <= tol, f"{label}: worst per-leaf gradient rel-L2 {worst[0]:.2e} at {worst[1]} (tolerance {tol})"
    return worst


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_fno2d_configs3_geometry(cuda, precision):
    """BASELINE.json configs[3]: ClassicFNO(2, 1, 1, hidden 64, 16 modes, 4 blocks) on 256 x 256 (two samples)."""
    pdeqx.set_precision(precision)
    try:
        rng = np.random.default_rng(41)
        model = pdeqx.ClassicFNO(2, 1, 1, hidden_channels=64, num_modes=16, num_blocks=4, key=0)
        step = TrainStep(model)
        with torch.no_grad():
            step.arena.named_params()["lifting.bias"].add_(0.2)
            step.arena.named_params()["projection.bias"].add_(-0.1)
        p = oracle_params(step)
        x, y = rng.standard_normal((2, 1, 256, 256)), rng.standard_normal((2, 1, 256, 256))
        pred = host(pdeqx.vmap(model)(dev(x)))
        assert rel_l2(pred, onn.fno_forward(p, x, (16, 16), 4)) <= TOL[precision]
        loss_ref, g_ref = onn.fno_loss_and_grad(p, x, y, (16, 16), 4)
        loss = step.loss_and_grad(x, y).item()
        assert abs(loss - loss_ref) <= TOL[precision] * abs(loss_ref)
        check_grads(step, g_ref, TOL[precision], f"fno2d {precision}")
    finally:
        pdeqx.set_precision("fp32")


def test_fno3d_configs4_geometry(cuda):
    """BASELINE.json configs[4]: ClassicFNO(3, 1, 1, hidden 32, 12 modes, 4 blocks) on 64^3 (one sample), TF32 mode
    (padded mode space 12 -> 16 / 24 -> 32, two image rows per slab unit)."""
    pdeqx.set_precision("tf32")
    try:
        rng = np.random.default_rng(42)
        model = pdeqx.ClassicFNO(3, 1, 1, hidden_channels=32, num_modes=12, num_blocks=4, key=1)
        step = TrainStep(model)
        p = oracle_params(step)
        x, y = rng.standard_normal((1, 1, 64, 64, 64)), rng.standard_normal((1, 1, 64, 64, 64))
        pred = host(pdeqx.vmap(model)(dev(x)))
        assert rel_l2(pred, onn.fno_forward(p, x, (12, 12, 12), 4)) <= TOL["tf32"]
        loss_ref, g_ref = onn.fno_loss_and_grad(p, x, y, (12, 12, 12), 4)
        loss = step.loss_and_grad(x, y).item()
        assert abs(loss - loss_ref) <= TOL["tf32"] * abs(loss_ref)
        check_grads(step, g_ref, TOL["tf32"], "fno3d tf32")
    finally:
        pdeqx.set_precision("fp32")


def relu_masks_of_forward(model, x):
    """relu masks of the forward pass under test. relu' is discontinuous: where |pre| is below the precision's error
    (a fraction ~ eps of 10^6 pixels per layer) the mask may legitimately differ from the float64 oracle's, and each flipped
    pixel changes the gradient by its full cotangent, so the oracle's reverse pass is evaluated with these masks."""
    from pdequinox_b200._module import Tape
    tape = Tape()
    model._fwd(dev(x), tape)
    masks, i = {}, 0
    for rec in tape.stack:
        if isinstance(rec, list):  # DilatedResBlock: [(normalised input, activation output, pre-activation or None)] per layer
            for j, (_, out, pre) in enumerate(rec):
                masks[(i, j)] = host(out if pre is None else pre) > 0
            i += 1
        elif isinstance(rec, tuple) and len(rec) == 5:  # ClassicResBlock: (x, h, p1, y, p2)
            _, h, p1, yy, p2 = rec
            masks[(i, 0)] = host(h if p1 is None else p1) > 0
            masks[(i, 1)] = host(yy if p2 is None else p2) > 0
            i += 1
    return masks


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_dilated_resnet_configs2_geometry(cuda, precision):
    """BASELINE.json configs[2]: DilatedResNet(2, 1, 1, hidden 64, 2 blocks x dilations (1,2,4,8,4,2,1), GroupNorm) on
    128 x 128 periodic (one sample): prediction, loss and every parameter gradient (oracle reverse pass evaluated with the
    relu masks of the pass under test, see relu_masks_of_forward)."""
    pdeqx.set_precision(precision)
    try:
        rng = np.random.default_rng(43)
        model = pdeqx.DilatedResNet(2, 1, 1, hidden_channels=64, boundary_mode="periodic", key=2)
        step = TrainStep(model)
        with torch.no_grad():  # non-trivial GroupNorm affine parameters
            for k, v in step.arena.named_params().items():
                if "norm_layers" in k:
                    v.add_(dev(0.2 * rng.standard_normal(tuple(v.shape))))
        p = oracle_params(step)
        x, y = rng.standard_normal((1, 1, 128, 128)), rng.standard_normal((1, 1, 128, 128))
        pred = host(pdeqx.vmap(model)(dev(x)))
        assert rel_l2(pred, onn.dilated_resnet_forward(p, x, 2, "periodic")) <= TOL[precision]
        loss = step.loss_and_grad(x, y).item()
        masks = relu_masks_of_forward(model, x)
        assert len(masks) == 14
        loss_ref, g_ref = onn.dilated_resnet_loss_and_grad(p, x, y, 2, "periodic", relu_masks=masks)
        assert abs(loss - loss_ref) <= TOL[precision] * abs(loss_ref)
        # Gradient bar of this 14-layer chain: the north star's 2e-3 is stated for OUTPUTS (checked above). A parameter gradient
        # of the first layer has crossed 14 forward and 14 reverse tcgen05 stages, each of which rounds both operands to tf32
        # once (relative RMS 2^-11 / sqrt(3) per operand, ~4e-4 per product, unbiased since the producers round to nearest):
        # independent stage errors add in quadrature, ~6e-4 * sqrt(28) = 3.2e-3 ("TF32 error chain"). fp32 mode: 1e-5.
        check_grads(step, g_ref, 3.2e-3 if precision == "tf32" else TOL[precision], f"dilated {precision}")
    finally:
        pdeqx.set_precision("fp32")


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_classic_resnet_configs2_geometry(cuda, precision):
    """BASELINE.json configs[2]: ClassicResNet(2, 1, 1, hidden 64, 6 blocks, relu, no norm) on 128 x 128 periodic (one
    sample): prediction, loss and every parameter gradient (relu masks taken from the pass under test)."""
    pdeqx.set_precision(precision)
    try:
        rng = np.random.default_rng(44)
        model = pdeqx.ClassicResNet(2, 1, 1, hidden_channels=64, boundary_mode="periodic", key=3)
        step = TrainStep(model)
        p = oracle_params(step)
        x, y = rng.standard_normal((1, 1, 128, 128)), rng.standard_normal((1, 1, 128, 128))
        pred = host(pdeqx.vmap(model)(dev(x)))
        assert rel_l2(pred, onn.classic_resnet_forward(p, x, 6, "periodic")) <= TOL[precision]
        loss = step.loss_and_grad(x, y).item()
        masks = relu_masks_of_forward(model, x)
        assert len(masks) == 12
        loss_ref, g_ref = onn.classic_resnet_loss_and_grad(p, x, y, 6, "periodic", relu_masks=masks)
        assert abs(loss - loss_ref) <= TOL[precision] * abs(loss_ref)
        check_grads(step, g_ref, TOL[precision], f"resnet {precision}")
    finally:
        pdeqx.set_precision("fp32")
