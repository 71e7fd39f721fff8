"""ClassicUNet path on the GPU: PhysicsConvTranspose and its two gradients, the Hierarchical
reverse pass (skip concat / split, stride-2 down-sampling) and one train step, against the float64 oracle.

The oracle has no hand-written reverse pass for this path; gradients are pinned through exact linearity (Jacobian
columns of the linear operator, evaluated with the oracle) and through central finite differences of the oracle's
float64 loss along a random direction in every parameter leaf."""
import numpy as np
import pytest
import torch

import pdequinox_b200 as pdeqx
from pdequinox_b200 import _lib
from pdequinox_b200.train import ParameterArena, TrainStep
from oracle import conv as oconv, nn as onn
from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["periodic", "dirichlet", "neumann"])
@pytest.mark.parametrize("D,ci,co,spatial,k,stride,op,dil", [
    (1, 2, 3, (8,), 3, 2, 1, 1), (1, 2, 2, (8,), 4, 2, 1, 1), (1, 2, 2, (9,), 3, 2, 1, 2), (2, 2, 2, (5, 6), 3, 2, 1, 1),
    (1, 3, 2, (7,), 3, 1, 0, 1), (1, 2, 2, (6,), 5, 3, 2, 1)])
def test_conv_transpose_forward_and_gradients(cuda, mode, D, ci, co, spatial, k, stride, op, dil):
    pdeqx.set_precision("fp32")
    rng = np.random.default_rng(11)
    m = pdeqx.PhysicsConvTranspose(D, ci, co, k, stride=stride, output_padding=op, dilation=dil, boundary_mode=mode, key=3)
    arena = ParameterArena(m)
    w = m.weight.detach().cpu().numpy().astype(np.float64)
    b = m.bias.detach().cpu().numpy().astype(np.float64)
    B = 2
    x = rng.standard_normal((B, ci) + spatial)

    def ref(xx, ww, bb):
        return oconv.physics_conv_transpose(xx, ww, bb, boundary_mode=mode, stride=stride, output_padding=op, dilation=dil)

    y_ref = ref(x, w, b)
    tape = pdeqx._module.Tape()
    xd = torch.from_numpy(x).float().cuda()
    y = m._fwd(xd, tape)
    assert tuple(y.shape) == y_ref.shape
    assert rel_l2(y.cpu().numpy(), y_ref) <= 1e-5
    gy = rng.standard_normal(y_ref.shape)
    gx = m._bwd(torch.from_numpy(gy).float().cuda(), tape, True).cpu().numpy()
    # exact adjoints through the Jacobian columns of the (linear) oracle
    gx_ref = np.zeros_like(x)
    for idx in np.ndindex(*x.shape):
        e = np.zeros_like(x)
        e[idx] = 1.0
        gx_ref[idx] = np.sum(ref(e, w, None) * gy)
    gw_ref = np.zeros_like(w)
    for idx in np.ndindex(*w.shape):
        e = np.zeros_like(w)
        e[idx] = 1.0
        gw_ref[idx] = np.sum(ref(x, e, None) * gy)
    gb_ref = gy.sum(axis=tuple(a for a in range(gy.ndim) if a != 1))
    g = arena.named_grads()
    assert rel_l2(gx, gx_ref) <= 1e-5
    assert rel_l2(g["weight"].cpu().numpy(), gw_ref) <= 1e-5
    assert rel_l2(g["bias"].cpu().numpy().ravel(), gb_ref) <= 1e-5


def test_reference_transposed_conv_known_answer(cuda):
    """x = 1..8, w = [10, 1, 100], k = 3, stride 2, output_padding 1 -> last element 80 / 180 / 780."""
    for mode, last in (("dirichlet", 80.0), ("periodic", 180.0), ("neumann", 780.0)):
        m = pdeqx.PhysicsConvTranspose(1, 1, 1, 3, stride=2, output_padding=1, use_bias=False, boundary_mode=mode, key=0)
        m.weight = torch.tensor([[[10.0, 1.0, 100.0]]], device="cuda")
        y = m(torch.arange(1.0, 9.0, device="cuda")[None]).cpu().numpy()[0]
        assert y.shape == (16,)
        assert y[-1] == last
        assert np.array_equal(y[0::2], np.arange(1.0, 9.0))  # even outputs see the centre tap only


def test_unet_rejects_indivisible_grids(cuda):
    m = pdeqx.ClassicUNet(1, 1, 1, hidden_channels=4, num_levels=2, key=0)
    with pytest.raises(ValueError, match="Spatial dim issue"):  # _hierarchical.py:252-255
        m(torch.zeros(1, 30, device="cuda"))


def _oracle_loss(p, x, y, mode, levels):
    pred = onn.unet_forward(p, x, mode, levels)
    return float(np.mean((pred - y) ** 2))


@pytest.mark.parametrize("D,spatial,mode,use_norm,levels", [
    (1, (32,), "dirichlet", True, 2), (1, (32,), "periodic", False, 2), (1, (16,), "neumann", True, 1),
    (2, (16, 8), "periodic", True, 2), (2, (8, 8), "dirichlet", False, 1), (2, (8, 12), "neumann", False, 2)])
def test_unet_train_step_matches_oracle(cuda, D, spatial, mode, use_norm, levels):
    pdeqx.set_precision("fp32")
    rng = np.random.default_rng(21)
    model = pdeqx.ClassicUNet(D, 1, 1, hidden_channels=4, num_levels=levels, use_norm=use_norm, boundary_mode=mode, key=5)
    step = TrainStep(model)
    # generic (non-zero) biases and norm affine parameters so that every gradient is exercised
    p32 = {k: (v.detach().cpu().numpy() + 0.05 * rng.standard_normal(tuple(v.shape))).astype(np.float32)
           for k, v in step.arena.named_params().items()}
    step.arena.load(p32)
    p = {k: v.astype(np.float64) for k, v in p32.items()}
    B = 3
    x = rng.standard_normal((B, 1) + spatial)
    y = rng.standard_normal((B, 1) + spatial)
    pred = pdeqx.vmap(model)(torch.from_numpy(x).float().cuda()).cpu().numpy()
    assert rel_l2(pred, onn.unet_forward(p, x, mode, levels)) <= 1e-5
    loss = step.loss_and_grad(x, y).item()
    loss_ref = _oracle_loss(p, x, y, mode, levels)
    assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
    g = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in step.arena.named_grads().items()}
    assert all(np.isfinite(v).all() for v in g.values())
    # per-leaf directional derivatives <grad_k, v_k> against central differences of the float64 oracle loss (one leaf
    # at a time: a small perturbation norm keeps the relu kinks out of the difference quotient)
    checked = 0
    for k in p:
        v = rng.standard_normal(p[k].shape)
        fds = []
        for eps in (1e-6, 1e-7):
            hi, lo = dict(p), dict(p)
            hi[k] = p[k] + eps * v
            lo[k] = p[k] - eps * v
            fds.append((_oracle_loss(hi, x, y, mode, levels) - _oracle_loss(lo, x, y, mode, levels)) / (2 * eps))
        if abs(fds[0] - fds[1]) > 2e-5 * max(abs(fds[0]), 1e-4):
            continue  # a relu kink inside the difference stencil: the quotient is not a derivative here
        checked += 1
        an = float(np.sum(g[k] * v))
        assert abs(an - fds[0]) <= 2e-4 * max(abs(fds[0]), 1e-4), (k, an, fds)
    assert checked >= 0.8 * len(p)
    before = step.arena.params.clone()
    step.apply_gradients()
    assert torch.isfinite(step.arena.params).all() and not torch.equal(before, step.arena.params)


def test_unet_parameter_counts(cuda):
    import golden_cases as gc
    for D in (1, 2, 3):  # tests/test_count_parameters.py of the reference (defaults: hidden 16, 4 levels, norm)
        assert pdeqx.count_parameters(pdeqx.ClassicUNet(D, 1, 1, key=0)) == gc.PARAMETER_COUNTS[f"ClassicUNet_{D}d"]
