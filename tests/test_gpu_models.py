"""Whole-model parity on the GPU: ClassicFNO / ClassicResNet forward, loss, gradients and one Adam
step against the CPU oracle with identical weights (rel-L2 <= 1e-5, fp32 mode)."""
import numpy as np
import pytest
import torch

import pdequinox_b200 as pdeqx
from pdequinox_b200.train import TrainStep
from oracle import nn as onn
from conftest import rel_l2

pytestmark = pytest.mark.gpu


def params_to_oracle(step):
    return {k: v.detach().cpu().numpy().astype(np.float64) for k, v in step.arena.named_params().items()}


@pytest.mark.parametrize("D,spatial,hidden,modes,B", [(1, (64,), 32, 12, 8), (2, (32, 32), 16, 8, 4),
                                                      (3, (16, 16, 16), 8, 6, 2)])
def test_fno_train_step_matches_oracle(cuda, D, spatial, hidden, modes, B):
    pdeqx.set_precision("fp32")
    rng = np.random.default_rng(7)
    model = pdeqx.ClassicFNO(D, 1, 1, hidden_channels=hidden, num_modes=modes, num_blocks=2, key=1)
    step = TrainStep(model, lr=3e-4)
    p = params_to_oracle(step)
    x = rng.standard_normal((B, 1) + spatial)
    y = rng.standard_normal((B, 1) + spatial)
    pred = pdeqx.vmap(model)(torch.from_numpy(x).float().cuda()).cpu().numpy()
    ref_pred = onn.fno_forward(p, x, (modes,) * D, 2)
    assert rel_l2(pred, ref_pred) <= 1e-5
    loss_ref, g_ref = onn.fno_loss_and_grad(p, x, y, (modes,) * D, 2)
    loss = step.loss_and_grad(x, y).item()
    assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
    g = {k: v.detach().cpu().numpy() for k, v in step.arena.named_grads().items()}
    for k in g_ref:
        assert rel_l2(g[k], g_ref[k]) <= 2e-5, k
    # Adam update
    state = onn.adam_init({k: v.astype(np.float32) for k, v in p.items()})
    p32 = {k: v.astype(np.float32) for k, v in p.items()}
    new_p, _ = onn.adam_step(p32, {k: g_ref[k].astype(np.float32).reshape(p32[k].shape) for k in p32}, state)
    step.apply_gradients()
    after = params_to_oracle(step)
    for k in new_p:
        assert np.max(np.abs(after[k] - new_p[k])) <= 3e-4 * 2e-2 + 1e-7, k  # lr * tolerance on mu/sqrt(nu)


@pytest.mark.parametrize("mode", ["periodic", "dirichlet", "neumann"])
def test_resnet_train_step_matches_oracle(cuda, mode):
    pdeqx.set_precision("fp32")
    rng = np.random.default_rng(8)
    model = pdeqx.ClassicResNet(2, 1, 1, hidden_channels=8, num_blocks=2, boundary_mode=mode, key=2)
    step = TrainStep(model)
    p = params_to_oracle(step)
    x = rng.standard_normal((3, 1, 12, 10))
    y = rng.standard_normal((3, 1, 12, 10))
    pred = pdeqx.vmap(model)(torch.from_numpy(x).float().cuda()).cpu().numpy()
    assert rel_l2(pred, onn.classic_resnet_forward(p, x, 2, mode)) <= 1e-5
    loss_ref, g_ref = onn.classic_resnet_loss_and_grad(p, x, y, 2, mode)
    loss = step.loss_and_grad(x, y).item()
    assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
    g = {k: v.detach().cpu().numpy() for k, v in step.arena.named_grads().items()}
    for k in g_ref:
        assert rel_l2(g[k], g_ref[k]) <= 2e-5, k


def test_dilated_resnet_forward_matches_oracle(cuda):
    pdeqx.set_precision("fp32")
    rng = np.random.default_rng(9)
    model = pdeqx.DilatedResNet(2, 1, 1, hidden_channels=8, num_blocks=1, boundary_mode="periodic", key=3)
    step = TrainStep(model)
    p = params_to_oracle(step)
    x = rng.standard_normal((2, 1, 32, 32))
    h = onn.oconv.pointwise_conv(x, p["lifting.weight"], p["lifting.bias"])
    h = onn.dilated_res_block(h, p, "blocks.0.", "periodic")
    ref = onn.oconv.pointwise_conv(h, p["projection.weight"], p["projection.bias"])
    pred = pdeqx.vmap(model)(torch.from_numpy(x).float().cuda()).cpu().numpy()
    assert rel_l2(pred, ref) <= 1e-5
    # the reverse pass runs and produces finite gradients of the right shapes
    step.loss_and_grad(x, rng.standard_normal(x.shape))
    assert torch.isfinite(step.arena.grads).all()


def test_graph_replayed_train_steps_equal_eager_steps(cuda):
    """GraphedTrainStep replays the captured launch sequence: after the same number of updates on the same batches the
    parameters must equal those of the eagerly launched steps (same kernels, same order; only the atomically reduced
    bypass gradients may differ in summation order)."""
    from pdequinox_b200.train import GraphedTrainStep
    pdeqx.set_precision("fp32")
    rng = np.random.default_rng(31)
    xs = [torch.from_numpy(rng.standard_normal((4, 1, 32, 32))).float().cuda() for _ in range(3)]
    ys = [torch.from_numpy(rng.standard_normal((4, 1, 32, 32))).float().cuda() for _ in range(3)]

    def make():
        return TrainStep(pdeqx.ClassicFNO(2, 1, 1, hidden_channels=8, num_modes=4, num_blocks=2, key=9), lr=1e-2)

    eager = make()
    losses_e = [eager(x, y).item() for x, y in zip(xs, ys)]
    other = make()
    snapshot = other.arena.params.clone()
    graphed = GraphedTrainStep(other, xs[0], ys[0], warmup=2)
    # construction (warm-up steps + capture) must leave the optimisation where it was
    assert torch.equal(other.arena.params, snapshot) and other.count == 0
    assert not other.mu.any() and not other.nu.any() and not other.adam_state.any()
    losses_g = [graphed(x, y).item() for x, y in zip(xs, ys)]
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) <= 1e-5 * abs(a)
    assert rel_l2(other.arena.params.cpu().numpy(), eager.arena.params.cpu().numpy()) <= 1e-5
