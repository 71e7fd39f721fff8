"""ModernResNet (pre-activation residual blocks) and ConvNet on the GPU (re-orderings of the same
PhysicsConv / GroupNorm / activation kernels): forward against the float64 oracle (itself pinned on golden vectors of the
reference's source, tests/test_golden_oracle.py), one train step with every parameter gradient checked through central
differences of the oracle's float64 loss along a random direction per leaf."""
import numpy as np
import pytest
import torch

import pdequinox_b200 as pdeqx
from pdequinox_b200.train import TrainStep
from oracle import nn as onn
from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _check_train_step(model, forward, B, cin, cout, spatial, seed):
    rng = np.random.default_rng(seed)
    step = TrainStep(model)
    # generic (non-zero) biases and norm affine parameters so that every gradient is exercised
    p32 = {k: (v.detach().cpu().numpy() + 0.05 * rng.standard_normal(tuple(v.shape))).astype(np.float32)
           for k, v in step.arena.named_params().items()}
    step.arena.load(p32)
    p = {k: v.astype(np.float64) for k, v in p32.items()}
    x = rng.standard_normal((B, cin) + spatial)
    y = rng.standard_normal((B, cout) + spatial)

    def loss_of(q):
        return float(np.mean((forward(q, x) - y) ** 2))

    pred = pdeqx.vmap(model)(torch.from_numpy(x).float().cuda()).cpu().numpy()
    assert rel_l2(pred, forward(p, x)) <= 1e-5
    loss = step.loss_and_grad(x, y).item()
    assert abs(loss - loss_of(p)) <= 1e-5 * abs(loss_of(p))
    g = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in step.arena.named_grads().items()}
    assert all(np.isfinite(v).all() for v in g.values())
    checked = 0
    for k in p:
        v = rng.standard_normal(p[k].shape)
        fds = []
        for eps in (1e-6, 1e-7):
            hi, lo = dict(p), dict(p)
            hi[k] = p[k] + eps * v
            lo[k] = p[k] - eps * v
            fds.append((loss_of(hi) - loss_of(lo)) / (2 * eps))
        if abs(fds[0] - fds[1]) > 2e-5 * max(abs(fds[0]), 1e-4):
            continue  # a relu kink inside the difference stencil: the quotient is not a derivative here
        checked += 1
        an = float(np.sum(g[k] * v))
        assert abs(an - fds[0]) <= 2e-4 * max(abs(fds[0]), 1e-4), (k, an, fds)
    assert checked >= 0.8 * len(p)
    before = step.arena.params.clone()
    step.apply_gradients()
    assert torch.isfinite(step.arena.params).all() and not torch.equal(before, step.arena.params)


@pytest.mark.parametrize("D,spatial,mode,use_norm", [
    (1, (24,), "periodic", True), (1, (16,), "dirichlet", False), (2, (10, 12), "neumann", True),
    (2, (8, 8), "periodic", False), (2, (12, 8), "dirichlet", True)])
def test_modern_resnet_train_step_matches_oracle(cuda, D, spatial, mode, use_norm):
    pdeqx.set_precision("fp32")
    model = pdeqx.ModernResNet(D, 1, 1, hidden_channels=5, num_blocks=2, use_norm=use_norm, boundary_mode=mode, key=4)

    def forward(q, x):
        return onn.sequential_forward(q, x, lambda h, pre: onn.modern_block_any(h, q, pre, mode, "relu", D), 2)

    _check_train_step(model, forward, 3, 1, 1, spatial, 31)


@pytest.mark.parametrize("use_norm", [False, True])
def test_modern_res_block_with_channel_change(cuda, use_norm):
    """in != out: the bypass is a bias-free 1x1 conv of the (optionally) normalised input (_modern_res_block.py:107-127)."""
    pdeqx.set_precision("fp32")
    mode = "periodic"
    model = pdeqx.blocks.ModernResBlock(2, 3, 5, use_norm=use_norm, boundary_mode=mode, key=2)
    assert ("bypass_norm.weight" in pdeqx.tree_paths(model)) == use_norm
    _check_train_step(model, lambda q, x: onn.modern_block_any(x, q, "", mode, "relu", 2), 2, 3, 5, (9, 10), 32)


@pytest.mark.parametrize("D,spatial,mode,depth", [
    (1, (20,), "periodic", 3), (2, (9, 11), "dirichlet", 2), (2, (8, 10), "neumann", 0), (3, (5, 6, 7), "periodic", 1)])
def test_conv_net_train_step_matches_oracle(cuda, D, spatial, mode, depth):
    pdeqx.set_precision("fp32")
    model = pdeqx.ConvNet(D, 2, 3, hidden_channels=4, depth=depth, boundary_mode=mode, key=6)
    assert len(model.layers) == depth + 1
    _check_train_step(model, lambda q, x: onn.conv_net_forward(q, x, mode), 2, 2, 3, spatial, 33)


def test_conv_net_gelu_hidden_and_relu_final(cuda):
    """An activation whose derivative needs the pre-activation (gelu) takes the unfused path; a final activation is applied."""
    pdeqx.set_precision("fp32")
    model = pdeqx.ConvNet(1, 1, 1, hidden_channels=3, depth=2, activation=pdeqx.nn.gelu, final_activation=pdeqx.nn.relu,
                          boundary_mode="periodic", key=1)
    _check_train_step(model, lambda q, x: onn.conv_net_forward(q, x, "periodic", "gelu", "relu"), 2, 1, 1, (16,), 34)


@pytest.mark.parametrize("D,spatial,mode,use_norm,levels,blocks", [
    (1, (32,), "dirichlet", True, 2, 2), (1, (16,), "periodic", False, 1, 1), (2, (8, 12), "neumann", True, 2, 1)])
def test_modern_unet_train_step_matches_oracle(cuda, D, spatial, mode, use_norm, levels, blocks):
    """Hierarchical with pre-activation ResNet arcs (pdequinox/arch/_modern_u_net.py:78-98): forward and every gradient."""
    pdeqx.set_precision("fp32")
    model = pdeqx.ModernUNet(D, 1, 1, hidden_channels=4, num_levels=levels, num_blocks=blocks, use_norm=use_norm,
                             boundary_mode=mode, key=8)
    _check_train_step(model, lambda q, x: onn.modern_unet_forward(q, x, mode, levels, blocks), 2, 1, 1, spatial, 35)
