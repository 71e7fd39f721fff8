"""Internal consistency of the oracle (CPU): the dense truncated-DFT formulation the kernels implement
equals the reference's rfftn/einsum/irfftn formulation, and every analytic VJP equals finite differences."""
import numpy as np
import pytest

from oracle import conv as oconv, nn as onn, spectral as ospec
from conftest import rel_l2


@pytest.mark.parametrize("spatial,modes,ci,co", [((20,), (7,), 3, 5), ((8,), (5,), 2, 3), ((9,), (5,), 2, 2),
                                                 ((20, 18), (6, 5), 3, 5), ((12, 8), (3, 5), 2, 2),
                                                 ((6, 16), (4, 3), 2, 2), ((12, 10, 8), (4, 3, 3), 2, 3),
                                                 ((4, 6, 8), (3, 2, 5), 1, 2)])
def test_dense_equals_fft_formulation(spatial, modes, ci, co):
    rng = np.random.default_rng(0)
    D = len(spatial)
    x = rng.standard_normal((2, ci) + spatial)
    wr = rng.standard_normal((2 ** (D - 1), co, ci) + modes)
    wi = rng.standard_normal((2 ** (D - 1), co, ci) + modes)
    assert rel_l2(ospec.spectral_conv_dense(x, wr, wi, modes), ospec.spectral_conv_nd(x, wr, wi, modes)) < 1e-12


def _fd(f, arrs, idx, eps=1e-6):
    """Directional finite difference of scalar f w.r.t. arrs[idx] along a random direction."""
    rng = np.random.default_rng(123)
    d = rng.standard_normal(arrs[idx].shape)
    ap = [a.copy() for a in arrs]
    am = [a.copy() for a in arrs]
    ap[idx] = ap[idx] + eps * d
    am[idx] = am[idx] - eps * d
    return (f(*ap) - f(*am)) / (2 * eps), d


@pytest.mark.parametrize("spatial,modes", [((12,), (5,)), ((8,), (5,)), ((10, 8), (3, 4)), ((6, 16), (4, 3)),
                                           ((6, 5, 8), (2, 2, 3))])
def test_spectral_vjp_finite_differences(spatial, modes):
    rng = np.random.default_rng(1)
    D = len(spatial)
    x = rng.standard_normal((2, 3) + spatial)
    wr = rng.standard_normal((2 ** (D - 1), 4, 3) + modes)
    wi = rng.standard_normal((2 ** (D - 1), 4, 3) + modes)
    gy = rng.standard_normal((2, 4) + spatial)
    f = lambda x_, wr_, wi_: float(np.sum(ospec.spectral_conv_nd(x_, wr_, wi_, modes) * gy))
    gx, gwr, gwi, _ = ospec.spectral_conv_vjp(x, wr, wi, modes, gy)
    for i, g in enumerate([gx, gwr, gwi]):
        num, d = _fd(f, [x, wr, wi], i)
        assert num == pytest.approx(float(np.sum(g * d)), rel=1e-6, abs=1e-7)


@pytest.mark.parametrize("mode", ["periodic", "dirichlet", "neumann"])
@pytest.mark.parametrize("spatial,k,stride,dilation,groups", [((11,), 3, 1, 1, 1), ((12,), 4, 2, 2, 2),
                                                              ((7, 8), 3, 1, 3, 1), ((6, 6), (3, 2), (2, 1), 1, 1),
                                                              ((12, 12), 3, 1, 8, 1), ((5, 6, 4), 3, 1, 1, 1)])
def test_physics_conv_vjp_finite_differences(mode, spatial, k, stride, dilation, groups):
    rng = np.random.default_rng(2)
    D = len(spatial)
    ci, co = 4, 2 * groups
    kk = (k,) * D if isinstance(k, int) else k
    x = rng.standard_normal((2, ci) + spatial)
    w = rng.standard_normal((co, ci // groups) + kk)
    b = rng.standard_normal((co,) + (1,) * D)
    kw = dict(boundary_mode=mode, stride=stride, dilation=dilation, groups=groups)
    y = oconv.physics_conv(x, w, b, **kw)
    gy = rng.standard_normal(y.shape)
    f = lambda x_, w_, b_: float(np.sum(oconv.physics_conv(x_, w_, b_, **kw) * gy))
    gx, gw, gb = oconv.physics_conv_vjp(x, w, gy, **kw)
    for i, g in enumerate([gx, gw, gb]):
        num, d = _fd(f, [x, w, b], i)
        assert num == pytest.approx(float(np.sum(g * d)), rel=1e-6, abs=1e-7)


def test_group_norm_vjp_finite_differences():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 6, 5, 4))
    w, b = rng.standard_normal(6), rng.standard_normal(6)
    gy = rng.standard_normal(x.shape)
    f = lambda x_, w_, b_: float(np.sum(onn.group_norm(x_, w_, b_, 2, 2) * gy))
    gx, gw, gb = onn.group_norm_vjp(x, w, 2, 2, gy)
    for i, g in enumerate([gx, gw, gb]):
        num, d = _fd(f, [x, w, b], i)
        assert num == pytest.approx(float(np.sum(g * d)), rel=1e-5, abs=1e-7)


def test_fno_and_resnet_loss_gradients_finite_differences():
    rng = np.random.default_rng(4)
    p = {k: v.astype(np.float64) for k, v in onn.init_fno(rng, 2, 1, 1, hidden_channels=4, num_modes=3, num_blocks=2).items()}
    x, y = rng.standard_normal((2, 1, 8, 8)), rng.standard_normal((2, 1, 8, 8))
    loss, g = onn.fno_loss_and_grad(p, x, y, (3, 3), 2)
    assert loss == pytest.approx(onn.mse_loss(onn.fno_forward(p, x, (3, 3), 2), y))
    for key in p:
        d = rng.standard_normal(p[key].shape)
        pp, pm = dict(p), dict(p)
        pp[key], pm[key] = p[key] + 1e-6 * d, p[key] - 1e-6 * d
        num = (onn.mse_loss(onn.fno_forward(pp, x, (3, 3), 2), y) - onn.mse_loss(onn.fno_forward(pm, x, (3, 3), 2), y)) / 2e-6
        assert num == pytest.approx(float(np.sum(g[key].reshape(d.shape) * d)), rel=1e-5, abs=1e-9), key
    q = {k: v.astype(np.float64) for k, v in onn.init_classic_resnet(rng, 2, 1, 1, hidden_channels=3, num_blocks=2).items()}
    for mode in ("periodic", "dirichlet", "neumann"):
        loss, g = onn.classic_resnet_loss_and_grad(q, x, y, 2, mode)
        for key in q:
            d = rng.standard_normal(q[key].shape)
            qp, qm = dict(q), dict(q)
            qp[key], qm[key] = q[key] + 1e-6 * d, q[key] - 1e-6 * d
            num = (onn.mse_loss(onn.classic_resnet_forward(qp, x, 2, mode), y)
                   - onn.mse_loss(onn.classic_resnet_forward(qm, x, 2, mode), y)) / 2e-6
            assert num == pytest.approx(float(np.sum(g[key].reshape(d.shape) * d)), rel=2e-5, abs=1e-9), (mode, key)


@pytest.mark.parametrize("mode", ["periodic", "dirichlet", "neumann"])
def test_dilated_resnet_loss_gradients_finite_differences(mode):
    """The analytic reverse pass of the DilatedResNet oracle (GroupNorm -> dilated conv -> act per rate, + skip) against central
    differences of its own float64 loss along a random direction in every leaf; its forward pass is pinned by the golden
    vectors generated from the reference's source (arch_dilated_resnet_* cases)."""
    rng = np.random.default_rng(6)
    rates = (1, 2, 1)
    p = {k: v.astype(np.float64) for k, v in
         onn.init_dilated_resnet(rng, 2, 1, 1, hidden_channels=3, num_blocks=2, dilation_rates=rates).items()}
    for k in p:  # non-trivial affine parameters
        if "norm_layers" in k:
            p[k] = p[k] + 0.3 * rng.standard_normal(p[k].shape)
    x, y = rng.standard_normal((2, 1, 8, 6)), rng.standard_normal((2, 1, 8, 6))
    loss, g = onn.dilated_resnet_loss_and_grad(p, x, y, 2, mode, rates, activation="gelu")
    assert loss == pytest.approx(onn.mse_loss(onn.dilated_resnet_forward(p, x, 2, mode, rates, "gelu"), y))
    assert list(g) == list(p)
    # the golden-pinned evaluator (tests/golden_cases.py: arch_dilated_* cases) and this forward pass are the same function
    golden_path = onn.sequential_forward(p, x, lambda h, pre: onn.dilated_block_any(h, p, pre, mode, "gelu", 2, rates), 2)
    np.testing.assert_allclose(onn.dilated_resnet_forward(p, x, 2, mode, rates, "gelu"), golden_path, rtol=0, atol=1e-13)
    for key in p:
        d = rng.standard_normal(p[key].shape)
        pp, pm = dict(p), dict(p)
        pp[key], pm[key] = p[key] + 1e-6 * d, p[key] - 1e-6 * d
        num = (onn.mse_loss(onn.dilated_resnet_forward(pp, x, 2, mode, rates, "gelu"), y)
               - onn.mse_loss(onn.dilated_resnet_forward(pm, x, 2, mode, rates, "gelu"), y)) / 2e-6
        assert num == pytest.approx(float(np.sum(g[key].reshape(d.shape) * d)), rel=2e-5, abs=1e-9), key
