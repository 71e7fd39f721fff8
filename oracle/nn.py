"""Oracle: blocks, architectures, loss, VJPs and Adam (TEST INFRASTRUCTURE ONLY).

Follows
  * /root/reference/pdequinox/blocks/_classic_spectral_block.py:72-75
  * /root/reference/pdequinox/blocks/_classic_res_block.py:112-121
  * /root/reference/pdequinox/blocks/_dilated_res_block.py:141-151
  * /root/reference/pdequinox/_sequential.py:111-116 (lifting -> blocks -> projection)
  * /root/reference/README.md:74-86 (MSE loss over all elements, Adam(3e-4))
Third-party pieces restated from their documented definitions ("parity
unpinned": jax/equinox/optax cannot be installed here): jax.nn.gelu
(approximate=True, tanh form), jax.nn.relu, eqx.nn.GroupNorm (biased variance,
eps=1e-5, per-channel affine), optax.adam (b1=.9, b2=.999, eps=1e-8,
eps_root=0, bias-corrected).

A model's parameters are a flat `dict[str, np.ndarray]` whose insertion order is
the reference pytree's leaf order.
"""
from __future__ import annotations

import numpy as np

from . import conv as oconv
from . import spectral as ospec

_SQRT_2_OVER_PI = 0.7978845608028654
_GELU_C = 0.044715


def gelu_tanh(x):
    u = _SQRT_2_OVER_PI * (x + _GELU_C * x * x * x)
    return (0.5 * x * (1.0 + np.tanh(u))).astype(x.dtype)


def gelu_tanh_grad(x):
    u = _SQRT_2_OVER_PI * (x + _GELU_C * x * x * x)
    t = np.tanh(u)
    du = _SQRT_2_OVER_PI * (1.0 + 3.0 * _GELU_C * x * x)
    return (0.5 * (1.0 + t) + 0.5 * x * (1.0 - t * t) * du).astype(x.dtype)


def relu(x):
    return np.maximum(x, 0).astype(x.dtype)


def relu_grad(x):
    return (x > 0).astype(x.dtype)


ACTIVATIONS = {"gelu": (gelu_tanh, gelu_tanh_grad), "relu": (relu, relu_grad),
               "identity": (lambda x: x, lambda x: np.ones_like(x))}


def group_norm(x, weight, bias, groups, D, eps=1e-5):
    """eqx.nn.GroupNorm on (..., C, *S): per-sample, per-group mean / biased var."""
    lead = x.ndim - D - 1
    C = x.shape[lead]
    xs = x.reshape(x.shape[:lead] + (groups, -1))
    mean = xs.mean(axis=-1, keepdims=True)
    var = xs.var(axis=-1, keepdims=True)
    out = ((xs - mean) / np.sqrt(var + eps)).reshape(x.shape)
    shp = (C,) + (1,) * D
    if weight is not None:
        out = out * weight.reshape(shp) + bias.reshape(shp)
    return out.astype(x.dtype)


# ---------------------------------------------------------------------------
# ClassicSpectralBlock / ClassicFNO
# ---------------------------------------------------------------------------

def spectral_block(x, p, prefix, modes, activation="gelu", dense=False):
    """activation(spectral_conv(x) + by_pass_conv(x)) (_classic_spectral_block.py:72-75)."""
    f = ospec.spectral_conv_dense if dense else ospec.spectral_conv_nd
    s = f(x, p[prefix + "spectral_conv.weights_real"], p[prefix + "spectral_conv.weights_imag"], modes)
    b = oconv.pointwise_conv(x, p[prefix + "by_pass_conv.weight"], p.get(prefix + "by_pass_conv.bias"))
    return ACTIVATIONS[activation][0](s + b)


def init_fno(rng, num_spatial_dims, in_channels, out_channels, *, hidden_channels=32, num_modes=12, num_blocks=4,
             dtype=np.float32):
    """Leaf order of ClassicFNO (Sequential: lifting, blocks[*], projection; App. C).
    Distributions follow the reference constructors; bit streams of jax.random are not reproducible."""
    D = num_spatial_dims
    p = {}
    w, b = oconv.init_conv_weights(rng, D, in_channels, hidden_channels, 1, dtype=dtype)
    p["lifting.weight"], p["lifting.bias"] = w, b
    for i in range(num_blocks):
        wr, wi = ospec.init_spectral_weights(rng, D, hidden_channels, hidden_channels, num_modes, dtype)
        p[f"blocks.{i}.spectral_conv.weights_real"] = wr
        p[f"blocks.{i}.spectral_conv.weights_imag"] = wi
        w, b = oconv.init_conv_weights(rng, D, hidden_channels, hidden_channels, 1, dtype=dtype)
        p[f"blocks.{i}.by_pass_conv.weight"], p[f"blocks.{i}.by_pass_conv.bias"] = w, b
    w, b = oconv.init_conv_weights(rng, D, hidden_channels, out_channels, 1, dtype=dtype)
    p["projection.weight"], p["projection.bias"] = w, b
    return p


def fno_forward(p, x, modes, num_blocks, activation="gelu", dense=False, return_acts=False):
    """Sequential.__call__ (_sequential.py:111-116) of a ClassicFNO, batched over the leading axis."""
    acts = []
    h = oconv.pointwise_conv(x, p["lifting.weight"], p.get("lifting.bias"))
    for i in range(num_blocks):
        acts.append(h)
        h = spectral_block(h, p, f"blocks.{i}.", modes, activation, dense)
    acts.append(h)
    y = oconv.pointwise_conv(h, p["projection.weight"], p.get("projection.bias"))
    return (y, acts) if return_acts else y


def mse_loss(pred, target):
    """jnp.mean((pred - target)**2) over ALL elements (README.md:74-76)."""
    d = pred.astype(np.float64) - target.astype(np.float64)
    return float(np.mean(d * d))


def fno_loss_and_grad(p, x, y, modes, num_blocks, activation="gelu"):
    """Loss and gradient dict (same keys/order as p) of mse(fno(x), y); analytic reverse mode."""
    act, act_grad = ACTIVATIONS[activation]
    D = len(modes)
    grads = {}
    hs = []          # block inputs
    pres = []        # block pre-activations
    h = oconv.pointwise_conv(x, p["lifting.weight"], p.get("lifting.bias"))
    for i in range(num_blocks):
        pre = (ospec.spectral_conv_dense(h, p[f"blocks.{i}.spectral_conv.weights_real"],
                                          p[f"blocks.{i}.spectral_conv.weights_imag"], modes)
               + oconv.pointwise_conv(h, p[f"blocks.{i}.by_pass_conv.weight"], p.get(f"blocks.{i}.by_pass_conv.bias")))
        hs.append(h)
        pres.append(pre)
        h = act(pre)
    pred = oconv.pointwise_conv(h, p["projection.weight"], p.get("projection.bias"))
    diff = pred - y
    loss = float(np.mean(diff.astype(np.float64) ** 2))
    g = (2.0 / diff.size) * diff
    g, gw, gb = oconv.pointwise_conv_vjp(h, p["projection.weight"], g, "projection.bias" in p)
    grads["projection.weight"], grads["projection.bias"] = gw, gb
    for i in reversed(range(num_blocks)):
        gpre = g * act_grad(pres[i])
        gx_s, gwr, gwi, _ = ospec.spectral_conv_vjp(hs[i], p[f"blocks.{i}.spectral_conv.weights_real"],
                                                    p[f"blocks.{i}.spectral_conv.weights_imag"], modes, gpre)
        gx_b, gw, gb = oconv.pointwise_conv_vjp(hs[i], p[f"blocks.{i}.by_pass_conv.weight"], gpre,
                                                f"blocks.{i}.by_pass_conv.bias" in p)
        grads[f"blocks.{i}.spectral_conv.weights_real"] = gwr
        grads[f"blocks.{i}.spectral_conv.weights_imag"] = gwi
        grads[f"blocks.{i}.by_pass_conv.weight"] = gw
        grads[f"blocks.{i}.by_pass_conv.bias"] = gb
        g = gx_s + gx_b
    _, gw, gb = oconv.pointwise_conv_vjp(x, p["lifting.weight"], g, "lifting.bias" in p)
    grads["lifting.weight"], grads["lifting.bias"] = gw, gb
    return loss, {k: grads[k] for k in p if grads.get(k) is not None}


# ---------------------------------------------------------------------------
# ClassicResBlock / DilatedResBlock / ResNets (boundary-aware conv path)
# ---------------------------------------------------------------------------

def classic_res_block(x, p, prefix, boundary_mode, activation="relu", use_norm=False, D=2):
    """_classic_res_block.py:112-121 with in==out (bypass = Identity)."""
    act = ACTIVATIONS[activation][0]
    h = oconv.physics_conv(x, p[prefix + "conv_1.weight"], p.get(prefix + "conv_1.bias"), boundary_mode=boundary_mode)
    if use_norm:
        h = group_norm(h, p[prefix + "norm_1.weight"], p[prefix + "norm_1.bias"], 1, D)
    h = act(h)
    h = oconv.physics_conv(h, p[prefix + "conv_2.weight"], p.get(prefix + "conv_2.bias"), boundary_mode=boundary_mode)
    if use_norm:
        h = group_norm(h, p[prefix + "norm_2.weight"], p[prefix + "norm_2.bias"], 1, D)
    if prefix + "bypass_conv.weight" in p:
        skip = oconv.pointwise_conv(x, p[prefix + "bypass_conv.weight"], None)
        if use_norm:
            skip = group_norm(skip, p[prefix + "bypass_norm.weight"], p[prefix + "bypass_norm.bias"], 1, D)
    else:
        skip = x
    return act(h + skip)


def dilated_res_block(x, p, prefix, boundary_mode, dilation_rates=(1, 2, 4, 8, 4, 2, 1), activation="relu",
                      use_norm=True, D=2):
    """_dilated_res_block.py:141-151 with in==out: per rate norm -> conv -> act; then + skip."""
    act = ACTIVATIONS[activation][0]
    h = x
    for j, d in enumerate(dilation_rates):
        if use_norm:
            h = group_norm(h, p[prefix + f"norm_layers.{j}.weight"], p[prefix + f"norm_layers.{j}.bias"], 1, D)
        h = oconv.physics_conv(h, p[prefix + f"conv_layers.{j}.weight"], p.get(prefix + f"conv_layers.{j}.bias"),
                               boundary_mode=boundary_mode, dilation=d)
        h = act(h)
    return h + x


def init_classic_resnet(rng, D, in_channels, out_channels, *, hidden_channels=32, num_blocks=6, kernel_size=3,
                        use_norm=False, dtype=np.float32):
    p = {}
    p["lifting.weight"], p["lifting.bias"] = oconv.init_conv_weights(rng, D, in_channels, hidden_channels, 1, dtype=dtype)
    for i in range(num_blocks):
        for c in ("conv_1", "conv_2"):
            w, b = oconv.init_conv_weights(rng, D, hidden_channels, hidden_channels, kernel_size, dtype=dtype)
            p[f"blocks.{i}.{c}.weight"], p[f"blocks.{i}.{c}.bias"] = w, b
            if use_norm:
                n = c.replace("conv", "norm")
                p[f"blocks.{i}.{n}.weight"] = np.ones(hidden_channels, dtype)
                p[f"blocks.{i}.{n}.bias"] = np.zeros(hidden_channels, dtype)
    p["projection.weight"], p["projection.bias"] = oconv.init_conv_weights(rng, D, hidden_channels, out_channels, 1, dtype=dtype)
    return p


def classic_resnet_forward(p, x, num_blocks, boundary_mode, activation="relu", use_norm=False):
    D = p["lifting.weight"].ndim - 2
    h = oconv.pointwise_conv(x, p["lifting.weight"], p.get("lifting.bias"))
    for i in range(num_blocks):
        h = classic_res_block(h, p, f"blocks.{i}.", boundary_mode, activation, use_norm, D)
    return oconv.pointwise_conv(h, p["projection.weight"], p.get("projection.bias"))


# ---------------------------------------------------------------------------
# optax.adam
# ---------------------------------------------------------------------------

def adam_init(p):
    return {"count": 0, "mu": {k: np.zeros_like(v) for k, v in p.items()},
            "nu": {k: np.zeros_like(v) for k, v in p.items()}}


def adam_step(p, g, state, lr=3e-4, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adam: mu,nu EMA -> bias correction -> p -= lr * mu_hat / (sqrt(nu_hat) + eps)."""
    t = state["count"] + 1
    new_p, mu, nu = {}, {}, {}
    c1 = 1.0 - b1 ** t
    c2 = 1.0 - b2 ** t
    for k in p:
        gk = g[k].astype(p[k].dtype)
        mu[k] = (b1 * state["mu"][k] + (1.0 - b1) * gk).astype(p[k].dtype)
        nu[k] = (b2 * state["nu"][k] + (1.0 - b2) * gk * gk).astype(p[k].dtype)
        upd = (mu[k] / c1) / (np.sqrt(nu[k] / c2) + eps)
        new_p[k] = (p[k] - lr * upd).astype(p[k].dtype)
    return new_p, {"count": t, "mu": mu, "nu": nu}


def count_parameters(p):
    return int(sum(v.size for v in p.values() if v is not None))


def poisson_1d_dirichlet(rng, num_points=64, num_samples=1000, domain_extent=5.0, dtype=np.float32):
    """NumPy restatement of pdequinox/_sample_data.py:6-71 (same distribution; jax.random stream not reproducible)."""
    grid = np.linspace(0, domain_extent, num_points + 2)[1:-1]
    dx = grid[1] - grid[0]
    A = (np.diag(np.ones(num_points - 1), -1) - 2 * np.diag(np.ones(num_points), 0)
         + np.diag(np.ones(num_points - 1), 1)) / dx ** 2
    lo = rng.uniform(0.2 * domain_extent, 0.4 * domain_extent, num_samples)
    hi = rng.uniform(0.6 * domain_extent, 0.8 * domain_extent, num_samples)
    f = ((grid[None, :] >= lo[:, None]) & (grid[None, :] <= hi[:, None])).astype(np.float64)
    u = np.linalg.solve(A, -f.T).T
    return f[:, None, :].astype(dtype), u[:, None, :].astype(dtype)


# ---------------------------------------------------------------------------
# VJPs for the ResNet path (the reference relies on JAX autodiff)
# ---------------------------------------------------------------------------

def group_norm_vjp(x, weight, groups, D, gy, eps=1e-5):
    """Returns (gx, gweight, gbias) of group_norm."""
    lead = x.ndim - D - 1
    C = x.shape[lead]
    xs = x.reshape(x.shape[:lead] + (groups, -1)).astype(np.float64)
    mean = xs.mean(axis=-1, keepdims=True)
    var = xs.var(axis=-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = ((xs - mean) * rstd).reshape(x.shape)
    shp = (C,) + (1,) * D
    w = np.ones(C) if weight is None else weight.astype(np.float64)
    g = gy.astype(np.float64) * w.reshape(shp)
    gs = g.reshape(xs.shape)
    xh = xhat.reshape(xs.shape)
    gx = rstd * (gs - gs.mean(axis=-1, keepdims=True) - xh * (gs * xh).mean(axis=-1, keepdims=True))
    red = tuple(a for a in range(x.ndim) if a != lead)
    gw = (gy.astype(np.float64) * xhat).sum(axis=red)
    gb = gy.astype(np.float64).sum(axis=red)
    return gx.reshape(x.shape).astype(x.dtype), gw.astype(x.dtype), gb.astype(x.dtype)


def classic_resnet_loss_and_grad(p, x, y, num_blocks, boundary_mode, activation="relu", relu_masks=None):
    """mse(classic_resnet(x), y) and its gradient (use_norm=False, in==out channels).
    `relu_masks` (optional, {(block, 0 | 1): bool array}): evaluate act' with the given masks (see dilated_resnet_loss_and_grad)."""
    act, act_grad = ACTIVATIONS[activation]
    grads = {}
    h = oconv.pointwise_conv(x, p["lifting.weight"], p.get("lifting.bias"))
    saved = []
    for i in range(num_blocks):
        pre1 = oconv.physics_conv(h, p[f"blocks.{i}.conv_1.weight"], p.get(f"blocks.{i}.conv_1.bias"), boundary_mode=boundary_mode)
        a1 = act(pre1)
        pre2 = oconv.physics_conv(a1, p[f"blocks.{i}.conv_2.weight"], p.get(f"blocks.{i}.conv_2.bias"), boundary_mode=boundary_mode) + h
        saved.append((h, pre1, a1, pre2))
        h = act(pre2)
    pred = oconv.pointwise_conv(h, p["projection.weight"], p.get("projection.bias"))
    diff = pred - y
    loss = float(np.mean(diff.astype(np.float64) ** 2))
    g = ((2.0 / diff.size) * diff).astype(x.dtype)
    g, gw, gb = oconv.pointwise_conv_vjp(h, p["projection.weight"], g, "projection.bias" in p)
    grads["projection.weight"], grads["projection.bias"] = gw, gb
    for i in reversed(range(num_blocks)):
        hin, pre1, a1, pre2 = saved[i]
        g2 = g * (relu_masks[(i, 1)].astype(pre2.dtype) if relu_masks is not None else act_grad(pre2))
        ga1, gw2, gb2 = oconv.physics_conv_vjp(a1, p[f"blocks.{i}.conv_2.weight"], g2, boundary_mode=boundary_mode)
        g1 = ga1 * (relu_masks[(i, 0)].astype(pre1.dtype) if relu_masks is not None else act_grad(pre1))
        gh, gw1, gb1 = oconv.physics_conv_vjp(hin, p[f"blocks.{i}.conv_1.weight"], g1, boundary_mode=boundary_mode)
        grads[f"blocks.{i}.conv_1.weight"], grads[f"blocks.{i}.conv_1.bias"] = gw1, gb1
        grads[f"blocks.{i}.conv_2.weight"], grads[f"blocks.{i}.conv_2.bias"] = gw2, gb2
        g = gh + g2
    _, gw, gb = oconv.pointwise_conv_vjp(x, p["lifting.weight"], g, "lifting.bias" in p)
    grads["lifting.weight"], grads["lifting.bias"] = gw, gb
    return loss, {k: grads[k] for k in p if grads.get(k) is not None}


def init_dilated_resnet(rng, D, in_channels, out_channels, *, hidden_channels=32, num_blocks=2, kernel_size=3,
                        dilation_rates=(1, 2, 4, 8, 4, 2, 1), use_norm=True, dtype=np.float32):
    """Leaf order of DilatedResNet (Sequential of DilatedResBlock: norm_layers[*], conv_layers[*]; App. C)."""
    p = {}
    p["lifting.weight"], p["lifting.bias"] = oconv.init_conv_weights(rng, D, in_channels, hidden_channels, 1, dtype=dtype)
    for i in range(num_blocks):
        if use_norm:
            for j in range(len(dilation_rates)):
                p[f"blocks.{i}.norm_layers.{j}.weight"] = np.ones(hidden_channels, dtype)
                p[f"blocks.{i}.norm_layers.{j}.bias"] = np.zeros(hidden_channels, dtype)
        for j in range(len(dilation_rates)):
            w, b = oconv.init_conv_weights(rng, D, hidden_channels, hidden_channels, kernel_size, dtype=dtype)
            p[f"blocks.{i}.conv_layers.{j}.weight"], p[f"blocks.{i}.conv_layers.{j}.bias"] = w, b
    p["projection.weight"], p["projection.bias"] = oconv.init_conv_weights(rng, D, hidden_channels, out_channels, 1, dtype=dtype)
    return p


def dilated_resnet_forward(p, x, num_blocks, boundary_mode, dilation_rates=(1, 2, 4, 8, 4, 2, 1), activation="relu"):
    D = p["lifting.weight"].ndim - 2
    use_norm = "blocks.0.norm_layers.0.weight" in p
    h = oconv.pointwise_conv(x, p["lifting.weight"], p.get("lifting.bias"))
    for i in range(num_blocks):
        h = dilated_res_block(h, p, f"blocks.{i}.", boundary_mode, dilation_rates, activation, use_norm, D)
    return oconv.pointwise_conv(h, p["projection.weight"], p.get("projection.bias"))


def dilated_resnet_loss_and_grad(p, x, y, num_blocks, boundary_mode, dilation_rates=(1, 2, 4, 8, 4, 2, 1), activation="relu",
                                 relu_masks=None):
    """mse(dilated_resnet(x), y) and its gradient: DilatedResBlock (_dilated_res_block.py:141-151, in == out) is per rate
    norm -> conv -> act, then + x; analytic reverse mode through group_norm_vjp / physics_conv_vjp.
    `relu_masks` (optional, {(block, layer): bool array}): evaluate act' with the given masks instead of the oracle's own
    pre-activation signs (used to compare against a reduced-precision forward pass whose masks may legitimately differ
    where |pre| is below that precision's error)."""
    act, act_grad = ACTIVATIONS[activation]
    D = p["lifting.weight"].ndim - 2
    use_norm = "blocks.0.norm_layers.0.weight" in p
    grads = {}
    h = oconv.pointwise_conv(x, p["lifting.weight"], p.get("lifting.bias"))
    saved = []
    for i in range(num_blocks):
        layers = []
        hin = h
        for j, d in enumerate(dilation_rates):
            pre_n = h
            if use_norm:
                h = group_norm(h, p[f"blocks.{i}.norm_layers.{j}.weight"], p[f"blocks.{i}.norm_layers.{j}.bias"], 1, D)
            hn = h
            pre = oconv.physics_conv(hn, p[f"blocks.{i}.conv_layers.{j}.weight"], p.get(f"blocks.{i}.conv_layers.{j}.bias"),
                                     boundary_mode=boundary_mode, dilation=d)
            h = act(pre)
            layers.append((pre_n, hn, pre))
        h = h + hin
        saved.append(layers)
    pred = oconv.pointwise_conv(h, p["projection.weight"], p.get("projection.bias"))
    diff = pred - y
    loss = float(np.mean(diff.astype(np.float64) ** 2))
    g = ((2.0 / diff.size) * diff).astype(x.dtype)
    g, gw, gb = oconv.pointwise_conv_vjp(h, p["projection.weight"], g, "projection.bias" in p)
    grads["projection.weight"], grads["projection.bias"] = gw, gb
    for i in reversed(range(num_blocks)):
        gskip = g
        for j in reversed(range(len(dilation_rates))):
            pre_n, hn, pre = saved[i][j]
            mask = relu_masks[(i, j)].astype(pre.dtype) if relu_masks is not None else act_grad(pre)
            gpre = g * mask
            g, gwc, gbc = oconv.physics_conv_vjp(hn, p[f"blocks.{i}.conv_layers.{j}.weight"], gpre, boundary_mode=boundary_mode,
                                                dilation=dilation_rates[j])
            grads[f"blocks.{i}.conv_layers.{j}.weight"], grads[f"blocks.{i}.conv_layers.{j}.bias"] = gwc, gbc
            if use_norm:
                g, gnw, gnb = group_norm_vjp(pre_n, p[f"blocks.{i}.norm_layers.{j}.weight"], 1, D, g)
                grads[f"blocks.{i}.norm_layers.{j}.weight"], grads[f"blocks.{i}.norm_layers.{j}.bias"] = gnw, gnb
        g = g + gskip
    _, gw, gb = oconv.pointwise_conv_vjp(x, p["lifting.weight"], g, "lifting.bias" in p)
    grads["lifting.weight"], grads["lifting.bias"] = gw, gb
    return loss, {k: grads[k] for k in p if grads.get(k) is not None}


# ---------------------------------------------------------------------------
# Generic evaluators over a {path: array} parameter dict (used with the golden fixtures)
# ---------------------------------------------------------------------------

def _norm(h, p, prefix, D, groups=1):
    if prefix + "weight" in p:
        return group_norm(h, p[prefix + "weight"], p[prefix + "bias"], groups, D)
    return h


def res_block_any(x, p, prefix, boundary_mode, activation, D):
    """ClassicResBlock with optional norms and optional 1x1 bypass (_classic_res_block.py:112-121)."""
    act = ACTIVATIONS[activation][0]
    h = oconv.physics_conv(x, p[prefix + "conv_1.weight"], p.get(prefix + "conv_1.bias"), boundary_mode=boundary_mode)
    h = act(_norm(h, p, prefix + "norm_1.", D))
    h = oconv.physics_conv(h, p[prefix + "conv_2.weight"], p.get(prefix + "conv_2.bias"), boundary_mode=boundary_mode)
    h = _norm(h, p, prefix + "norm_2.", D)
    skip = x
    if prefix + "bypass_conv.weight" in p:
        skip = oconv.pointwise_conv(x, p[prefix + "bypass_conv.weight"], p.get(prefix + "bypass_conv.bias"))
        skip = _norm(skip, p, prefix + "bypass_norm.", D)
    return act(h + skip)


def dilated_block_any(x, p, prefix, boundary_mode, activation, D, rates=(1, 2, 4, 8, 4, 2, 1)):
    """DilatedResBlock (_dilated_res_block.py:141-151), in == out."""
    act = ACTIVATIONS[activation][0]
    h = x
    for j, d in enumerate(rates):
        h = _norm(h, p, prefix + f"norm_layers.{j}.", D)
        h = act(oconv.physics_conv(h, p[prefix + f"conv_layers.{j}.weight"], p.get(prefix + f"conv_layers.{j}.bias"),
                                   boundary_mode=boundary_mode, dilation=d))
    return h + x


def modern_block_any(x, p, prefix, boundary_mode, activation, D):
    """ModernResBlock (_modern_res_block.py:133-140): pre-activation ordering, norm -> act -> conv twice, plus the bypass
    (1x1 conv of the normalised input when the channel counts differ, else the identity)."""
    act = ACTIVATIONS[activation][0]
    h = oconv.physics_conv(act(_norm(x, p, prefix + "norm_1.", D)), p[prefix + "conv_1.weight"], p.get(prefix + "conv_1.bias"),
                           boundary_mode=boundary_mode)
    h = oconv.physics_conv(act(_norm(h, p, prefix + "norm_2.", D)), p[prefix + "conv_2.weight"], p.get(prefix + "conv_2.bias"),
                           boundary_mode=boundary_mode)
    skip = x
    if prefix + "bypass_conv.weight" in p:
        skip = oconv.pointwise_conv(_norm(x, p, prefix + "bypass_norm.", D), p[prefix + "bypass_conv.weight"],
                                    p.get(prefix + "bypass_conv.bias"))
    return h + skip


def conv_net_forward(p, x, boundary_mode, activation="relu", final_activation="identity"):
    """ConvNet.__call__ (_conv_net.py:111-114): activation after every layer but the last, final_activation after it."""
    n = 0
    while f"layers.{n}.weight" in p:
        n += 1
    h = x
    for j in range(n):
        h = oconv.physics_conv(h, p[f"layers.{j}.weight"], p.get(f"layers.{j}.bias"), boundary_mode=boundary_mode)
        h = ACTIVATIONS[activation if j < n - 1 else final_activation][0](h)
    return h


def double_conv_block(x, p, prefix, boundary_mode, activation, D):
    """ClassicDoubleConvBlock (_classic_double_conv_block.py:87-90)."""
    act = ACTIVATIONS[activation][0]
    h = x
    for c in ("1", "2"):
        h = oconv.physics_conv(h, p[prefix + f"conv_{c}.weight"], p.get(prefix + f"conv_{c}.bias"), boundary_mode=boundary_mode)
        h = act(_norm(h, p, prefix + f"norm_{c}.", D))
    return h


def sequential_forward(p, x, block_fn, num_blocks):
    """Sequential.__call__ (_sequential.py:111-116) with `block_fn(h, prefix)` per block."""
    h = oconv.pointwise_conv(x, p["lifting.weight"], p.get("lifting.bias"))
    for i in range(num_blocks):
        h = block_fn(h, f"blocks.{i}.")
    return oconv.pointwise_conv(h, p["projection.weight"], p.get("projection.bias"))


def unet_forward(p, x, boundary_mode, num_levels, activation="relu"):
    """Hierarchical.__call__ (_hierarchical.py:251-284) with ClassicUNet's factories (_classic_u_net.py:62-83):
    double-conv lifting, stride-2 conv down (NOT pooling), double-conv arcs, transposed-conv up, concat((skip, x))."""
    D = p["projection.weight"].ndim - 2
    lead = x.ndim - D - 1
    h = double_conv_block(x, p, "lifting.", boundary_mode, activation, D)
    skips = []
    for i in range(num_levels):
        skips.append(h)
        h = oconv.physics_conv(h, p[f"down_sampling_blocks.{i}.weight"], p.get(f"down_sampling_blocks.{i}.bias"),
                               boundary_mode=boundary_mode, stride=2)
        h = double_conv_block(h, p, f"left_arc_blocks.{i}.0.", boundary_mode, activation, D)
    for i in reversed(range(num_levels)):
        skip = skips.pop()
        h = oconv.physics_conv_transpose(h, p[f"up_sampling_blocks.{i}.weight"], p.get(f"up_sampling_blocks.{i}.bias"),
                                         boundary_mode=boundary_mode, stride=2, output_padding=1)
        h = np.concatenate((skip, h), axis=lead)
        h = double_conv_block(h, p, f"right_arc_blocks.{i}.0.", boundary_mode, activation, D)
    return oconv.pointwise_conv(h, p["projection.weight"], p.get("projection.bias"))


def modern_unet_forward(p, x, boundary_mode, num_levels, num_blocks=2, activation="relu"):
    """Hierarchical.__call__ (_hierarchical.py:251-284) with ModernUNet's factories (_modern_u_net.py:78-98): pre-activation
    ResNet blocks as lifting and arcs (num_blocks per level and arc), stride-2 conv down, transposed-conv up,
    concat((skip, x)), 1x1 projection."""
    D = p["projection.weight"].ndim - 2
    lead = x.ndim - D - 1
    h = modern_block_any(x, p, "lifting.", boundary_mode, activation, D)
    skips = []
    for i in range(num_levels):
        skips.append(h)
        h = oconv.physics_conv(h, p[f"down_sampling_blocks.{i}.weight"], p.get(f"down_sampling_blocks.{i}.bias"),
                               boundary_mode=boundary_mode, stride=2)
        for j in range(num_blocks):
            h = modern_block_any(h, p, f"left_arc_blocks.{i}.{j}.", boundary_mode, activation, D)
    for i in reversed(range(num_levels)):
        skip = skips.pop()
        h = oconv.physics_conv_transpose(h, p[f"up_sampling_blocks.{i}.weight"], p.get(f"up_sampling_blocks.{i}.bias"),
                                         boundary_mode=boundary_mode, stride=2, output_padding=1)
        h = np.concatenate((skip, h), axis=lead)
        for j in range(num_blocks):
            h = modern_block_any(h, p, f"right_arc_blocks.{i}.{j}.", boundary_mode, activation, D)
    return oconv.pointwise_conv(h, p["projection.weight"], p.get("projection.bias"))
