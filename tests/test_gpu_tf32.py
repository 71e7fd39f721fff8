"""TF32 mode: the tcgen05 kernels (first-stage r2c DFT, last-stage slab kernel, bypass weight gradient)
against the float64 oracle, tolerance rel-L2 <= 2e-3 (BASELINE.json north_star, "stated TF32 mode"),
and against the library's own fp32 path."""
import numpy as np
import pytest
import torch

import pdequinox_b200 as pdeqx
from pdequinox_b200 import _lib
from pdequinox_b200.conv._spectral_conv import get_plan
from pdequinox_b200.train import ParameterArena
from oracle import conv as oconv, nn as onn, spectral as ospec
from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 2e-3


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def host(t):
    return t.detach().cpu().numpy()


CASES = [
    # (D, B, ci, co, spatial, modes)
    (2, 2, 32, 32, (16, 128), (4, 16)),
    (2, 1, 64, 64, (8, 256), (4, 16)),
    (2, 3, 32, 64, (12, 128), (3, 16)),
    (2, 2, 64, 32, (6, 256), (3, 16)),
    (1, 4, 32, 32, (128,), (16,)),
    (3, 1, 32, 32, (4, 6, 128), (2, 3, 16)),
    (2, 2, 64, 64, (64, 256), (16, 16)),
    (2, 3, 64, 64, (32, 128), (8, 16)),
    (2, 1, 64, 64, (96, 384), (16, 16)),   # three w-tiles, odd number of units per CTA
    # mode counts other than 16 are zero-padded in the data layout; grids narrower than 128 put several rows in a slab unit
    (3, 2, 32, 32, (32, 32, 64), (12, 12, 12)),  # BASELINE configs[4] geometry scaled down: 12 modes, W = 64
    (2, 2, 32, 32, (64, 64), (12, 12)),
    (2, 2, 64, 64, (64, 128), (8, 12)),
    (2, 3, 32, 32, (32, 32), (4, 8)),            # W = 32: four rows per unit
    # >= 16 samples and >= 64 channels: the per-mode channel mix and its two adjoints run on tcgen05 (tc_modemix.cu)
    (2, 16, 64, 64, (16, 128), (4, 16)),
    (2, 40, 64, 64, (8, 128), (4, 12)),          # ragged batch tile, a half-valid 8-mode tile (12 modes in a row of 16)
    (2, 72, 64, 64, (8, 128), (4, 8)),           # two batch tiles (the second one ragged), 18 K chunks in the weight gradient
    (1, 32, 64, 64, (128,), (16,)),
    (3, 32, 64, 64, (4, 6, 128), (2, 3, 16)),
]


@pytest.mark.parametrize("D,B,ci,co,spatial,modes", CASES)
@pytest.mark.parametrize("act", ["gelu", "identity", "relu"])
def test_tf32_spectral_block(cuda, D, B, ci, co, spatial, modes, act):
    pdeqx.set_precision("tf32")
    try:
        rng = np.random.default_rng(11)
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

        plan = get_plan(D, B, ci, co, spatial, modes)
        mask = _lib.lib().pdeqx_spectral_plan_tcgen05_stages(plan.handle)
        assert mask & 7 == 7, f"tcgen05 stages not enabled for this shape (mask {mask})"
        assert bool(mask & 8) == (B >= 16 and min(ci, co) >= 64), f"channel mix on the wrong path (mask {mask})"

        blk = pdeqx.ClassicSpectralBlock(D, ci, co, activation=getattr(pdeqx.nn, act), num_modes=modes, key=0)
        blk.spectral_conv.weights_real, blk.spectral_conv.weights_imag = dev(wr), dev(wi)
        blk.by_pass_conv.weight, blk.by_pass_conv.bias = dev(bw), dev(bb)
        arena = ParameterArena(blk)
        tape = pdeqx._module.Tape()
        y = blk._fwd(dev(x), tape)
        torch.cuda.synchronize()
        assert rel_l2(host(y), ref_y) <= TOL
        if act == "relu":
            # relu' is discontinuous: where |pre| is below the TF32 error the mask may legitimately differ from the
            # float64 oracle's; evaluate the oracle's reverse pass with the mask of the forward pass under test
            gpre = gy * (host(y) > 0)
            gx_s, gwr, gwi, _ = ospec.spectral_conv_vjp(x, wr, wi, modes, gpre)
            gx_b, gbw, gbb = oconv.pointwise_conv_vjp(x, bw, gpre)
        gx = blk._bwd(dev(gy), tape, True)
        torch.cuda.synchronize()
        g = {k: host(v) for k, v in arena.named_grads().items()}
        assert rel_l2(host(gx), gx_s + gx_b) <= TOL
        assert rel_l2(g["spectral_conv.weights_real"], gwr) <= TOL
        assert rel_l2(g["spectral_conv.weights_imag"], gwi) <= TOL
        assert rel_l2(g["by_pass_conv.weight"], gbw) <= TOL
        assert rel_l2(g["by_pass_conv.bias"], gbb) <= TOL
    finally:
        pdeqx.set_precision("fp32")


def test_tf32_plain_spectral_conv(cuda):
    pdeqx.set_precision("tf32")
    try:
        rng = np.random.default_rng(12)
        x = rng.standard_normal((2, 32, 16, 128))
        wr = rng.standard_normal((2, 32, 32, 4, 16)) / 1024
        wi = rng.standard_normal((2, 32, 32, 4, 16)) / 1024
        conv = pdeqx.SpectralConv(2, 32, 32, (4, 16), key=0)
        conv.weights_real, conv.weights_imag = dev(wr), dev(wi)
        y = host(pdeqx.vmap(conv)(dev(x)))
        assert rel_l2(y, ospec.spectral_conv_nd(x, wr, wi, (4, 16))) <= TOL
    finally:
        pdeqx.set_precision("fp32")


@pytest.mark.parametrize("D,spatial,hidden,modes,blocks", [
    (2, (64, 128), 32, 16, 2),   # projection folded into the last block (rank-one cotangent), lifting into the first
    (2, (32, 128), 64, 16, 1),   # one block: both folds in the same reverse pass
    (2, (64, 64), 32, 12, 3),    # padded modes, two rows per slab unit
    (3, (32, 32, 64), 32, 12, 2),
])
def test_tf32_fno_train_step(cuda, D, spatial, hidden, modes, blocks):
    """Whole ClassicFNO train step in TF32 mode vs the oracle."""
    from pdequinox_b200.train import TrainStep
    pdeqx.set_precision("tf32")
    try:
        rng = np.random.default_rng(13)
        model = pdeqx.ClassicFNO(D, 1, 1, hidden_channels=hidden, num_modes=modes, num_blocks=blocks, key=5)
        step = TrainStep(model)
        # non-zero lifting / projection biases so that every folded gradient is exercised
        with torch.no_grad():
            named = step.arena.named_params()
            named["lifting.bias"].add_(0.3)
            named["projection.bias"].add_(-0.2)
        p = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in step.arena.named_params().items()}
        x = rng.standard_normal((2, 1) + spatial)
        y = rng.standard_normal((2, 1) + spatial)
        loss_ref, g_ref = onn.fno_loss_and_grad(p, x, y, (modes,) * D, blocks)
        loss = step.loss_and_grad(x, y).item()
        assert abs(loss - loss_ref) <= TOL * abs(loss_ref)
        g = {k: v.detach().cpu().numpy() for k, v in step.arena.named_grads().items()}
        for k in g_ref:
            assert rel_l2(g[k], g_ref[k]) <= TOL, k  # whole-model bar = the north star's 2e-3, chained stages included
    finally:
        pdeqx.set_precision("fp32")


@pytest.mark.parametrize("spatial,hidden,blocks,act", [
    ((64, 128), 32, 2, "gelu"),
    ((32, 256), 64, 1, "relu"),   # one block: lifting and projection folded into the same forward call
    ((128, 128), 64, 3, "gelu"),
])
def test_tf32_folded_forward_equals_plain_forward(cuda, spatial, hidden, blocks, act):
    """The training-tape forward pass (1 -> C lifting and C -> 1 projection folded into the first / last spectral block:
    neither full-size tensor is written) gives the same prediction as the plain layer-by-layer forward pass and as the
    oracle (pdequinox/arch/_classic_fno.py:71-83 around pdequinox/blocks/_classic_spectral_block.py:72-75)."""
    from pdequinox_b200._module import Tape
    from pdequinox_b200 import nn as pnn
    pdeqx.set_precision("tf32")
    try:
        rng = np.random.default_rng(21)
        model = pdeqx.ClassicFNO(2, 1, 1, hidden_channels=hidden, num_modes=16, num_blocks=blocks,
                                 activation=getattr(pnn, act), key=9)
        with torch.no_grad():
            model.lifting.bias.add_(0.25)
            model.projection.bias.add_(-0.4)
        x = rng.standard_normal((3, 1) + spatial)
        plain = host(model._fwd(dev(x), None))
        tape = Tape()
        tape.fold_lifting = True
        folded = host(model._fwd(dev(x), tape))
        assert tape.stack[-1] is None, "the projection was not folded into the last block"
        assert rel_l2(folded, plain) <= 2e-4
        p = {k: host(getattr(o, f)).astype(np.float64) for k, o, f in model.named_leaves()}
        ref = onn.fno_forward(p, x, (16, 16), blocks, activation=act)
        assert rel_l2(folded, ref) <= TOL
    finally:
        pdeqx.set_precision("fp32")


@pytest.mark.parametrize("D,spatial,hidden,modes,blocks,fold", [
    (2, (64, 128), 32, 16, 3, True),     # lifted first block (rank-one bypass in the chained pass), projected last block
    (2, (32, 256), 64, 16, 3, False),    # plain chain: every block reads its stored input
    (3, (32, 32, 64), 32, 12, 2, True),  # two image rows per unit
])
def test_tf32_chained_reverse_pass_equals_unchained(cuda, D, spatial, hidden, modes, blocks, fold):
    """The chained reverse pass (data gradient of block k and pre-activation cotangent of block k-1 in ONE kernel: the tensor
    between them is never written) gives the gradients of the block-by-block reverse pass; the only arithmetic difference is
    that the input gradient stays in the fp32 accumulator instead of taking a round trip through HBM."""
    from pdequinox_b200._module import Tape
    from pdequinox_b200._sequential import Sequential
    from pdequinox_b200 import _lib
    pdeqx.set_precision("tf32")
    try:
        rng = np.random.default_rng(17)
        model = pdeqx.ClassicFNO(D, 1, 1, hidden_channels=hidden, num_modes=modes, num_blocks=blocks, key=6)
        arena = ParameterArena(model)
        x = dev(rng.standard_normal((2, 1) + spatial))
        gy = dev(rng.standard_normal((2, 1) + spatial))
        grads = {}
        for chained in (False, True):
            Sequential.chain_reverse = chained
            tape = Tape()
            tape.fold_lifting = fold
            arena.grads.zero_()
            before = _lib.lib().pdeqx_launch_count()
            model._fwd(x, tape)
            model._bwd(gy, tape, False)
            torch.cuda.synchronize()
            grads[chained] = ({k: host(v).copy() for k, v in arena.named_grads().items()}, _lib.lib().pdeqx_launch_count() - before)
        assert grads[True][1] < grads[False][1], "the chained reverse pass did not run (same number of kernel launches)"
        for k in grads[False][0]:
            assert rel_l2(grads[True][0][k], grads[False][0][k]) <= 5e-4, k
    finally:
        Sequential.chain_reverse = True
        pdeqx.set_precision("fp32")
