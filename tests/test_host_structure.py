"""CPU-only checks of the host layer: the C ABI loads and exports every declared symbol, constructors
mirror the reference (signatures, errors, pytree leaf order, parameter counts), and the product never
touches the oracle."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

import pdequinox_b200 as pdeqx
from pdequinox_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shared_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pdeqx_b200.h")).read()
    declared = set(re.findall(r"\b(pdeqx_[A-Za-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} is declared in include/pdeqx_b200.h but not exported"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert b"sm_100a" in _lib.lib().pdeqx_version()


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    conv = pdeqx.PhysicsConv(1, 1, 1, 3, key=0, boundary_mode="periodic")
    with pytest.raises(_lib.PdeqxError):
        conv(np.ones((1, 7), np.float32))


@pytest.mark.parametrize("ctor,expect", [
    (lambda: pdeqx.ClassicFNO(1, 1, 1, key=0), 102625),            # parameter_count_and_receptive_field.ipynb:541
    (lambda: pdeqx.ClassicResNet(1, 1, 1, key=0), 37345),          # :380
    (lambda: pdeqx.ClassicResNet(1, 1, 1, use_norm=True, key=0), 38113),  # :427
    (lambda: pdeqx.DilatedResNet(1, 1, 1, key=0), 44449),          # :683
    (lambda: pdeqx.ClassicFNO(2, 1, 1, hidden_channels=64, num_modes=16, key=0), 16794049),   # BASELINE configs[3]
    (lambda: pdeqx.ClassicResNet(2, 1, 1, hidden_channels=64, key=0), 443329),                # BASELINE configs[2]
    (lambda: pdeqx.DilatedResNet(2, 1, 1, hidden_channels=64, key=0), 518977),
    (lambda: pdeqx.ClassicUNet(1, 1, 1, key=0), 789089),           # README defaults (SURVEY.md 8d cfg 2)
    (lambda: pdeqx.ClassicUNet(1, 1, 1, hidden_channels=32, num_levels=2, boundary_mode="dirichlet", key=0), 189633),  # train_unet_as_poisson_solver.ipynb:225-235
    (lambda: pdeqx.ClassicUNet(2, 1, 1, key=0), 2357441),
    (lambda: pdeqx.ClassicUNet(3, 1, 1, key=0), 7062497),
])
def test_parameter_counts_match_reference(ctor, expect):
    assert pdeqx.count_parameters(ctor()) == expect


@pytest.mark.parametrize("D", [1, 2, 3])
def test_primitive_parameter_counts(D):
    # reference tests/test_count_parameters.py:15-145
    for k, stride, dil in [(2, 1, 1), (3, 2, 2), (4, 1, 3), (5, 2, 1)]:
        conv = pdeqx.PhysicsConv(D, 2, 5, k, stride=stride, dilation=dil, key=0, boundary_mode="periodic")
        assert pdeqx.count_parameters(conv) == 2 * 5 * k ** D + 5
        conv = pdeqx.PhysicsConv(D, 2, 5, k, use_bias=False, key=0, boundary_mode="dirichlet")
        assert pdeqx.count_parameters(conv) == 2 * 5 * k ** D
    sc = pdeqx.SpectralConv(D, 2, 5, 4, key=0)
    assert pdeqx.count_parameters(sc) == 2 * 5 * 4 ** D * 2 * 2 ** (D - 1)
    assert tuple(sc.weights_real.shape) == (2 ** (D - 1), 2, 5) + (4,) * D  # ALLOCATION shape (G, in, out, *modes)
    pw = pdeqx.PointwiseLinearConv(D, 2, 5, key=0)
    assert pdeqx.count_parameters(pw) == 2 * 5 + 5
    assert tuple(pw.weight.shape) == (5, 2) + (1,) * D and tuple(pw.bias.shape) == (5,) + (1,) * D


def test_leaf_order_matches_reference_pytree():
    m = pdeqx.ClassicFNO(2, 1, 1, hidden_channels=4, num_modes=2, num_blocks=2, key=0)
    assert pdeqx.tree_paths(m) == [
        "lifting.weight", "lifting.bias",
        "blocks.0.spectral_conv.weights_real", "blocks.0.spectral_conv.weights_imag",
        "blocks.0.by_pass_conv.weight", "blocks.0.by_pass_conv.bias",
        "blocks.1.spectral_conv.weights_real", "blocks.1.spectral_conv.weights_imag",
        "blocks.1.by_pass_conv.weight", "blocks.1.by_pass_conv.bias",
        "projection.weight", "projection.bias"]
    r = pdeqx.ClassicResNet(1, 1, 1, hidden_channels=4, num_blocks=1, use_norm=True, key=0)
    assert pdeqx.tree_paths(r)[2:10] == [
        "blocks.0.conv_1.weight", "blocks.0.conv_1.bias", "blocks.0.norm_1.weight", "blocks.0.norm_1.bias",
        "blocks.0.conv_2.weight", "blocks.0.conv_2.bias", "blocks.0.norm_2.weight", "blocks.0.norm_2.bias"]


def test_unet_leaf_order_and_shared_arc_keys():
    m = pdeqx.ClassicUNet(1, 1, 1, hidden_channels=4, num_levels=2, key=0)
    paths = pdeqx.tree_paths(m)
    # dataclass-field order of Hierarchical (_hierarchical.py:21-30): lifting, down, left, up, right, projection
    first = [p.split(".")[0] for p in paths]
    order = [k for i, k in enumerate(first) if i == 0 or first[i - 1] != k]
    assert order == ["lifting", "down_sampling_blocks", "left_arc_blocks", "up_sampling_blocks", "right_arc_blocks",
                     "projection"]
    assert "left_arc_blocks.1.0.conv_2.weight" in paths and "right_arc_blocks.0.0.norm_1.bias" in paths
    assert tuple(m.up_sampling_blocks[1].weight.shape) == (8, 16, 3)       # fan_out -> fan_out // 2, weight (C_out, C_in, k)
    assert tuple(m.right_arc_blocks[1][0].conv_1.weight.shape) == (8, 16, 3)  # concat(skip 8, up 8) -> 8
    assert m.down_sampling_blocks[0].stride == (2,) and m.up_sampling_blocks[0].output_padding == (1,)
    # every block of one arc is built from the same key (_hierarchical.py:196,208): identical initial weights
    h = pdeqx.Hierarchical(1, 1, 1, hidden_channels=4, num_levels=1, num_blocks=3, activation=pdeqx.nn.relu, key=0,
                           boundary_mode="periodic")
    a, b = h.left_arc_blocks[0][1], h.left_arc_blocks[0][2]
    assert bool((a.conv_1.weight == b.conv_1.weight).all())
    assert pdeqx.ClassicUNet(1, 1, 1, key=0).receptive_field == tuple(pdeqx.ClassicUNet(1, 1, 1, key=1).receptive_field)
    with pytest.raises(ValueError):  # _physics_conv.py:211-215
        pdeqx.PhysicsConvTranspose(1, 1, 1, 3, key=0, boundary_mode="robin")


def test_constructor_signatures_and_errors():
    sig = inspect.signature(pdeqx.PhysicsConv.__init__)
    assert list(sig.parameters)[1:9] == ["num_spatial_dims", "in_channels", "out_channels", "kernel_size", "stride",
                                         "dilation", "groups", "use_bias"]
    assert sig.parameters["key"].kind is inspect.Parameter.KEYWORD_ONLY
    assert sig.parameters["boundary_mode"].kind is inspect.Parameter.KEYWORD_ONLY
    sig = inspect.signature(pdeqx.SpectralConv.__init__)
    assert list(sig.parameters)[1:6] == ["num_spatial_dims", "in_channels", "out_channels", "num_modes", "key"]
    with pytest.raises(ValueError):  # _spectral_conv.py:48-49
        pdeqx.SpectralConv(2, 1, 1, (3, 3, 3), key=0)
    with pytest.raises(ValueError):  # _physics_conv.py:98-101
        pdeqx.PhysicsConv(1, 1, 1, 3, key=0, boundary_mode="robin")
    with pytest.raises(ValueError):  # _conv.py:110-114
        pdeqx.PhysicsConv(1, 3, 4, 3, groups=2, key=0, boundary_mode="periodic")
    with pytest.raises(ValueError):  # _sequential.py:67-68
        pdeqx.ClassicFNO(1, 1, 1, num_blocks=0, key=0)
    assert pdeqx.PhysicsConv(1, 1, 1, 3, key=0, boundary_mode="PERIODIC").boundary_mode == "periodic"


def test_receptive_fields():
    # reference tests/test_receptive_field.py:8-29 and the notebook values (:959, :1093, :1124)
    assert pdeqx.PhysicsConv(1, 1, 1, 3, key=0, boundary_mode="periodic").receptive_field == ((1.0, 1.0),)
    assert pdeqx.SpectralConv(1, 1, 1, 4, key=0).receptive_field == ((float("inf"), float("inf")),)
    assert pdeqx.PointwiseLinearConv(1, 1, 1, key=0).receptive_field == ((0.0, 0.0),)
    assert pdeqx.ClassicResNet(1, 1, 1, key=0).receptive_field == ((12.0, 12.0),)
    assert pdeqx.DilatedResNet(1, 1, 1, key=0).receptive_field == ((44.0, 44.0),)
    assert pdeqx.ClassicFNO(1, 1, 1, key=0).receptive_field == ((float("inf"), float("inf")),)


def test_init_distributions_follow_reference():
    conv = pdeqx.PhysicsConv(2, 8, 16, 3, key=1, boundary_mode="periodic")
    lim = 1.0 / np.sqrt(8 * 9)  # _conv.py:129
    w = conv.weight.cpu().numpy()
    assert np.abs(w).max() <= lim and np.abs(w).max() > 0.9 * lim
    sc = pdeqx.SpectralConv(1, 8, 8, 16, key=2)
    assert abs(sc.weights_real.cpu().numpy().std() * 64 - 1.0) < 0.1  # scale 1/(in*out) * N(0,1)
    z = pdeqx.PhysicsConv(1, 2, 2, 3, key=0, boundary_mode="periodic", zero_bias_init=True)
    assert float(z.bias.abs().sum()) == 0.0
    a = pdeqx.ClassicFNO(1, 1, 1, key=7)
    b = pdeqx.ClassicFNO(1, 1, 1, key=7)
    assert all(bool((p == q).all()) for p, q in zip(pdeqx.tree_leaves(a), pdeqx.tree_leaves(b)))  # same key, same init


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pdequinox_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "oracle/" not in src and "oracle." not in src.replace("oracle.md", ""), f"{f} references the oracle"


def test_gradient_buckets_tile_the_arena():
    """The overlapped exchange all-reduces one contiguous arena range per finished pytree prefix (projection, blocks in
    reverse, lifting): the ranges must tile the flat gradient arena exactly, or a gradient would stay un-reduced."""
    from pdequinox_b200.train import TrainStep
    m = pdeqx.ClassicFNO(2, 1, 1, hidden_channels=8, num_modes=4, num_blocks=3, key=0)
    st = TrainStep(m, distributed=False)
    ranges = sorted(st._prefix_range(p) for p in ["projection.", "blocks.2.", "blocks.1.", "blocks.0.", "lifting."])
    assert ranges[0][0] == 0 and ranges[-1][1] == st.arena.size
    for (a0, a1), (b0, b1) in zip(ranges[:-1], ranges[1:]):
        assert a1 == b0
    assert st._prefix_range("no_such_prefix.") == (None, None)
