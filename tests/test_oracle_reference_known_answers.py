"""Pins the oracle against every known-answer test the reference holds for the hot path (CPU only)."""
import numpy as np
import pytest

import known_answers as ka
from oracle import conv as oconv, nn as onn, spectral as ospec


def _cases(D, limit=None):
    cases = ka.derivative_cases(D)
    if limit:  # 3D: 432 cases in the reference; a strided subset keeps the CPU suite fast
        cases = cases[::max(1, len(cases) // limit)]
    return cases


@pytest.mark.parametrize("D", [1, 2, 3])
@pytest.mark.parametrize("formulation", ["fft", "dense"])
def test_spectral_derivative_known_answers(D, formulation):
    rel, abs_ = ka.TOLERANCE[D]
    f = ospec.spectral_conv_nd if formulation == "fft" else ospec.spectral_conv_dense
    for sines, cosines, order in _cases(D, None if D < 3 else 48):
        field, wr, wi, expected = ka.derivative_problem(D, sines, cosines, order)
        pred = f(field, wr, wi, (ka.NUM_MODES,) * D)
        assert pred.shape == expected.shape
        assert pred == pytest.approx(expected, rel=rel, abs=abs_), (sines, cosines, order)


def test_mode_slice_order_matches_reference_tests():
    # 2D: [:M,:M] then [-M:,:M] (test_spectral_conv.py:180-186); 3D: (lo,lo),(hi,lo),(lo,hi),(hi,hi) (:370-391)
    s2 = ospec.generate_modes_slices((3, 4))
    assert [tuple(s[1:]) for s in s2] == [(slice(None, 3), slice(None, 4)), (slice(-3, None), slice(None, 4))]
    s3 = ospec.generate_modes_slices((2, 3, 4))
    lo0, hi0, lo1, hi1 = slice(None, 2), slice(-2, None), slice(None, 3), slice(-3, None)
    assert [tuple(s[1:3]) for s in s3] == [(lo0, lo1), (hi0, lo1), (lo0, hi1), (hi0, hi1)]


@pytest.mark.parametrize("mode", list(ka.BOUNDARY_CASES))
def test_boundary_handling_known_answers(mode):
    y = oconv.physics_conv(np.ones((1, 7), np.float32), np.ones((1, 1, 3), np.float32), None, boundary_mode=mode)
    assert y == pytest.approx(np.array([ka.BOUNDARY_CASES[mode]], np.float32))


def test_conv_transpose_alignment_known_answers():
    # App. A.3: k=3, stride 2, output_padding 1; x = 1..8, w = [10, 1, 100] -> last element 80 / 180 / 780
    x = np.arange(1, 9, dtype=np.float64)[None]
    w = np.array([10.0, 1.0, 100.0])[None, None]
    for mode, last in [("dirichlet", 80.0), ("periodic", 180.0), ("neumann", 780.0)]:
        y = oconv.physics_conv_transpose(x, w, None, boundary_mode=mode, stride=2, output_padding=1)
        assert y.shape == (1, 16)
        assert y[0, -1] == pytest.approx(last)
        assert y[0, 0::2] == pytest.approx(x[0])  # centre tap lands on even outputs (test_unet_translation.py)


@pytest.mark.parametrize("D,expect", [(1, 102625)])
def test_fno_parameter_count(D, expect):
    # docs/examples/parameter_count_and_receptive_field.ipynb:541
    p = onn.init_fno(np.random.default_rng(0), D, 1, 1)
    assert onn.count_parameters(p) == expect


def test_resnet_parameter_counts():
    # notebook :380 (37345) and :427 with norm (38113)
    rng = np.random.default_rng(0)
    assert onn.count_parameters(onn.init_classic_resnet(rng, 1, 1, 1)) == 37345
    assert onn.count_parameters(onn.init_classic_resnet(rng, 1, 1, 1, use_norm=True)) == 38113


def test_same_padding():
    # _physics_conv.py:11-26: odd kernels symmetric, even kernels put the extra on the high side
    assert oconv.compute_same_padding(1, 3, 1) == ((1, 1),)
    assert oconv.compute_same_padding(1, 2, 1) == ((0, 1),)
    assert oconv.compute_same_padding(2, (4, 3), (2, 8)) == ((2, 4), (8, 8))
