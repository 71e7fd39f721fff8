"""Known-answer cases restated from the reference's own tests (shared by the oracle tests and the
GPU parity tests).

* spectral derivatives: /root/reference/tests/test_spectral_conv.py:19-411 -- a SpectralConv whose
  weights are overwritten with (i k_d)^order on the retained corner blocks differentiates
  prod_a (sin(s_a x_a) + cos(c_a x_a)) analytically; grids 48 (1D), 48x36 (2D), 48x36x24 (3D),
  M = 10; weight tensor shaped (G, C_o = D, C_i = 1, *M) with the corner blocks stacked in
  generate_modes_slices order.
* boundary handling: /root/reference/tests/test_boundary_handling.py:35-52.
"""
from itertools import product

import numpy as np

GRID = {1: (48,), 2: (48, 36), 3: (48, 36, 24)}
NUM_MODES = 10
# (rel, abs) of the reference's pytest.approx per dimension
TOLERANCE = {1: (1e-5, 3e-5), 2: (1e-4, 3e-4), 3: (1e-4, 3e-4)}


def derivative_cases(D):
    """All (sine modes, cosine modes, order) combinations the reference parametrises over."""
    per_axis = list(product([2, 5], [0, 2, 3]))
    return [(tuple(c[0] for c in combo), tuple(c[1] for c in combo), order)
            for combo in product(per_axis, repeat=D) for order in (1, 2)]


def derivative_problem(D, sines, cosines, order, dtype=np.float32):
    """Returns (field (1,*S), weights_real, weights_imag (G, D, 1, *M), expected (D,*S))."""
    S = GRID[D]
    grids = np.meshgrid(*[np.linspace(0, 2 * np.pi, n, endpoint=False) for n in S], indexing="ij")
    f = [np.sin(s * g) + np.cos(c * g) for s, c, g in zip(sines, cosines, grids)]
    if order == 1:
        df = [s * np.cos(s * g) - c * np.sin(c * g) for s, c, g in zip(sines, cosines, grids)]
    else:
        df = [-(s ** 2) * np.sin(s * g) - (c ** 2) * np.cos(c * g) for s, c, g in zip(sines, cosines, grids)]
    field = np.prod(f, axis=0)[None]
    expected = np.stack([np.prod([df[a] if a == d else f[a] for a in range(D)], axis=0) for d in range(D)])
    # (i k_d)^order on the full fft grid, then cut to the corner blocks
    k = np.meshgrid(*[np.fft.fftfreq(n, 1.0 / n) for n in S], indexing="ij")
    op = (1j * np.stack(k)) ** order  # (D, *S)
    M = NUM_MODES
    blocks = []
    for g in range(2 ** (D - 1)):
        idx = [slice(None)]
        for a in range(D - 1):
            idx.append(slice(-M, None) if (g >> a) & 1 else slice(None, M))
        idx.append(slice(None, M))
        blocks.append(op[tuple(idx)])
    w = np.stack(blocks)[:, :, None]  # (G, D, 1, *M)
    return field.astype(dtype), np.real(w).astype(dtype), np.imag(w).astype(dtype), expected.astype(dtype)


BOUNDARY_CASES = {"periodic": [3.0] * 7, "dirichlet": [2.0, 3.0, 3.0, 3.0, 3.0, 3.0, 2.0], "neumann": [3.0] * 7}
