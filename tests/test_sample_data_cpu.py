"""pdequinox_b200.sample_data.poisson_1d_dirichlet: the contract of pdequinox/_sample_data.py:6-71 (shapes, the
distribution of the force fields, displacement = exact solution of the three-point finite-difference system)."""
import numpy as np

import pdequinox_b200 as pdeqx
from oracle import nn as onn


def test_poisson_1d_dirichlet_shapes_distribution_and_solution():
    n, m, L = 64, 200, 5.0
    f, u = pdeqx.sample_data.poisson_1d_dirichlet(n, m, domain_extent=L, key=pdeqx.random.PRNGKey(0))
    assert f.shape == u.shape == (m, 1, n) and f.dtype == u.dtype == np.float32
    grid = np.linspace(0, L, n + 2)[1:-1]
    dx = grid[1] - grid[0]
    assert set(np.unique(f)) <= {0.0, 1.0}
    for s in range(m):  # one interval, starting in [0.2 L, 0.4 L + dx) and ending in (0.6 L - dx, 0.8 L]
        on = np.flatnonzero(f[s, 0])
        assert np.array_equal(on, np.arange(on[0], on[-1] + 1))
        assert 0.2 * L <= grid[on[0]] < 0.4 * L + dx and 0.6 * L - dx < grid[on[-1]] <= 0.8 * L
    assert len({tuple(r) for r in f[:, 0]}) > 20  # the samples differ
    # residual of the discrete operator with homogeneous Dirichlet ends
    up = np.pad(u[:, 0].astype(np.float64), ((0, 0), (1, 1)))
    lap = (up[:, :-2] - 2 * up[:, 1:-1] + up[:, 2:]) / dx ** 2
    assert np.abs(lap + f[:, 0]).max() < 1e-3
    assert (u >= 0).all()  # a non-negative load bends the string one way


def test_poisson_1d_dirichlet_matches_the_dense_solve_of_the_oracle_restatement():
    n, L = 33, 2.0
    f, u = pdeqx.sample_data.poisson_1d_dirichlet(n, 16, domain_extent=L, key=3)
    grid = np.linspace(0, L, n + 2)[1:-1]
    dx = grid[1] - grid[0]
    A = (np.diag(np.ones(n - 1), -1) - 2 * np.eye(n) + np.diag(np.ones(n - 1), 1)) / dx ** 2
    ref = np.linalg.solve(A, -f[:, 0].astype(np.float64).T).T  # what pdequinox/_sample_data.py:45-46 does
    assert np.abs(u[:, 0] - ref).max() <= 1e-6 * np.abs(ref).max()
    f2, u2 = onn.poisson_1d_dirichlet(np.random.default_rng(0), n, 16, L)
    assert f2.shape == f.shape and u2.shape == u.shape


def test_same_key_same_data():
    a = pdeqx.sample_data.poisson_1d_dirichlet(16, 8, key=5)
    b = pdeqx.sample_data.poisson_1d_dirichlet(16, 8, key=5)
    c = pdeqx.sample_data.poisson_1d_dirichlet(16, 8, key=6)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and not np.array_equal(a[0], c[0])
