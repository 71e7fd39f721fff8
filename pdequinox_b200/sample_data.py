"""Synthetic training data of the reference's quickstart (host side, NumPy): pdequinox/_sample_data.py:6-71.

BASELINE.json configs[0] / [1] train on `poisson_1d_dirichlet`; the distribution of the samples is the reference's, the
random stream is this package's host PRNG (a JAX key cannot be replayed here)."""
import numpy as np

from . import random as jr


def poisson_1d_dirichlet(num_points: int = 64, num_samples: int = 1000, *, domain_extent: float = 5.0, key):
    """Pairs (force, displacement), each `[num_samples, 1, num_points]` float32, of the 1-D Poisson problem -u'' = f with
    homogeneous Dirichlet ends on `num_points` interior grid points of (0, domain_extent). A force field is the indicator
    of a random interval [a, b], a ~ U(0.2, 0.4) L, b ~ U(0.6, 0.8) L; the displacement solves the three-point
    finite-difference system exactly."""
    grid = np.linspace(0.0, domain_extent, num_points + 2)[1:-1]
    dx = grid[1] - grid[0]
    lower, upper = np.empty(num_samples), np.empty(num_samples)
    for i, k in enumerate(jr.split(key, num_samples)):  # one key per sample, split into the two interval ends (:48-56,:62)
        k_lo, k_hi = jr.split(k)
        lower[i] = jr.uniform(k_lo, (), 0.2 * domain_extent, 0.4 * domain_extent)
        upper[i] = jr.uniform(k_hi, (), 0.6 * domain_extent, 0.8 * domain_extent)
    force = ((grid[None, :] >= lower[:, None]) & (grid[None, :] <= upper[:, None])).astype(np.float64)
    # (u[i-1] - 2 u[i] + u[i+1]) / dx^2 = -f[i], u[-1] = u[n] = 0: Thomas algorithm on the constant tridiagonal matrix,
    # all right-hand sides at once
    n = num_points
    rhs = (-force * dx * dx).T.copy()  # [n, samples]
    cp = np.empty(n)
    cp[0] = 1.0 / -2.0
    rhs[0] /= -2.0
    for i in range(1, n):
        denom = -2.0 - cp[i - 1]
        cp[i] = 1.0 / denom
        rhs[i] = (rhs[i] - rhs[i - 1]) / denom
    for i in range(n - 2, -1, -1):
        rhs[i] -= cp[i] * rhs[i + 1]
    displacement = rhs.T
    return force[:, None, :].astype(np.float32), displacement[:, None, :].astype(np.float32)
