"""jax.random stand-in (splittable numpy SeedSequence keys; bit streams differ from JAX's threefry)."""
import numpy as _np


class _Key:
    def __init__(self, entropy, path=()):
        self.entropy, self.path = int(entropy), tuple(path)

    def rng(self):
        return _np.random.default_rng(_np.random.SeedSequence(self.entropy, spawn_key=self.path))


def PRNGKey(seed):  # noqa: N802
    return _Key(seed)


key = PRNGKey


def split(k, num=2):
    return tuple(_Key(k.entropy, k.path + (i,)) for i in range(num))


def normal(k, shape=(), dtype=_np.float32):
    return k.rng().standard_normal(shape).astype(dtype)


def uniform(k, shape=(), dtype=_np.float32, minval=0.0, maxval=1.0):
    return k.rng().uniform(minval, maxval, shape).astype(dtype)


def permutation(k, n):
    return k.rng().permutation(n)
