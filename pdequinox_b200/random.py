"""Splittable host-side PRNG keys for parameter initialisation.

The reference takes `jax.random.PRNGKey`s and splits them in a fixed order. JAX is
not available here, so keys are numpy SeedSequences with the same functional `split` discipline; the
split ORDER and the DISTRIBUTIONS follow the reference constructors, the bit streams cannot.
"""
from __future__ import annotations

import numpy as np


class Key:
    __slots__ = ("entropy", "path")

    def __init__(self, entropy, path=()):
        self.entropy = int(entropy)
        self.path = tuple(path)

    def generator(self) -> np.random.Generator:
        return np.random.default_rng(np.random.SeedSequence(self.entropy, spawn_key=self.path))

    def __repr__(self):
        return f"Key({self.entropy}, {self.path})"


def PRNGKey(seed: int) -> Key:
    return Key(seed)


def _as_key(key) -> Key:
    if isinstance(key, Key):
        return key
    if isinstance(key, (int, np.integer)):
        return Key(int(key))
    raise TypeError(f"key must be a pdequinox_b200.random.Key or an int seed, got {type(key)}")


def split(key, num: int = 2):
    key = _as_key(key)
    return tuple(Key(key.entropy, key.path + (i,)) for i in range(num))


def normal(key, shape):
    return _as_key(key).generator().standard_normal(shape).astype(np.float32)


def uniform(key, shape, minval=0.0, maxval=1.0):
    return _as_key(key).generator().uniform(minval, maxval, shape).astype(np.float32)


def permutation(key, n: int):
    """A random permutation of range(n) (the reference shuffles an epoch with jax.random.permutation, _utils.py:110)."""
    return _as_key(key).generator().permutation(int(n))
