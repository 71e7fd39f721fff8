"""Minimal NumPy stand-in for the parts of `jax` the reference's hot path uses (TEST TOOLING ONLY).

jax, equinox and optax cannot be installed in this image. To pin the oracle against the reference's OWN
Python source (slice order, padding arithmetic, weight interpretation, block wiring), make_golden.py
executes /root/reference/pdequinox unmodified on top of this shim and records input/output vectors.
The third-party numerics are restated independently of the oracle: FFTs are numpy's pocketfft (the
algorithm the JAX CPU backend uses), conv_general_dilated is evaluated with torch.nn.functional.conv*d.
"""
import functools

import numpy as _np

from . import lax, nn, numpy, random, tree_util  # noqa: F401

Array = _np.ndarray


def named_scope(name):
    def deco(fn):
        return fn
    return deco


def vmap(fn, in_axes=0, out_axes=0):
    @functools.wraps(fn)
    def mapped(x):
        return numpy.stack([fn(xi) for xi in x])
    return mapped


def jit(fn, **kw):
    return fn


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()
