"""equinox stand-in (TEST TOOLING ONLY): Module as a plain Python class, nn.Conv / GroupNorm / Identity from
their documented definitions. Only what /root/reference/pdequinox imports."""
import numpy as _np

from . import nn  # noqa: F401
from .nn import Module, field  # noqa: F401


def is_array(x):
    return isinstance(x, _np.ndarray)


def tree_at(where, pytree, replace):
    """Functional attribute replacement for the simple `lambda m: m.attr[.attr...]` selectors the tests use."""
    import copy
    new = copy.deepcopy(pytree)

    class _Probe:
        def __init__(self, path=()):
            object.__setattr__(self, "_path", path)

        def __getattr__(self, name):
            return _Probe(self._path + (name,))

        def __getitem__(self, i):
            return _Probe(self._path + (i,))

    path = where(_Probe())._path
    obj = new
    for p in path[:-1]:
        obj = obj[p] if isinstance(p, int) else getattr(obj, p)
    last = path[-1]
    if isinstance(last, int):
        obj[last] = replace
    else:
        object.__setattr__(obj, last, replace)
    return new


def filter(tree, fn):  # noqa: A001
    """Only used as eqx.filter(model, eqx.is_array) -> the array leaves."""
    from jax.tree_util import tree_leaves
    return [leaf for leaf in tree_leaves(tree) if fn(leaf)]


def partition(*a, **k):
    raise NotImplementedError


def combine(*a, **k):
    raise NotImplementedError


def filter_vmap(*a, **k):
    raise NotImplementedError


def filter_jit(f):
    return f
