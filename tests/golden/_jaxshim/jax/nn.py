"""jax.nn stand-in: relu and gelu(approximate=True) from their documented definitions."""
import numpy as _np


def relu(x):
    return _np.maximum(x, 0)


def gelu(x, approximate=True):
    x = _np.asarray(x)
    if approximate:
        c = _np.sqrt(2.0 / _np.pi).astype(x.dtype) if hasattr(_np.sqrt(2.0 / _np.pi), "astype") else _np.sqrt(2.0 / _np.pi)
        return (0.5 * x * (1.0 + _np.tanh(c * (x + 0.044715 * x ** 3)))).astype(x.dtype)
    from scipy.special import erf
    return (0.5 * x * (1.0 + erf(x / _np.sqrt(2.0)))).astype(x.dtype)
