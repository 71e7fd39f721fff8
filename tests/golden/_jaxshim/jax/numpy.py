"""jax.numpy stand-in: numpy with float32 defaults and the functional `.at[idx].set(v)` update."""
import numpy as _np
from numpy import (abs, arange, ceil, concatenate, cos, diag, exp, expand_dims, inf, linspace, maximum, mean, meshgrid,  # noqa: F401,A004
                   ones, pi, prod, sin, sqrt, squeeze, stack, sum, tanh, where, zeros_like, float32, float64, complex64)
from numpy import linalg  # noqa: F401

ndarray = _np.ndarray


class _AtIndex:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, value):
        out = _np.array(self.arr, copy=True).view(Arr)
        out[self.idx] = value
        return out


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self.arr, idx)


class Arr(_np.ndarray):
    @property
    def at(self):
        return _At(self)


def asarray(x, dtype=None):
    return _np.asarray(x, dtype=dtype).view(Arr)


array = asarray


def zeros(shape, dtype=_np.float32):
    return _np.zeros(shape, dtype=dtype).view(Arr)


def pad(x, pad_width, mode="constant", **kw):
    return _np.pad(x, pad_width, mode=mode, **kw).view(Arr)


def einsum(subscripts, *operands):
    return _np.einsum(subscripts, *operands).view(Arr)


class fft:  # noqa: N801
    @staticmethod
    def rfftn(x, s=None, axes=None):
        x = _np.asarray(x)
        axes = tuple(range(x.ndim - len(s), x.ndim)) if (axes is None and s is not None) else axes
        out = _np.fft.rfftn(x, s=s, axes=axes)
        return out.astype(_np.complex64 if x.dtype == _np.float32 else _np.complex128).view(Arr)

    @staticmethod
    def irfftn(x, s=None, axes=None):
        x = _np.asarray(x)
        axes = tuple(range(x.ndim - len(s), x.ndim)) if (axes is None and s is not None) else axes
        out = _np.fft.irfftn(x, s=s, axes=axes)
        return out.astype(_np.float32 if x.dtype == _np.complex64 else _np.float64).view(Arr)

    rfftfreq = staticmethod(_np.fft.rfftfreq)
    fftfreq = staticmethod(_np.fft.fftfreq)
