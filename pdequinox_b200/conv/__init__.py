"""Convolutions of the hot path (reference package: pdequinox/conv)."""
from . import _physics_conv, _pointwise_linear_conv, _spectral_conv

__all__ = ["MorePaddingConv", "MorePaddingConvTranspose", "PhysicsConv", "PhysicsConvTranspose", "PointwiseLinearConv",
           "SpectralConv", "compute_same_padding"]
_HOMES = (_physics_conv, _pointwise_linear_conv, _spectral_conv)
for _name in __all__:
    globals()[_name] = next(getattr(m, _name) for m in _HOMES if hasattr(m, _name))
del _name
