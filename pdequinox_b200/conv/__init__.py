from ._physics_conv import MorePaddingConv, PhysicsConv, compute_same_padding
from ._pointwise_linear_conv import PointwiseLinearConv
from ._spectral_conv import SpectralConv

__all__ = ["MorePaddingConv", "PhysicsConv", "PointwiseLinearConv", "SpectralConv", "compute_same_padding"]
