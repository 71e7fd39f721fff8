from ._physics_conv import (MorePaddingConv, MorePaddingConvTranspose, PhysicsConv, PhysicsConvTranspose,
                            compute_same_padding)
from ._pointwise_linear_conv import PointwiseLinearConv
from ._spectral_conv import SpectralConv

__all__ = ["MorePaddingConv", "MorePaddingConvTranspose", "PhysicsConv", "PhysicsConvTranspose", "PointwiseLinearConv",
           "SpectralConv", "compute_same_padding"]
