"""LinearConvDownBlock: strided PhysicsConv (reference: pdequinox/blocks/_linear_conv_down_block.py:9-84).
The factory drops `zero_bias_init` exactly like the reference's (`:74-83`)."""
from ..conv import PhysicsConv
from ._base_block import BlockFactory


class LinearConvDownBlock(PhysicsConv):
    def __init__(self, num_spatial_dims, in_channels, out_channels, *, kernel_size=3, factor=2, use_bias=True,
                 zero_bias_init=False, boundary_mode, key):
        super().__init__(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels,
                         kernel_size=kernel_size, stride=factor, dilation=1, boundary_mode=boundary_mode,
                         use_bias=use_bias, zero_bias_init=zero_bias_init, key=key)


class LinearConvDownBlockFactory(BlockFactory):
    def __init__(self, *, kernel_size=3, factor=2, use_bias=True):
        self.kernel_size = kernel_size
        self.factor = factor
        self.use_bias = use_bias

    def __call__(self, num_spatial_dims, in_channels, out_channels, *, activation, boundary_mode, key):
        return LinearConvDownBlock(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels,
                                   kernel_size=self.kernel_size, factor=self.factor, boundary_mode=boundary_mode,
                                   use_bias=self.use_bias, key=key)
