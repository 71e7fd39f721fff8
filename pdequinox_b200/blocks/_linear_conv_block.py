"""LinearConvBlock: a plain stride-1 PhysicsConv as a block (reference: pdequinox/blocks/_linear_conv_block.py:9-76)."""
from ..conv import PhysicsConv
from ._base_block import BlockFactory


class LinearConvBlock(PhysicsConv):
    def __init__(self, num_spatial_dims, in_channels, out_channels, *, kernel_size=3, use_bias=True, zero_bias_init=False,
                 boundary_mode, key):
        super().__init__(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels,
                         kernel_size=kernel_size, stride=1, dilation=1, boundary_mode=boundary_mode, use_bias=use_bias,
                         zero_bias_init=zero_bias_init, key=key)


class LinearConvBlockFactory(BlockFactory):
    def __init__(self, *, kernel_size=3, use_bias=True):
        self.kernel_size = kernel_size
        self.use_bias = use_bias

    def __call__(self, num_spatial_dims, in_channels, out_channels, *, activation, boundary_mode, key):
        # (`activation` is part of the factory protocol and unused: the block is linear)
        return LinearConvBlock(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels,
                               kernel_size=self.kernel_size, boundary_mode=boundary_mode, use_bias=self.use_bias, key=key)
