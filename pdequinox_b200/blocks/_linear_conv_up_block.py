"""LinearConvUpBlock: strided PhysicsConvTranspose (reference: pdequinox/blocks/_linear_conv_up_block.py:9-96)."""
from ..conv import PhysicsConvTranspose
from ._base_block import BlockFactory


class LinearConvUpBlock(PhysicsConvTranspose):
    def __init__(self, num_spatial_dims, in_channels, out_channels, *, kernel_size=3, factor=2, output_padding=1,
                 use_bias=True, zero_bias_init=False, boundary_mode, key):
        super().__init__(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels,
                         kernel_size=kernel_size, stride=factor, output_padding=output_padding, dilation=1,
                         boundary_mode=boundary_mode, use_bias=use_bias, zero_bias_init=zero_bias_init, key=key)


class LinearConvUpBlockFactory(BlockFactory):
    def __init__(self, *, kernel_size=3, factor=2, use_bias=True, output_padding=1):
        self.kernel_size = kernel_size
        self.factor = factor
        self.use_bias = use_bias
        self.output_padding = output_padding

    def __call__(self, num_spatial_dims, in_channels, out_channels, *, activation, boundary_mode, key):
        return LinearConvUpBlock(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels,
                                 kernel_size=self.kernel_size, factor=self.factor, output_padding=self.output_padding,
                                 boundary_mode=boundary_mode, use_bias=self.use_bias, key=key)
