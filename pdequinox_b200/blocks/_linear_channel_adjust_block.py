"""LinearChannelAdjustBlock: the 1x1 lifting / projection of every architecture
(reference: pdequinox/blocks/_linear_channel_adjust_block.py:9-66)."""
from ..conv import PointwiseLinearConv
from ._base_block import BlockFactory


class LinearChannelAdjustBlock(PointwiseLinearConv):
    def __init__(self, num_spatial_dims, in_channels, out_channels, *, use_bias, zero_bias_init, key):
        super().__init__(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels,
                         use_bias=use_bias, zero_bias_init=zero_bias_init, key=key)


class LinearChannelAdjustBlockFactory(BlockFactory):
    def __init__(self, *, use_bias: bool = True, zero_bias_init: bool = False):
        self.use_bias = use_bias
        self.zero_bias_init = zero_bias_init

    def __call__(self, num_spatial_dims, in_channels, out_channels, *, activation, boundary_mode, key):
        return LinearChannelAdjustBlock(num_spatial_dims=num_spatial_dims, in_channels=in_channels,
                                        out_channels=out_channels, use_bias=self.use_bias,
                                        zero_bias_init=self.zero_bias_init, key=key)
