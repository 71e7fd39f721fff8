"""ModernResNet (reference: pdequinox/arch/_modern_res_net.py:10-66): pointwise lifting, pre-activation residual blocks,
pointwise projection."""
from .. import nn
from .._sequential import Sequential
from ..blocks import LinearChannelAdjustBlockFactory, ModernResBlockFactory


class ModernResNet(Sequential):
    def __init__(self, num_spatial_dims, in_channels, out_channels, *, hidden_channels=32, num_blocks=6, use_norm=True,
                 activation=nn.relu, boundary_mode="periodic", key):
        super().__init__(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels,
                         hidden_channels=hidden_channels, num_blocks=num_blocks, activation=activation, key=key,
                         boundary_mode=boundary_mode,
                         lifting_factory=LinearChannelAdjustBlockFactory(use_bias=True, zero_bias_init=False),
                         block_factory=ModernResBlockFactory(use_norm=use_norm, use_bias=True, zero_bias_init=False),
                         projection_factory=LinearChannelAdjustBlockFactory(use_bias=True, zero_bias_init=False))
