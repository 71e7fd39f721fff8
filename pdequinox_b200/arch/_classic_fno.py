"""ClassicFNO (reference: pdequinox/arch/_classic_fno.py:10-84)."""
from .. import nn
from .._sequential import Sequential
from ..blocks import ClassicSpectralBlockFactory, LinearChannelAdjustBlockFactory


class ClassicFNO(Sequential):
    def __init__(self, num_spatial_dims, in_channels, out_channels, *, hidden_channels=32, num_modes=12,
                 num_blocks=4, activation=nn.gelu, boundary_mode=None, key):
        super().__init__(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels,
                         hidden_channels=hidden_channels, num_blocks=num_blocks, activation=activation, key=key,
                         boundary_mode="periodic",  # does not matter (_classic_fno.py:70)
                         lifting_factory=LinearChannelAdjustBlockFactory(use_bias=True, zero_bias_init=False),
                         block_factory=ClassicSpectralBlockFactory(num_modes=num_modes, use_bias=True,
                                                                   zero_bias_init=False),
                         projection_factory=LinearChannelAdjustBlockFactory(use_bias=True, zero_bias_init=False))
