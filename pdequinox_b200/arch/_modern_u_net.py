"""ModernUNet (reference: pdequinox/arch/_modern_u_net.py:15-99): a Hierarchical whose lifting and arcs are pre-activation
ResNet blocks (`num_blocks` per level and arc), stride-2 conv down-sampling, transposed-conv up-sampling with a concat
skip, 1x1 projection."""
from .. import nn
from .._hierarchical import Hierarchical
from ..blocks import (LinearChannelAdjustBlockFactory, LinearConvDownBlockFactory, LinearConvUpBlockFactory,
                      ModernResBlockFactory)


class ModernUNet(Hierarchical):
    def __init__(self, num_spatial_dims, in_channels, out_channels, *, hidden_channels=16, num_levels=4, num_blocks=2,
                 channel_multipliers=None, use_norm=True, activation=nn.relu, key, boundary_mode="periodic"):
        super().__init__(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels,
                         hidden_channels=hidden_channels, num_levels=num_levels, num_blocks=num_blocks,
                         channel_multipliers=channel_multipliers, activation=activation, key=key,
                         boundary_mode=boundary_mode,
                         lifting_factory=ModernResBlockFactory(use_norm=use_norm),
                         down_sampling_factory=LinearConvDownBlockFactory(),
                         left_arc_factory=ModernResBlockFactory(use_norm=use_norm),
                         up_sampling_factory=LinearConvUpBlockFactory(),
                         right_arc_factory=ModernResBlockFactory(use_norm=use_norm),
                         projection_factory=LinearChannelAdjustBlockFactory())
