"""ClassicUNet (reference: pdequinox/arch/_classic_u_net.py:15-84): double-conv lifting, stride-2 conv down-sampling
(not pooling), double-conv arcs, transposed-conv up-sampling with a concat skip, 1x1 projection."""
from .. import nn
from .._hierarchical import Hierarchical
from ..blocks import (ClassicDoubleConvBlockFactory, LinearChannelAdjustBlockFactory, LinearConvDownBlockFactory,
                      LinearConvUpBlockFactory)


class ClassicUNet(Hierarchical):
    def __init__(self, num_spatial_dims, in_channels, out_channels, *, hidden_channels=16, num_levels=4, use_norm=True,
                 activation=nn.relu, key, boundary_mode="periodic"):
        super().__init__(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels,
                         hidden_channels=hidden_channels, num_levels=num_levels, num_blocks=1, activation=activation,
                         key=key, boundary_mode=boundary_mode,
                         lifting_factory=ClassicDoubleConvBlockFactory(use_norm=use_norm),
                         down_sampling_factory=LinearConvDownBlockFactory(),
                         left_arc_factory=ClassicDoubleConvBlockFactory(use_norm=use_norm),
                         up_sampling_factory=LinearConvUpBlockFactory(),
                         right_arc_factory=ClassicDoubleConvBlockFactory(use_norm=use_norm),
                         projection_factory=LinearChannelAdjustBlockFactory())
