"""The four linear (activation-free) blocks, each one convolution with fixed structural arguments, and their factories:

* `LinearChannelAdjustBlock` -- 1x1 conv: the lifting / projection of every architecture
  (reference: pdequinox/blocks/_linear_channel_adjust_block.py:9-66)
* `LinearConvBlock`          -- stride-1 PhysicsConv (pdequinox/blocks/_linear_conv_block.py:9-76)
* `LinearConvDownBlock`      -- stride-`factor` PhysicsConv, the UNet's down-sampling (…/_linear_conv_down_block.py:9-84)
* `LinearConvUpBlock`        -- stride-`factor` PhysicsConvTranspose, its up-sampling (…/_linear_conv_up_block.py:9-96)

A factory stores its keyword defaults and forwards them together with the structural arguments of the `BlockFactory`
protocol; like the reference's factories they never forward `zero_bias_init` for the convolutional blocks
(_linear_conv_down_block.py:74-83) and ignore `activation`.
"""
from ..conv import PhysicsConv, PhysicsConvTranspose, PointwiseLinearConv
from ._base_block import BlockFactory


class LinearChannelAdjustBlock(PointwiseLinearConv):
    def __init__(self, num_spatial_dims, in_channels, out_channels, *, use_bias, zero_bias_init, key):
        PointwiseLinearConv.__init__(self, num_spatial_dims, in_channels, out_channels, use_bias,
                                     zero_bias_init=zero_bias_init, key=key)


class LinearConvBlock(PhysicsConv):
    def __init__(self, num_spatial_dims, in_channels, out_channels, *, kernel_size=3, use_bias=True, zero_bias_init=False,
                 boundary_mode, key):
        PhysicsConv.__init__(self, num_spatial_dims, in_channels, out_channels, kernel_size, stride=1, dilation=1,
                             use_bias=use_bias, zero_bias_init=zero_bias_init, boundary_mode=boundary_mode, key=key)


class LinearConvDownBlock(PhysicsConv):
    def __init__(self, num_spatial_dims, in_channels, out_channels, *, kernel_size=3, factor=2, use_bias=True,
                 zero_bias_init=False, boundary_mode, key):
        PhysicsConv.__init__(self, num_spatial_dims, in_channels, out_channels, kernel_size, stride=factor, dilation=1,
                             use_bias=use_bias, zero_bias_init=zero_bias_init, boundary_mode=boundary_mode, key=key)


class LinearConvUpBlock(PhysicsConvTranspose):
    def __init__(self, num_spatial_dims, in_channels, out_channels, *, kernel_size=3, factor=2, output_padding=1,
                 use_bias=True, zero_bias_init=False, boundary_mode, key):
        PhysicsConvTranspose.__init__(self, num_spatial_dims, in_channels, out_channels, kernel_size, stride=factor,
                                      output_padding=output_padding, dilation=1, use_bias=use_bias,
                                      zero_bias_init=zero_bias_init, boundary_mode=boundary_mode, key=key)


class _StoredKwargsFactory(BlockFactory):
    """Keeps the keyword arguments given at construction as attributes (the reference's factories are small dataclasses
    with exactly these fields) and builds `block_cls` from them."""
    block_cls = None
    defaults = {}
    needs_boundary = True

    def __init__(self, **kwargs):
        unknown = set(kwargs) - set(self.defaults)
        if unknown:
            raise TypeError(f"{type(self).__name__}() got unexpected keyword arguments {sorted(unknown)}")
        for name, value in {**self.defaults, **kwargs}.items():
            setattr(self, name, value)

    def __call__(self, num_spatial_dims, in_channels, out_channels, *, activation, boundary_mode, key):
        extra = {name: getattr(self, name) for name in self.defaults}
        if self.needs_boundary:
            extra["boundary_mode"] = boundary_mode
        return self.block_cls(num_spatial_dims=num_spatial_dims, in_channels=in_channels, out_channels=out_channels, key=key,
                              **extra)


class LinearChannelAdjustBlockFactory(_StoredKwargsFactory):
    block_cls, defaults, needs_boundary = LinearChannelAdjustBlock, {"use_bias": True, "zero_bias_init": False}, False


class LinearConvBlockFactory(_StoredKwargsFactory):
    block_cls, defaults = LinearConvBlock, {"kernel_size": 3, "use_bias": True}


class LinearConvDownBlockFactory(_StoredKwargsFactory):
    block_cls, defaults = LinearConvDownBlock, {"kernel_size": 3, "factor": 2, "use_bias": True}


class LinearConvUpBlockFactory(_StoredKwargsFactory):
    block_cls, defaults = LinearConvUpBlock, {"kernel_size": 3, "factor": 2, "use_bias": True, "output_padding": 1}
