"""Sequential: lifting -> N blocks -> projection (reference: pdequinox/_sequential.py:16-128)."""
from . import random as jr
from ._module import Module
from ._utils import sum_receptive_fields
from .blocks import ClassicResBlockFactory, LinearChannelAdjustBlockFactory


class Sequential(Module):
    _fields = ("lifting", "blocks", "projection")

    def __init__(self, num_spatial_dims, in_channels, out_channels, *, hidden_channels, num_blocks, activation, key,
                 boundary_mode, lifting_factory=None, block_factory=None, projection_factory=None):
        lifting_factory = lifting_factory or LinearChannelAdjustBlockFactory()
        block_factory = block_factory or ClassicResBlockFactory()
        projection_factory = projection_factory or LinearChannelAdjustBlockFactory()
        self.num_spatial_dims = num_spatial_dims
        subkey, key = jr.split(key)  # _sequential.py:66
        if num_blocks < 1:
            raise ValueError("num_blocks must be at least 1")
        if isinstance(hidden_channels, int):
            hidden_channels = (hidden_channels,) * (num_blocks + 1)
        elif len(hidden_channels) != (num_blocks + 1):
            raise ValueError("The list of hidden channels must be one longer than the number of blocks")
        self.lifting = lifting_factory(num_spatial_dims=num_spatial_dims, in_channels=in_channels,
                                       out_channels=hidden_channels[0], activation=activation,
                                       boundary_mode=boundary_mode, key=subkey)
        self.blocks = []
        for fan_in, fan_out in zip(hidden_channels[:-1], hidden_channels[1:]):
            subkey, key = jr.split(key)  # _sequential.py:91
            self.blocks.append(block_factory(num_spatial_dims=num_spatial_dims, in_channels=fan_in,
                                             out_channels=fan_out, activation=activation,
                                             boundary_mode=boundary_mode, key=subkey))
        self.projection = projection_factory(num_spatial_dims=num_spatial_dims, in_channels=hidden_channels[-1],
                                             out_channels=out_channels, activation=activation,
                                             boundary_mode=boundary_mode, key=key)

    def _fwd(self, x, tape):
        x = self.lifting._fwd(x, tape)
        for block in self.blocks:
            x = block._fwd(x, tape)
        return self.projection._fwd(x, tape)

    def _bwd(self, gy, tape, need_gx=False):
        # `_grad_ready_hook(prefix)` (set by the data-parallel train step) is told as soon as every gradient under a
        # pytree prefix is final, so that its all-reduce can start while the rest of the reverse pass still runs
        hook = getattr(self, "_grad_ready_hook", None) or (lambda prefix: None)
        g = self.projection._bwd(gy, tape, True)
        hook("projection.")
        for i in reversed(range(len(self.blocks))):
            g = self.blocks[i]._bwd(g, tape, True)
            hook(f"blocks.{i}.")
        gx = self.lifting._bwd(g, tape, need_gx)
        hook("lifting.")
        return gx

    @property
    def receptive_field(self):
        fields = (self.lifting.receptive_field,) + tuple(b.receptive_field for b in self.blocks) + \
            (self.projection.receptive_field,)
        return sum_receptive_fields(fields)
