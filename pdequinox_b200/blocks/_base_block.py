"""Block / BlockFactory base classes (reference: pdequinox/blocks/_base_block.py:8-42)."""
from .._module import Module


class Block(Module):
    pass


class BlockFactory:
    def __call__(self, num_spatial_dims, in_channels, out_channels, *, activation, boundary_mode, key):
        raise NotImplementedError
