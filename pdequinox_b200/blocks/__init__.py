from ._base_block import Block, BlockFactory
from ._classic_res_block import ClassicResBlock, ClassicResBlockFactory
from ._classic_spectral_block import ClassicSpectralBlock, ClassicSpectralBlockFactory
from ._dilated_res_block import DilatedResBlock, DilatedResBlockFactory
from ._linear_channel_adjust_block import LinearChannelAdjustBlock, LinearChannelAdjustBlockFactory

__all__ = ["Block", "BlockFactory", "ClassicResBlock", "ClassicResBlockFactory", "ClassicSpectralBlock",
           "ClassicSpectralBlockFactory", "DilatedResBlock", "DilatedResBlockFactory", "LinearChannelAdjustBlock",
           "LinearChannelAdjustBlockFactory"]
