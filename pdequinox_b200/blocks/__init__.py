from ._base_block import Block, BlockFactory
from ._classic_double_conv_block import ClassicDoubleConvBlock, ClassicDoubleConvBlockFactory
from ._classic_res_block import ClassicResBlock, ClassicResBlockFactory
from ._classic_spectral_block import ClassicSpectralBlock, ClassicSpectralBlockFactory
from ._dilated_res_block import DilatedResBlock, DilatedResBlockFactory
from ._linear_channel_adjust_block import LinearChannelAdjustBlock, LinearChannelAdjustBlockFactory
from ._linear_conv_block import LinearConvBlock, LinearConvBlockFactory
from ._linear_conv_down_block import LinearConvDownBlock, LinearConvDownBlockFactory
from ._linear_conv_up_block import LinearConvUpBlock, LinearConvUpBlockFactory
from ._modern_res_block import ModernResBlock, ModernResBlockFactory

__all__ = ["Block", "BlockFactory", "ClassicDoubleConvBlock", "ClassicDoubleConvBlockFactory", "ClassicResBlock",
           "ClassicResBlockFactory", "ClassicSpectralBlock", "ClassicSpectralBlockFactory", "DilatedResBlock",
           "DilatedResBlockFactory", "LinearChannelAdjustBlock", "LinearChannelAdjustBlockFactory",
           "LinearConvBlock", "LinearConvBlockFactory", "LinearConvDownBlock", "LinearConvDownBlockFactory", "LinearConvUpBlock", "LinearConvUpBlockFactory",
           "ModernResBlock", "ModernResBlockFactory"]
