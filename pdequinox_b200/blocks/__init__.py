"""Blocks and their factories (reference package: pdequinox/blocks). Every block class `X` comes with `XFactory`."""
from . import (_base_block, _classic_double_conv_block, _classic_res_block, _classic_spectral_block, _dilated_res_block,
               _linear_blocks, _modern_res_block)

_BLOCKS = {
    _classic_double_conv_block: ["ClassicDoubleConvBlock"],
    _classic_res_block: ["ClassicResBlock"],
    _classic_spectral_block: ["ClassicSpectralBlock"],
    _dilated_res_block: ["DilatedResBlock"],
    _linear_blocks: ["LinearChannelAdjustBlock", "LinearConvBlock", "LinearConvDownBlock", "LinearConvUpBlock"],
    _modern_res_block: ["ModernResBlock"],
}
__all__ = ["Block", "BlockFactory"]
Block, BlockFactory = _base_block.Block, _base_block.BlockFactory
for _module, _names in _BLOCKS.items():
    for _name in _names:
        for _export in (_name, _name + "Factory"):
            globals()[_export] = getattr(_module, _export)
            __all__.append(_export)
del _module, _names, _name, _export
