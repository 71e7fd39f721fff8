"""Architectures (reference package: pdequinox/arch): Sequential- and Hierarchical-based networks plus the plain ConvNet."""
import importlib

_ARCHS = {"ClassicFNO": "_classic_fno", "ClassicResNet": "_classic_res_net", "ClassicUNet": "_classic_u_net",
          "ConvNet": "_conv_net", "DilatedResNet": "_dilated_res_net", "ModernResNet": "_modern_res_net",
          "ModernUNet": "_modern_u_net"}
__all__ = sorted(_ARCHS)
for _name, _module in _ARCHS.items():
    globals()[_name] = getattr(importlib.import_module(f"{__name__}.{_module}"), _name)
del _name, _module
