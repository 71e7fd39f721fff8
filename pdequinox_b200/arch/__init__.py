from ._classic_fno import ClassicFNO
from ._classic_res_net import ClassicResNet
from ._classic_u_net import ClassicUNet
from ._dilated_res_net import DilatedResNet

__all__ = ["ClassicFNO", "ClassicResNet", "ClassicUNet", "DilatedResNet"]
