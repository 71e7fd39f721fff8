"""pdequinox_b200: B200-native (sm_100a) implementation of pdequinox's hot path -- SpectralConv,
PhysicsConv / PointwiseLinearConv, the FNO / ResNet blocks and architectures built from them, and a
batch-sharded Adam training step. Same constructors, call convention and parameter pytree as the
reference; all arithmetic runs in libpdeqx_b200.so (hand-written CUDA). There is no CPU fallback."""
from . import arch, blocks, conv, nn, random, sample_data, train  # noqa: F401
from ._module import (count_parameters, get_precision, set_precision, tree_leaves, tree_paths, vmap,  # noqa: F401
                      Module)
from ._hierarchical import Hierarchical  # noqa: F401
from ._sequential import Sequential  # noqa: F401
from ._utils import cycling_dataloader, dataloader, sum_receptive_fields  # noqa: F401
from .arch import (ClassicFNO, ClassicResNet, ClassicUNet, ConvNet, DilatedResNet, ModernResNet,  # noqa: F401
                   ModernUNet)
from .blocks import *  # noqa: F401,F403
from .conv import (MorePaddingConv, MorePaddingConvTranspose, PhysicsConv, PhysicsConvTranspose,  # noqa: F401
                   PointwiseLinearConv, SpectralConv)

__version__ = "0.1.0"
