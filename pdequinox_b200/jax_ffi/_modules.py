"""The reference's modules with `__call__` rebound to the B200 kernels. They SUBCLASS the reference's own classes, so the
constructors, field names, static fields, RNG split order and pytree layout are the reference's by construction
(pdequinox/conv/_spectral_conv.py:16-63, _conv.py:43-167, _physics_conv.py:29-118, _pointwise_linear_conv.py:6-53,
blocks/_classic_spectral_block.py:10-75). `convert(model)` swaps every such module inside an existing model."""
from __future__ import annotations

import dataclasses

import jax
import pdequinox as _ref

from . import _ops

_ACT_NAMES = {jax.nn.relu: "relu", jax.nn.gelu: "gelu"}


class SpectralConv(_ref.conv.SpectralConv):
    def __call__(self, x):  # replaces _spectral_conv.py:65-66 (rfftn -> corner einsum -> irfftn)
        return _ops.spectral_block(x, self.weights_real, self.weights_imag)


class PointwiseLinearConv(_ref.conv.PointwiseLinearConv):
    def __call__(self, x, *, key=None):  # replaces eqx.nn.Conv.__call__ for kernel_size = 1
        if x.ndim != self.num_spatial_dims + 1:
            raise ValueError(f"Input to `Conv` needs to have rank {self.num_spatial_dims + 1}, but input has shape {x.shape}.")
        return _ops.pointwise_conv(x, self.weight, self.bias)


class PhysicsConv(_ref.conv.PhysicsConv):
    def __call__(self, x, *, key=None):  # replaces _conv.py:169-219 (jnp.pad + lax.conv_general_dilated + bias)
        if x.ndim != self.num_spatial_dims + 1:  # _conv.py:183-188
            raise ValueError(f"Input to `Conv` needs to have rank {self.num_spatial_dims + 1}, but input has shape {x.shape}.")
        return _ops.physics_conv(x, self.weight, self.bias, stride=self.stride, dilation=self.dilation, padding=self.padding,
                                 padding_mode=self.padding_mode, groups=self.groups)


class ClassicSpectralBlock(_ref.blocks.ClassicSpectralBlock):
    def __call__(self, x):  # replaces _classic_spectral_block.py:72-75: ONE fused op instead of three
        act = _ACT_NAMES.get(self.activation)
        if act is None:  # an activation the epilogue does not know: keep the reference's composition around the fused linear part
            return self.activation(_ops.spectral_block(x, self.spectral_conv.weights_real, self.spectral_conv.weights_imag,
                                                       self.by_pass_conv.weight, self.by_pass_conv.bias))
        return _ops.spectral_block(x, self.spectral_conv.weights_real, self.spectral_conv.weights_imag, self.by_pass_conv.weight,
                                   self.by_pass_conv.bias, activation=act)


_TWINS = ((_ref.blocks.ClassicSpectralBlock, ClassicSpectralBlock), (_ref.conv.SpectralConv, SpectralConv),
          (_ref.conv.PhysicsConv, PhysicsConv), (_ref.conv.PointwiseLinearConv, PointwiseLinearConv))


def _rebind(module):
    for ref_cls, twin in _TWINS:
        if type(module) is ref_cls:
            new = object.__new__(twin)
            for f in dataclasses.fields(module):
                object.__setattr__(new, f.name, getattr(module, f.name))
            return new
    return module


def convert(model):
    """The same model (same leaves, same pytree structure up to the class of four module types) with every
    ClassicSpectralBlock, SpectralConv, PhysicsConv and PointwiseLinearConv replaced by its B200 twin. A
    ClassicSpectralBlock becomes ONE fused op (its two children only lend their weights); convolutions nested inside
    ResNet / UNet blocks are rebound individually."""
    ref_types = tuple(r for r, _ in _TWINS)
    return jax.tree_util.tree_map(_rebind, model, is_leaf=lambda m: isinstance(m, ref_types))
