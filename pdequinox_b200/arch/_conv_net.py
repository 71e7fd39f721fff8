"""ConvNet (reference: pdequinox/arch/_conv_net.py:16-121): depth + 1 PhysicsConv layers, activation after all but the
last, `final_activation` after the last. relu / identity ride in the convolution epilogues."""
from .. import _lib, nn, random as jr
from .._module import Module
from .._utils import sum_receptive_fields
from ..conv import PhysicsConv


class ConvNet(Module):
    _fields = ("layers", "activation", "final_activation")

    def __init__(self, num_spatial_dims, in_channels, out_channels, *, hidden_channels=16, depth=10, activation=nn.relu,
                 kernel_size=3, final_activation=nn.identity, use_bias=True, use_final_bias=True, boundary_mode="periodic",
                 key, zero_bias_init=False):
        self.num_spatial_dims = num_spatial_dims
        self.activation = activation
        self.final_activation = final_activation
        keys = jr.split(key, depth + 1)  # _conv_net.py:76

        def conv_constructor(i, o, b, k):
            return PhysicsConv(num_spatial_dims=num_spatial_dims, in_channels=i, out_channels=o, kernel_size=kernel_size,
                               stride=1, dilation=1, use_bias=b, zero_bias_init=zero_bias_init,
                               boundary_mode=boundary_mode, key=k)

        if depth == 0:
            layers = [conv_constructor(in_channels, out_channels, use_final_bias, keys[0])]
        else:
            layers = [conv_constructor(in_channels, hidden_channels, use_bias, keys[0])]
            layers += [conv_constructor(hidden_channels, hidden_channels, use_bias, keys[i + 1]) for i in range(depth - 1)]
            layers.append(conv_constructor(hidden_channels, out_channels, use_final_bias, keys[-1]))
        self.layers = tuple(layers)

    @property
    def receptive_field(self):
        return sum_receptive_fields(tuple(conv.receptive_field for conv in self.layers))

    def _acts(self):
        inner, final = nn.activation_code(self.activation), nn.activation_code(self.final_activation)
        return [inner] * (len(self.layers) - 1) + [final]

    def _fwd(self, x, tape):
        saved = []
        for j, (conv, act) in enumerate(zip(self.layers, self._acts())):
            if act in (_lib.ACT_RELU, _lib.ACT_IDENTITY):  # act'(pre) is recoverable from the output: fused epilogue
                out, pre = conv.apply(x, tape, act=act, tag="a"), None
            else:
                pre = conv.apply(x, tape, tag="p")
                out = nn.apply_activation(pre, act, tape, conv, "a")
            saved.append((x, out, pre))
            x = out
        if tape is not None:
            tape.push(saved)
        return x

    def _bwd(self, gy, tape, need_gx=True):
        saved = tape.pop()
        g = gy
        for j in reversed(range(len(self.layers))):
            x, out, pre = saved[j]
            g = nn.activation_bwd(pre if pre is not None else out, g, self._acts()[j], tape, self.layers[j], "ga")
            g = self.layers[j].apply_bwd(x, g, tape, need_gx or j > 0, tag="gx")
        return g
