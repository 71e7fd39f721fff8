"""ClassicDoubleConvBlock: conv -> norm -> act -> conv -> norm -> act
(reference: pdequinox/blocks/_classic_double_conv_block.py:11-172). Without normalisation and with relu / identity the
bias and activation run in the convolution epilogues."""
from .. import _lib, nn, random as jr
from ..conv import PhysicsConv
from ..nn import Identity
from ._base_block import Block, BlockFactory


class ClassicDoubleConvBlock(Block):
    _fields = ("conv_1", "norm_1", "conv_2", "norm_2", "activation")

    def __init__(self, num_spatial_dims, in_channels, out_channels, *, activation=nn.relu, kernel_size=3, use_norm=True,
                 num_groups=1, use_bias=True, zero_bias_init=False, boundary_mode, key):
        self.num_spatial_dims = num_spatial_dims

        def conv_constructor(i, o, b, k):
            return PhysicsConv(num_spatial_dims=num_spatial_dims, in_channels=i, out_channels=o, kernel_size=kernel_size,
                               stride=1, dilation=1, boundary_mode=boundary_mode, use_bias=b,
                               zero_bias_init=zero_bias_init, key=k)

        k_1, k_2 = jr.split(key)  # _classic_double_conv_block.py:73
        self.conv_1 = conv_constructor(in_channels, out_channels, use_bias, k_1)
        self.norm_1 = nn.GroupNorm(groups=num_groups, channels=out_channels) if use_norm else Identity()
        self.conv_2 = conv_constructor(out_channels, out_channels, use_bias, k_2)
        self.norm_2 = nn.GroupNorm(groups=num_groups, channels=out_channels) if use_norm else Identity()
        self.activation = activation
        self.use_norm = use_norm

    @property
    def receptive_field(self):
        return tuple((a0 + b0, a1 + b1) for (a0, a1), (b0, b1) in
                     zip(self.conv_1.receptive_field, self.conv_2.receptive_field))

    def _fwd(self, x, tape):  # _classic_double_conv_block.py:87-90
        act = nn.activation_code(self.activation)
        saved = []
        h = x
        fuse = (not self.use_norm) and act in (_lib.ACT_RELU, _lib.ACT_IDENTITY)
        for conv, norm, tag in ((self.conv_1, self.norm_1, "1"), (self.conv_2, self.norm_2, "2")):
            if fuse:
                xin = h
                h = conv.apply(xin, tape, act=act, tag="y")
                saved.append((xin, h))
            else:
                p = conv._fwd(h, tape)
                n = norm._fwd(p, tape)
                h = nn.apply_activation(n, act, tape, self, "a" + tag)
                saved.append((None, n))
        if tape is not None:
            tape.push((fuse, saved))
        return h

    def _bwd(self, gy, tape, need_gx=True):
        act = nn.activation_code(self.activation)
        fuse, saved = tape.pop()
        g = gy
        for idx, (conv, norm) in ((1, (self.conv_2, self.norm_2)), (0, (self.conv_1, self.norm_1))):
            xin, pre = saved[idx]
            g = nn.activation_bwd(pre, g, act, tape, self, f"g{idx}")
            need = need_gx or idx == 1
            if fuse:
                g = conv.apply_bwd(xin, g, tape, need, tag="gx")
            else:
                g = norm._bwd(g, tape)
                g = conv._bwd(g, tape, need)
        return g


class ClassicDoubleConvBlockFactory(BlockFactory):
    def __init__(self, *, kernel_size=3, use_norm=True, num_groups=1, use_bias=True, zero_bias_init=False):
        self.kernel_size = kernel_size
        self.use_norm = use_norm
        self.num_groups = num_groups
        self.use_bias = use_bias
        self.zero_bias_init = zero_bias_init

    def __call__(self, num_spatial_dims, in_channels, out_channels, *, activation, boundary_mode, key):
        return ClassicDoubleConvBlock(num_spatial_dims=num_spatial_dims, in_channels=in_channels,
                                      out_channels=out_channels, activation=activation, boundary_mode=boundary_mode,
                                      key=key, kernel_size=self.kernel_size, use_norm=self.use_norm,
                                      num_groups=self.num_groups, use_bias=self.use_bias,
                                      zero_bias_init=self.zero_bias_init)
