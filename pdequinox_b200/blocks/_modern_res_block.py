"""ModernResBlock (pre-activation ResNet block; reference: pdequinox/blocks/_modern_res_block.py:20-140).

norm_1 -> act -> conv_1 -> norm_2 -> act -> conv_2, plus bypass_conv(bypass_norm(x)) when the channel counts differ
(else the identity). The pre-activation ordering leaves nothing for the convolution epilogues to fuse except the bias
and the final residual add; every stage is one of the path's kernels (GroupNorm, activation, PhysicsConv, 1x1 conv).
"""
from .. import nn, random as jr
from ..conv import PhysicsConv, PointwiseLinearConv
from ..nn import Identity
from ._base_block import Block, BlockFactory


class ModernResBlock(Block):
    _fields = ("conv_1", "norm_1", "conv_2", "norm_2", "bypass_conv", "bypass_norm", "activation")

    def __init__(self, num_spatial_dims, in_channels, out_channels, *, activation=nn.relu, kernel_size=3, use_norm=True,
                 num_groups=1, use_bias=True, zero_bias_init=False, boundary_mode, key):
        self.num_spatial_dims = num_spatial_dims

        def conv_constructor(i, o, b, k):
            return PhysicsConv(num_spatial_dims=num_spatial_dims, in_channels=i, out_channels=o, kernel_size=kernel_size,
                               stride=1, dilation=1, boundary_mode=boundary_mode, use_bias=b,
                               zero_bias_init=zero_bias_init, key=k)

        k_1, k_2, key = jr.split(key, 3)  # _modern_res_block.py:88
        self.norm_1 = nn.GroupNorm(groups=num_groups, channels=in_channels) if use_norm else Identity()
        self.conv_1 = conv_constructor(in_channels, out_channels, use_bias, k_1)
        self.norm_2 = nn.GroupNorm(groups=num_groups, channels=out_channels) if use_norm else Identity()
        self.conv_2 = conv_constructor(out_channels, out_channels, use_bias, k_2)
        self.activation = activation
        if out_channels != in_channels:
            bypass_conv_key, _ = jr.split(key)
            # the bypass norm sits IN FRONT of the 1x1 conv here: in_channels (_modern_res_block.py:113-117)
            self.bypass_norm = nn.GroupNorm(groups=num_groups, channels=in_channels) if use_norm else Identity()
            self.bypass_conv = PointwiseLinearConv(num_spatial_dims=num_spatial_dims, in_channels=in_channels,
                                                   out_channels=out_channels, use_bias=False, key=bypass_conv_key)
        else:
            self.bypass_norm = Identity()
            self.bypass_conv = Identity()

    @property
    def receptive_field(self):
        return tuple((a0 + b0, a1 + b1) for (a0, a1), (b0, b1) in
                     zip(self.conv_1.receptive_field, self.conv_2.receptive_field))

    def _fwd(self, x, tape):
        act = nn.activation_code(self.activation)
        t = tape
        n1 = self.norm_1._fwd(x, t)
        p1 = self.conv_1._fwd(nn.apply_activation(n1, act, t, self, "a1"), t)
        n2 = self.norm_2._fwd(p1, t)
        a2 = nn.apply_activation(n2, act, t, self, "a2")
        if isinstance(self.bypass_conv, Identity):
            skip = x
        else:
            skip = self.bypass_conv._fwd(self.bypass_norm._fwd(x, t), t)
        # the residual join rides in conv_2's epilogue (PhysicsConv.apply(residual=...)); its record is pushed last
        y = self.conv_2.apply(a2, t, residual=skip, tag="y")
        if t is not None:
            t.push((n1, n2, a2))
        return y

    def _bwd(self, gy, tape, need_gx=True):
        act = nn.activation_code(self.activation)
        n1, n2, a2 = tape.pop()
        g = self.conv_2.apply_bwd(a2, gy, tape, True, tag="ga2")
        gskip = gy  # the residual join passes the cotangent on unchanged
        if not isinstance(self.bypass_conv, Identity):
            gskip = self.bypass_norm._bwd(self.bypass_conv._bwd(gy, tape, True), tape)
        g = nn.activation_bwd(n2, g, act, tape, self, "gn2")
        g = self.norm_2._bwd(g, tape)
        g = self.conv_1._bwd(g, tape, True)
        g = nn.activation_bwd(n1, g, act, tape, self, "gn1")
        g = self.norm_1._bwd(g, tape)
        return nn.add(g, gskip, tape, self, "gx")


class ModernResBlockFactory(BlockFactory):
    def __init__(self, kernel_size=3, *, use_norm=True, num_groups=1, use_bias=True, zero_bias_init=False):
        self.kernel_size = kernel_size
        self.use_norm = use_norm
        self.num_groups = num_groups
        self.use_bias = use_bias
        self.zero_bias_init = zero_bias_init

    def __call__(self, num_spatial_dims, in_channels, out_channels, *, activation, boundary_mode, key):
        return ModernResBlock(num_spatial_dims, in_channels, out_channels, activation=activation,
                              kernel_size=self.kernel_size, use_norm=self.use_norm, num_groups=self.num_groups,
                              boundary_mode=boundary_mode, key=key, use_bias=self.use_bias,
                              zero_bias_init=self.zero_bias_init)
