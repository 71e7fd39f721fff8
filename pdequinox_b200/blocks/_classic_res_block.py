"""ClassicResBlock (post-activation ResNet block; reference: pdequinox/blocks/_classic_res_block.py:10-188).

conv_1 -> norm_1 -> act -> conv_2 -> norm_2 -> (+ bypass_norm(bypass_conv(x))) -> act. Without
normalisation (the ClassicResNet default) bias, activation and the residual add are fused into the
two convolution epilogues.
"""
from .. import _lib, nn, random as jr
from ..conv import PhysicsConv, PointwiseLinearConv
from ..nn import Identity
from ._base_block import Block, BlockFactory


class ClassicResBlock(Block):
    _fields = ("conv_1", "norm_1", "conv_2", "norm_2", "bypass_conv", "bypass_norm", "activation")

    def __init__(self, num_spatial_dims, in_channels, out_channels, *, activation=nn.relu, kernel_size=3,
                 use_norm=False, num_groups=1, use_bias=True, zero_bias_init=False, boundary_mode, key):
        self.num_spatial_dims = num_spatial_dims

        def conv_constructor(i, o, b, k):
            return PhysicsConv(num_spatial_dims=num_spatial_dims, in_channels=i, out_channels=o,
                               kernel_size=kernel_size, stride=1, dilation=1, boundary_mode=boundary_mode,
                               use_bias=b, zero_bias_init=zero_bias_init, key=k)

        k_1, k_2, key = jr.split(key, 3)  # _classic_res_block.py:78
        self.conv_1 = conv_constructor(in_channels, out_channels, use_bias, k_1)
        self.norm_1 = nn.GroupNorm(groups=num_groups, channels=out_channels) if use_norm else Identity()
        self.conv_2 = conv_constructor(out_channels, out_channels, use_bias, k_2)
        self.norm_2 = nn.GroupNorm(groups=num_groups, channels=out_channels) if use_norm else Identity()
        if out_channels != in_channels:
            bypass_conv_key, _ = jr.split(key)
            self.bypass_conv = PointwiseLinearConv(num_spatial_dims=num_spatial_dims, in_channels=in_channels,
                                                   out_channels=out_channels, use_bias=False, key=bypass_conv_key)
            self.bypass_norm = nn.GroupNorm(groups=num_groups, channels=out_channels) if use_norm else Identity()
        else:
            self.bypass_conv = Identity()
            self.bypass_norm = Identity()
        self.activation = activation
        self.use_norm = use_norm

    @property
    def receptive_field(self):
        return tuple((a0 + b0, a1 + b1) for (a0, a1), (b0, b1) in
                     zip(self.conv_1.receptive_field, self.conv_2.receptive_field))

    def _fwd(self, x, tape):
        act = nn.activation_code(self.activation)
        if self.use_norm:
            return self._fwd_norm(x, tape, act)
        # epilogue fusion needs act'(pre) from the OUTPUT: exact for relu / identity
        fuse = act in (_lib.ACT_RELU, _lib.ACT_IDENTITY)
        if isinstance(self.bypass_conv, Identity):
            skip = x
        else:
            skip = self.bypass_conv.apply(x, tape, tag="skip")
        if fuse:
            h = self.conv_1.apply(x, tape, act=act, tag="h")
            y = self.conv_2.apply(h, tape, residual=skip, act=act, tag="y")
            if tape is not None:
                tape.push((x, h, None, y, None))
            return y
        p1 = self.conv_1.apply(x, tape, tag="p1")
        h = nn.apply_activation(p1, act, tape, self, "h")
        p2 = self.conv_2.apply(h, tape, residual=skip, tag="p2")
        y = nn.apply_activation(p2, act, tape, self, "y")
        if tape is not None:
            tape.push((x, h, p1, y, p2))
        return y

    def _bwd(self, gy, tape, need_gx=True, gy_premasked=False, mask_gx=False):
        """gy_premasked: `gy` already carries this block's output relu' (the block behind applied it while forming its input
        gradient, see mask_gx). mask_gx: this block's input is the relu output of the block in front: its relu' is applied by
        the kernel that forms the input gradient, which is then that block's pre-activation cotangent (no pass of its own)."""
        act = nn.activation_code(self.activation)
        if self.use_norm:
            return self._bwd_norm(gy, tape, need_gx, act)
        x, h, p1, y, p2 = tape.pop()
        g2 = gy if gy_premasked else nn.activation_bwd(p2 if p2 is not None else y, gy, act, tape, self, "g2")
        # relu' of the inner activation rides on the data-gradient kernel of conv_2 (relu'(p1) = [h > 0]); other activations
        # keep their elementwise pass
        inner_fused = act == _lib.ACT_RELU and p1 is None
        gh = self.conv_2.apply_bwd(h, g2, tape, True, tag="gh", relu_mask=h if inner_fused else None)
        g1 = gh if inner_fused else nn.activation_bwd(p1 if p1 is not None else h, gh, act, tape, self, "g1")
        mask = x if (mask_gx and need_gx) else None
        identity_skip = isinstance(self.bypass_conv, Identity)
        if identity_skip:  # cotangent of the skip joins inside the data-gradient pass
            return self.conv_1.apply_bwd(x, g1, tape, need_gx, tag="gx", add=g2, relu_mask=mask)
        gskip = self.bypass_conv.apply_bwd(x, g2, tape, need_gx, tag="gskip")
        return self.conv_1.apply_bwd(x, g1, tape, need_gx, tag="gx", add=gskip, relu_mask=mask)

    def relu_output_saved(self, record):
        """True if `record` (what this block's forward pass pushed) holds the block's relu OUTPUT, i.e. the block behind may
        apply this block's relu' through `mask_gx`."""
        return (not self.use_norm and nn.activation_code(self.activation) == _lib.ACT_RELU and isinstance(record, tuple)
                and len(record) == 5 and record[4] is None)

    # ---- with GroupNorm (unfused ordering) ----
    def _fwd_norm(self, x, tape, act):
        t = tape
        p1 = self.conv_1._fwd(x, t)
        n1 = self.norm_1._fwd(p1, t)
        h = nn.apply_activation(n1, act, t, self, "h")
        p2 = self.conv_2._fwd(h, t)
        n2 = self.norm_2._fwd(p2, t)
        if isinstance(self.bypass_conv, Identity):
            skip = x
        else:
            skip = self.bypass_norm._fwd(self.bypass_conv._fwd(x, t), t)
        s = nn.add(n2, skip, t, self, "s")
        y = nn.apply_activation(s, act, t, self, "y")
        if t is not None:
            t.push((n1, s))
        return y

    def _bwd_norm(self, gy, tape, need_gx, act):
        n1, s = tape.pop()
        gs = nn.activation_bwd(s, gy, act, tape, self, "gs")
        gskip = gs
        if not isinstance(self.bypass_conv, Identity):
            gskip = self.bypass_conv._bwd(self.bypass_norm._bwd(gs, tape), tape, need_gx)
        g = self.norm_2._bwd(gs, tape)
        g = self.conv_2._bwd(g, tape)
        g = nn.activation_bwd(n1, g, act, tape, self, "gn1")
        g = self.norm_1._bwd(g, tape)
        gx = self.conv_1._bwd(g, tape, need_gx)
        return nn.add(gx, gskip, tape, self, "gx") if need_gx else None


class ClassicResBlockFactory(BlockFactory):
    def __init__(self, kernel_size=3, *, use_norm=False, num_groups=1, use_bias=True, zero_bias_init=False):
        self.kernel_size = kernel_size
        self.use_norm = use_norm
        self.num_groups = num_groups
        self.use_bias = use_bias
        self.zero_bias_init = zero_bias_init

    def __call__(self, num_spatial_dims, in_channels, out_channels, *, activation, boundary_mode, key):
        return ClassicResBlock(num_spatial_dims, in_channels, out_channels, activation=activation,
                               kernel_size=self.kernel_size, use_norm=self.use_norm, num_groups=self.num_groups,
                               boundary_mode=boundary_mode, key=key, use_bias=self.use_bias,
                               zero_bias_init=self.zero_bias_init)
