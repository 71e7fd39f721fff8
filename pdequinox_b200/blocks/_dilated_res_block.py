"""DilatedResBlock: per dilation rate norm -> conv -> act, then + bypass
(reference: pdequinox/blocks/_dilated_res_block.py:18-158)."""
import os

from .. import _lib, nn, random as jr
from ..conv import PhysicsConv, PointwiseLinearConv
from ..nn import Identity
from ._base_block import Block, BlockFactory


class DilatedResBlock(Block):
    _fields = ("norm_layers", "conv_layers", "activation", "bypass_conv", "bypass_norm")

    def __init__(self, num_spatial_dims, in_channels, out_channels, *, activation=nn.relu, kernel_size=3,
                 dilation_rates=(1, 2, 4, 8, 4, 2, 1), use_norm=True, num_groups=1, use_bias=True,
                 zero_bias_init=False, boundary_mode, key):
        self.num_spatial_dims = num_spatial_dims

        def conv_constructor(i, o, d, b, k):
            return PhysicsConv(num_spatial_dims=num_spatial_dims, in_channels=i, out_channels=o,
                               kernel_size=kernel_size, stride=1, dilation=d, boundary_mode=boundary_mode,
                               use_bias=b, zero_bias_init=zero_bias_init, key=k)

        if use_norm:
            norm_layers = [nn.GroupNorm(groups=num_groups, channels=in_channels)]
            for _ in dilation_rates[1:]:
                norm_layers.append(nn.GroupNorm(groups=num_groups, channels=out_channels))
            self.norm_layers = tuple(norm_layers)
        else:
            self.norm_layers = tuple(Identity() for _ in dilation_rates)
        key, *keys = jr.split(key, len(dilation_rates) + 1)  # _dilated_res_block.py:103
        conv_layers = [conv_constructor(in_channels, out_channels, dilation_rates[0], use_bias, keys[0])]
        for d, k in zip(dilation_rates[1:], keys[1:]):
            conv_layers.append(conv_constructor(out_channels, out_channels, d, use_bias, k))
        self.conv_layers = tuple(conv_layers)
        self.activation = activation
        if out_channels != in_channels:
            self.bypass_norm = nn.GroupNorm(groups=num_groups, channels=in_channels) if use_norm else Identity()
            bypass_conv_key, _ = jr.split(key)
            self.bypass_conv = PointwiseLinearConv(num_spatial_dims=num_spatial_dims, in_channels=in_channels,
                                                   out_channels=out_channels, use_bias=use_bias,
                                                   zero_bias_init=zero_bias_init, key=bypass_conv_key)
        else:
            self.bypass_conv = Identity()
            self.bypass_norm = Identity()

    @property
    def receptive_field(self):
        from .._utils import sum_receptive_fields
        return sum_receptive_fields(tuple(c.receptive_field for c in self.conv_layers))

    # ---- GroupNorm folded into the convolutions (statistics in the producer's epilogue, normalise on
    # load): every norm is groups = 1, every conv a periodic tcgen05 CTA-pair shape, the skip is the identity ----
    def _gn_folded(self, x, act):
        if os.environ.get("PDEQX_GN_FOLD", "1") == "0":  # experiment knob: separate GroupNorm passes
            return False
        if act not in (_lib.ACT_RELU, _lib.ACT_IDENTITY) or not isinstance(self.bypass_conv, Identity):
            return False
        if not all(isinstance(n, nn.GroupNorm) and n.groups == 1 for n in self.norm_layers):
            return False
        return all(c.gn_fusable(x) for c in self.conv_layers)

    def _fwd_folded(self, x, tape, act):
        B, count = x.shape[0], x[0].numel()
        h, parts, slots = x, *getattr(x, "_pdeqx_parts", (None, 0))
        saved, stats_all = [], []
        last = len(self.conv_layers) - 1
        for j, (norm, conv) in enumerate(zip(self.norm_layers, self.conv_layers)):
            stats = norm.stats_of(h, tape) if parts is None else norm.stats_from_parts(parts, slots, B, count, tape)
            out, parts, slots = conv.apply_gn(h, norm, stats, tape, act=act, tag="a", want_parts=j < last)
            saved.append((h, out, None))
            stats_all.append(stats)
            h = out
        y, yparts, yslots = nn.add_parts(h, x, tape, self, "y")
        y._pdeqx_parts = (yparts, yslots)  # the next block's first norm reads its statistics from these
        if tape is not None:
            tape.push({"dilated_folded": stats_all})
            tape.push(saved)
        return y

    def _bwd_folded(self, gy, tape, saved, stats_all, act):
        last = len(self.conv_layers) - 1
        g = gy
        for j in reversed(range(len(self.conv_layers))):
            h, out, _ = saved[j]
            conv, norm = self.conv_layers[j], self.norm_layers[j]
            # below the last layer `g` already carries relu' of this layer's output (applied by the norm pass behind it)
            ga = nn.activation_bwd(out, g, act, tape, conv, "ga") if j == last else g
            v, parts, slots = conv.apply_bwd_gn(h, ga, norm, stats_all[j], tape)
            g = norm.bwd_fused(h, v, stats_all[j], parts, slots, tape, mask_h=(j > 0 and act == _lib.ACT_RELU),
                               add=gy if j == 0 else None)
        return g

    def _fwd(self, x, tape):
        act = nn.activation_code(self.activation)
        if self._gn_folded(x, act):
            return self._fwd_folded(x, tape, act)
        fuse = act in (_lib.ACT_RELU, _lib.ACT_IDENTITY)
        identity_skip = isinstance(self.bypass_conv, Identity)
        skip = x if identity_skip else self.bypass_conv._fwd(self.bypass_norm._fwd(x, tape), tape)
        h = x
        saved = []
        last = len(self.conv_layers) - 1
        for j, (norm, conv) in enumerate(zip(self.norm_layers, self.conv_layers)):
            hn = norm._fwd(h, tape)
            if fuse:
                out = conv.apply(hn, tape, act=act, tag="a")
                saved.append((hn, out, None))
            else:
                pre = conv.apply(hn, tape, tag="p")
                out = nn.apply_activation(pre, act, tape, conv, "a")
                saved.append((hn, out, pre))
            h = out
        y = nn.add(h, skip, tape, self, "y")
        if hasattr(y, "_pdeqx_parts"):  # (a persistent buffer that an earlier, folded pass annotated)
            del y._pdeqx_parts
        if tape is not None:
            tape.push({"dilated_folded": None})
            tape.push(saved)
        return y

    def _bwd(self, gy, tape, need_gx=True):
        act = nn.activation_code(self.activation)
        saved = tape.pop()
        stats_all = tape.pop()["dilated_folded"]
        if stats_all is not None:
            g = self._bwd_folded(gy, tape, saved, stats_all, act)
            return g if need_gx else None
        identity_skip = isinstance(self.bypass_conv, Identity)
        g = gy
        premasked = False  # True: `g` already carries relu' of layer j's output (applied by the GroupNorm reverse pass behind it)
        for j in reversed(range(len(self.conv_layers))):
            hn, out, pre = saved[j]
            conv, norm = self.conv_layers[j], self.norm_layers[j]
            ga = g if premasked else nn.activation_bwd(pre if pre is not None else out, g, act, tape, conv, "ga")
            premasked = False
            first = j == 0
            no_norm = isinstance(norm, Identity)
            # a GroupNorm in front of the conv has parameters of its own: their gradients need the conv's input gradient
            # even when the caller does not want the block's (and the norm's tape record must be popped either way)
            want = need_gx or not first or not no_norm
            add = gy if (first and identity_skip and no_norm) else None
            # without a norm in between, the conv's input IS the relu output of layer j-1: its relu' rides on this data gradient
            cmask = saved[j - 1][1] if (no_norm and j > 0 and act == _lib.ACT_RELU and saved[j - 1][2] is None) else None
            g = conv.apply_bwd(hn, ga, tape, want, tag="gx", add=add, relu_mask=cmask)
            premasked = cmask is not None
            if want and not no_norm:
                # the norm's input is the relu output of layer j-1: that layer's relu' rides on this norm's reverse pass
                mask = saved[j - 1][1] if (j > 0 and act == _lib.ACT_RELU and saved[j - 1][2] is None) else None
                g = norm._bwd(g, tape, relu_mask=mask)
                premasked = mask is not None
                if first and identity_skip:
                    g = nn.add(g, gy, tape, self, "gx")
        if not identity_skip:
            skip_norm = isinstance(self.bypass_norm, Identity)
            gs = self.bypass_conv._bwd(gy, tape, need_gx or not skip_norm)
            if not skip_norm:
                gs = self.bypass_norm._bwd(gs, tape)
            if need_gx:
                g = nn.add(g, gs, tape, self, "gx")
        return g if need_gx else None


class DilatedResBlockFactory(BlockFactory):
    def __init__(self, kernel_size=3, dilation_rates=(1, 2, 4, 8, 4, 2, 1), *, use_norm=True, num_groups=1,
                 use_bias=True, zero_bias_init=False):
        self.kernel_size = kernel_size
        self.dilation_rates = dilation_rates
        self.use_norm = use_norm
        self.num_groups = num_groups
        self.use_bias = use_bias
        self.zero_bias_init = zero_bias_init

    def __call__(self, num_spatial_dims, in_channels, out_channels, *, activation, boundary_mode, key):
        return DilatedResBlock(num_spatial_dims, in_channels, out_channels, activation=activation,
                               kernel_size=self.kernel_size, dilation_rates=self.dilation_rates,
                               boundary_mode=boundary_mode, key=key, use_norm=self.use_norm,
                               num_groups=self.num_groups, use_bias=self.use_bias,
                               zero_bias_init=self.zero_bias_init)
