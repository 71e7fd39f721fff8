"""Hierarchical: UNet-style encoder / decoder over several spatial resolutions
(reference: pdequinox/_hierarchical.py:20-366). Same constructor, field names, leaf order and RNG split order
(`:130`, `:175`; every block of one arc re-uses the same key, `:196,208,234,246`)."""
from . import _lib, nn, random as jr
from ._module import Module, new_buf
from ._utils import sum_receptive_fields
from .blocks import (ClassicDoubleConvBlockFactory, LinearChannelAdjustBlockFactory, LinearConvDownBlockFactory,
                     LinearConvUpBlockFactory)


class Hierarchical(Module):
    _fields = ("lifting", "down_sampling_blocks", "left_arc_blocks", "up_sampling_blocks", "right_arc_blocks",
               "projection")

    def __init__(self, num_spatial_dims, in_channels, out_channels, *, hidden_channels, num_levels, num_blocks,
                 activation, key, reduction_factor=2, boundary_mode, channel_multipliers=None, lifting_factory=None,
                 down_sampling_factory=None, left_arc_factory=None, up_sampling_factory=None, right_arc_factory=None,
                 projection_factory=None):
        lifting_factory = lifting_factory or ClassicDoubleConvBlockFactory()
        down_sampling_factory = down_sampling_factory or LinearConvDownBlockFactory()
        left_arc_factory = left_arc_factory or ClassicDoubleConvBlockFactory()
        up_sampling_factory = up_sampling_factory or LinearConvUpBlockFactory()
        right_arc_factory = right_arc_factory or ClassicDoubleConvBlockFactory()
        projection_factory = projection_factory or LinearChannelAdjustBlockFactory()
        self.num_spatial_dims = num_spatial_dims
        self.down_sampling_blocks, self.left_arc_blocks = [], []
        self.up_sampling_blocks, self.right_arc_blocks = [], []
        self.reduction_factor = reduction_factor
        self.num_levels = num_levels
        key, lifting_key, projection_key = jr.split(key, 3)  # _hierarchical.py:130
        common = dict(num_spatial_dims=num_spatial_dims, activation=activation, boundary_mode=boundary_mode)
        self.lifting = lifting_factory(in_channels=in_channels, out_channels=hidden_channels, key=lifting_key, **common)
        self.projection = projection_factory(in_channels=hidden_channels, out_channels=out_channels, key=projection_key,
                                             **common)
        if channel_multipliers is not None:
            if len(channel_multipliers) != num_levels:
                raise ValueError("len(channel_multipliers) must match num_levels")
        else:
            channel_multipliers = tuple(reduction_factor ** i for i in range(1, num_levels + 1))
        self.channel_multipliers = channel_multipliers
        channel_list = [hidden_channels] + [hidden_channels * m for m in channel_multipliers]
        if num_blocks < 1:
            raise ValueError("num_blocks must be at least 1")
        self.num_blocks = num_blocks
        for fan_in, fan_out in zip(channel_list[:-1], channel_list[1:]):
            key, down_key, left_key, up_key, right_key = jr.split(key, 5)  # _hierarchical.py:175
            self.down_sampling_blocks.append(down_sampling_factory(in_channels=fan_in, out_channels=fan_in, key=down_key,
                                                                   **common))
            left = [left_arc_factory(in_channels=fan_in, out_channels=fan_out, key=left_key, **common)]
            for _ in range(num_blocks - 1):
                left.append(left_arc_factory(in_channels=fan_out, out_channels=fan_out, key=left_key, **common))
            self.left_arc_blocks.append(left)
            self.up_sampling_blocks.append(up_sampling_factory(in_channels=fan_out, out_channels=fan_out // reduction_factor,
                                                               key=up_key, **common))
            right = [right_arc_factory(in_channels=fan_out // reduction_factor + fan_in, out_channels=fan_in,
                                       key=right_key, **common)]
            for _ in range(num_blocks - 1):
                right.append(right_arc_factory(in_channels=fan_in, out_channels=fan_in, key=right_key, **common))
            self.right_arc_blocks.append(right)

    # nested lists of blocks: the generic leaf walker only descends one list level
    def named_leaves(self, prefix=""):
        out = list(self.lifting.named_leaves(prefix + "lifting."))
        for i, blk in enumerate(self.down_sampling_blocks):
            out.extend(blk.named_leaves(f"{prefix}down_sampling_blocks.{i}."))
        for i, arc in enumerate(self.left_arc_blocks):
            for j, blk in enumerate(arc):
                out.extend(blk.named_leaves(f"{prefix}left_arc_blocks.{i}.{j}."))
        for i, blk in enumerate(self.up_sampling_blocks):
            out.extend(blk.named_leaves(f"{prefix}up_sampling_blocks.{i}."))
        for i, arc in enumerate(self.right_arc_blocks):
            for j, blk in enumerate(arc):
                out.extend(blk.named_leaves(f"{prefix}right_arc_blocks.{i}.{j}."))
        out.extend(self.projection.named_leaves(prefix + "projection."))
        return out

    def _fwd(self, x, tape):  # _hierarchical.py:251-284
        for dims in x.shape[2:]:
            if dims % self.reduction_factor ** self.num_levels != 0:
                raise ValueError("Spatial dim issue")
        L = _lib.lib()
        x = self.lifting._fwd(x, tape)
        skips = []
        for down, left_list in zip(self.down_sampling_blocks, self.left_arc_blocks):
            skips.append(x)
            x = down._fwd(x, tape)
            for left in left_list:
                x = left._fwd(x, tape)
        for level, up, right_list in zip(reversed(range(self.num_levels)), reversed(self.up_sampling_blocks),
                                         reversed(self.right_arc_blocks)):
            skip = skips.pop()
            x = up._fwd(x, tape)
            if skip.shape[0] != x.shape[0] or tuple(skip.shape[2:]) != tuple(x.shape[2:]):
                # the reference raises from jnp.concatenate (_hierarchical.py:272-275) when a custom up-sampling factory
                # does not restore the skip's grid
                raise ValueError(f"Cannot concatenate the skip connection {tuple(skip.shape)} with the up-sampled "
                                 f"tensor {tuple(x.shape)}: all dimensions but the channel axis must match")
            B, ca, cb = skip.shape[0], skip.shape[1], x.shape[1]
            n = skip[0, 0].numel()
            cat = new_buf(tape, self, f"cat{level}", (B, ca + cb) + tuple(skip.shape[2:]))
            _lib.check(L.pdeqx_concat_channels(B, ca, cb, n, _lib.ptr(skip), _lib.ptr(x), _lib.ptr(cat), _lib.stream()))
            x = cat
            if tape is not None:
                tape.push((ca, cb))
            for right in right_list:
                x = right._fwd(x, tape)
        return self.projection._fwd(x, tape)

    def _bwd(self, gy, tape, need_gx=False):
        L = _lib.lib()
        g = self.projection._bwd(gy, tape, True)
        gskips = []
        for level, up, right_list in zip(range(self.num_levels), self.up_sampling_blocks, self.right_arc_blocks):
            for right in reversed(right_list):
                g = right._bwd(g, tape, True)
            ca, cb = tape.pop()
            B = g.shape[0]
            n = g[0, 0].numel()
            gs = new_buf(tape, self, f"gskip{level}", (B, ca) + tuple(g.shape[2:]))
            gu = new_buf(tape, self, f"gup{level}", (B, cb) + tuple(g.shape[2:]))
            _lib.check(L.pdeqx_split_channels(B, ca, cb, n, _lib.ptr(g), _lib.ptr(gs), _lib.ptr(gu), _lib.stream()))
            gskips.append(gs)
            g = up._bwd(gu, tape, True)
        for level, down, left_list in zip(reversed(range(self.num_levels)), reversed(self.down_sampling_blocks),
                                          reversed(self.left_arc_blocks)):
            for left in reversed(left_list):
                g = left._bwd(g, tape, True)
            g = down._bwd(g, tape, True)
            g = nn.add(g, gskips[level], tape, self, f"gjoin{level}")
        return self.lifting._bwd(g, tape, need_gx)

    @property
    def receptive_field(self):  # _hierarchical.py:286-366
        red = tuple(self.reduction_factor ** level for level in range(self.num_levels + 1))

        def scaled(rf, r):
            return tuple((b * r, f * r) for b, f in rf)

        fields = [self.lifting.receptive_field]
        fields += [scaled(b.receptive_field, r) for b, r in zip(self.down_sampling_blocks, red[:-1])]
        fields += [scaled(b.receptive_field, r) for arc, r in zip(self.left_arc_blocks, red[1:]) for b in arc]
        fields += [scaled(b.receptive_field, r) for b, r in zip(self.up_sampling_blocks, red[1:])]
        fields += [scaled(b.receptive_field, r) for arc, r in zip(self.right_arc_blocks, red[:-1]) for b in arc]
        fields.append(self.projection.receptive_field)
        return sum_receptive_fields(tuple(fields))
