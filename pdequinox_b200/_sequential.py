"""Sequential: lifting -> N blocks -> projection (reference: pdequinox/_sequential.py:16-128)."""
from . import random as jr
from ._module import Module
from ._utils import sum_receptive_fields
from .blocks import ClassicResBlockFactory, LinearChannelAdjustBlockFactory
from .conv import PointwiseLinearConv


class Sequential(Module):
    _fields = ("lifting", "blocks", "projection")

    def __init__(self, num_spatial_dims, in_channels, out_channels, *, hidden_channels, num_blocks, activation, key,
                 boundary_mode, lifting_factory=None, block_factory=None, projection_factory=None):
        lifting_factory = lifting_factory or LinearChannelAdjustBlockFactory()
        block_factory = block_factory or ClassicResBlockFactory()
        projection_factory = projection_factory or LinearChannelAdjustBlockFactory()
        self.num_spatial_dims = num_spatial_dims
        subkey, key = jr.split(key)  # _sequential.py:66
        if num_blocks < 1:
            raise ValueError("num_blocks must be at least 1")
        if isinstance(hidden_channels, int):
            hidden_channels = (hidden_channels,) * (num_blocks + 1)
        elif len(hidden_channels) != (num_blocks + 1):
            raise ValueError("The list of hidden channels must be one longer than the number of blocks")
        self.lifting = lifting_factory(num_spatial_dims=num_spatial_dims, in_channels=in_channels,
                                       out_channels=hidden_channels[0], activation=activation,
                                       boundary_mode=boundary_mode, key=subkey)
        self.blocks = []
        for fan_in, fan_out in zip(hidden_channels[:-1], hidden_channels[1:]):
            subkey, key = jr.split(key)  # _sequential.py:91
            self.blocks.append(block_factory(num_spatial_dims=num_spatial_dims, in_channels=fan_in,
                                             out_channels=fan_out, activation=activation,
                                             boundary_mode=boundary_mode, key=subkey))
        self.projection = projection_factory(num_spatial_dims=num_spatial_dims, in_channels=hidden_channels[-1],
                                             out_channels=out_channels, activation=activation,
                                             boundary_mode=boundary_mode, key=key)

    def _fwd(self, x, tape):
        first = self.blocks[0]
        # a 1 -> C pointwise lifting in front of a spectral block is folded into that block (the lifted tensor is never
        # written; see ClassicSpectralBlock._fwd_lifted). Only inside a training step's tape with no input gradient wanted
        # downstream of it: the matching reverse pass accumulates the lifting's gradients instead of an input gradient.
        fold = (tape is not None and getattr(tape, "fold_lifting", False) and isinstance(self.lifting, PointwiseLinearConv)
                and self.lifting.in_channels == 1 and hasattr(first, "lifted_ok") and first.lifted_ok(x))
        if fold:
            tape.push(x)  # stands for the lifting's record (its input)
            x = first._fwd_lifted(x, self.lifting, tape)
            rest = self.blocks[1:]
        else:
            x = self.lifting._fwd(x, tape)
            rest = self.blocks
        for block in rest:
            x = block._fwd(x, tape)
        return self.projection._fwd(x, tape)

    def _bwd(self, gy, tape, need_gx=False):
        # `_grad_ready_hook(prefix)` (set by the data-parallel train step) is told as soon as every gradient under a
        # pytree prefix is final, so that its all-reduce can start while the rest of the reverse pass still runs
        hook = getattr(self, "_grad_ready_hook", None) or (lambda prefix: None)
        nb = len(self.blocks)
        first, last = self.blocks[0], self.blocks[-1]
        # Both ends of the architecture are pointwise and single-channel on the outside (1 -> C lifting, C -> 1 projection):
        #  * the projection's data gradient is the outer product weight[o] * gy[b, pixel]; the last spectral block takes the
        #    two factors and forms the cotangent inside its kernel instead of reading a materialised full-size tensor;
        #  * the lifting's parameter gradients are reductions of the first block's input gradient against the input field;
        #    the first block's last kernel accumulates them instead of writing that gradient out.
        # Both need the block's tcgen05 path (`rank1_ok`); otherwise the plain chain below runs.
        fold_proj = fold_lift = False
        x_proj = tape.pop()  # what the projection saved (pushed back below)
        if hasattr(last, "rank1_ok") and last.rank1_ok(tape):
            fold_proj = isinstance(self.projection, PointwiseLinearConv) and self.projection.out_channels == 1
            # the first block's forward record sits below the later blocks' records: same plan family, ask the block itself
            fold_lift = (isinstance(self.lifting, PointwiseLinearConv) and self.lifting.in_channels == 1 and not need_gx
                         and hasattr(first, "rank1_ok"))
        tape.push(x_proj)
        if fold_proj:
            self.projection._bwd(gy, tape, False)  # weight / bias gradients only
            hook("projection.")
        else:
            g = self.projection._bwd(gy, tape, True)
            hook("projection.")
        for i in reversed(range(nb)):
            kw = {}
            if fold_proj and i == nb - 1:
                kw["gy_w"] = self.projection.weight
                g = gy
            lifted_fwd = i == 0 and tape.stack[-1][0] is None  # the forward pass folded the lifting in: it must be unfolded here
            if i == 0 and (lifted_fwd or (fold_lift and self.blocks[0].rank1_ok(tape))):
                u = tape.stack[-2]  # the lifting's saved input, right below the first block's record
                lw = self.lifting.grad("weight")
                lb = self.lifting.grad("bias") if self.lifting.bias is not None else None
                kw["lift"] = (u, lw, lb)
                if lifted_fwd:
                    kw["lifted"] = (self.lifting.weight, self.lifting.bias)
            if kw:
                g = self.blocks[i]._bwd_ex(g, tape, True, **kw)
            else:
                g = self.blocks[i]._bwd(g, tape, True)
            hook(f"blocks.{i}.")
            if "lift" in kw:
                tape.pop()  # the lifting's record: its gradients were just written by the block's kernel
                hook("lifting.")
                return None
        gx = self.lifting._bwd(g, tape, need_gx)
        hook("lifting.")
        return gx

    @property
    def receptive_field(self):
        fields = (self.lifting.receptive_field,) + tuple(b.receptive_field for b in self.blocks) + \
            (self.projection.receptive_field,)
        return sum_receptive_fields(fields)
