"""Sequential: lifting -> N blocks -> projection (reference: pdequinox/_sequential.py:16-128)."""
from . import random as jr
from ._module import Module
from ._utils import sum_receptive_fields
from .blocks import ClassicResBlockFactory, LinearChannelAdjustBlockFactory
from .conv import PointwiseLinearConv


class Sequential(Module):
    _fields = ("lifting", "blocks", "projection")
    # reverse pass: let consecutive spectral blocks share the kernel between them (ClassicSpectralBlock.chain_with);
    # False = every block runs its own cotangent and data-gradient passes (tests compare the two)
    chain_reverse = True

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
        first, last = self.blocks[0], self.blocks[-1]
        # Inside a training step's tape (no gradient w.r.t. the network input wanted) both pointwise ends of the architecture
        # are folded into the spectral blocks next to them, so that neither full-size tensor is ever written:
        #  * a 1 -> C lifting in front of the first block (ClassicSpectralBlock._fwd_ex(lifting=...));
        #  * a C -> 1 projection behind the last block (…(proj=...)): reduced in that block's last kernel.
        # The matching reverse pass accumulates the lifting's / projection's gradients inside the blocks' kernels.
        folding = tape is not None and getattr(tape, "fold_lifting", False)
        fold = (folding and isinstance(self.lifting, PointwiseLinearConv)
                and self.lifting.in_channels == 1 and hasattr(first, "lifted_ok") and first.lifted_ok(x))
        fold_proj = (folding and isinstance(self.projection, PointwiseLinearConv) and self.projection.out_channels == 1
                     and self.projection.in_channels == getattr(last, "out_channels", self.projection.in_channels)
                     and hasattr(last, "projected_ok") and last.projected_ok(x))
        nb = len(self.blocks)
        if fold:
            tape.push(x)  # stands for the lifting's record (its input)
        else:
            x = self.lifting._fwd(x, tape)
        for i, block in enumerate(self.blocks):
            kw = {}
            if fold and i == 0:
                kw["lifting"] = self.lifting
            if fold_proj and i == nb - 1:
                kw["proj"] = self.projection
            x = block._fwd_ex(x, tape, **kw) if kw else block._fwd(x, tape)
        if fold_proj:
            tape.push(None)  # the projection's record: its input was never written
            return x
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
        x_proj = tape.pop()  # what the projection saved (pushed back below); None: the forward pass folded it away
        projected_fwd = x_proj is None
        if hasattr(last, "rank1_ok") and last.rank1_ok(tape):
            fold_proj = isinstance(self.projection, PointwiseLinearConv) and self.projection.out_channels == 1
            # the first block's forward record sits below the later blocks' records: same plan family, ask the block itself
            fold_lift = (isinstance(self.lifting, PointwiseLinearConv) and self.lifting.in_channels == 1 and not need_gx
                         and hasattr(first, "rank1_ok"))
        if projected_fwd:
            if not fold_proj:
                raise RuntimeError("the projected forward pass needs the rank-one reverse pass of the last block")
            # bias gradient = sum of the cotangent; the weight gradient comes out of the last block's cotangent kernel
            self.projection.bias_grad_only(gy, tape)
        else:
            tape.push(x_proj)
            if fold_proj:
                self.projection._bwd(gy, tape, False)  # weight / bias gradients only
            else:
                g = self.projection._bwd(gy, tape, True)
            hook("projection.")
        gpre_ready = False  # True: `g` is the cotangent of the next block's pre-activation (chained reverse pass)
        premasked = False   # True: `g` already carries the relu' of the next ResNet block's output
        for i in reversed(range(nb)):
            kw = {}
            if fold_proj and i == nb - 1:
                kw["gy_w"] = self.projection.weight
                if projected_fwd:
                    kw["proj_gw"] = self.projection.grad("weight")
                g = gy
            lifted_fwd = i == 0 and tape.stack[-1][0] is None  # the forward pass folded the lifting in: it must be unfolded here
            if i == 0 and (lifted_fwd or (fold_lift and self.blocks[0].rank1_ok(tape))):
                u = tape.stack[-2]  # the lifting's saved input, right below the first block's record
                lw = self.lifting.grad("weight")
                lb = self.lifting.grad("bias") if self.lifting.bias is not None else None
                kw["lift"] = (u, lw, lb)
                if lifted_fwd:
                    kw["lifted"] = (self.lifting.weight, self.lifting.bias)
            # two consecutive spectral blocks share the pass between them: block i's last kernel hands block i-1 the
            # cotangent of its PRE-activation, so block i's input gradient is never written (ClassicSpectralBlock.chain_with)
            if gpre_ready:
                kw["gy_is_gpre"] = True
            gpre_ready = False
            if i > 0 and self.chain_reverse and hasattr(self.blocks[i], "chain_with"):
                up_rec = tape.stack[-2]
                up_lifted = isinstance(up_rec, tuple) and len(up_rec) == 2 and up_rec[0] is None
                chain = self.blocks[i].chain_with(self.blocks[i - 1], tape.stack[-1], up_rec,
                                                  lifting=self.lifting if up_lifted else None,
                                                  u=tape.stack[-3] if up_lifted else None) \
                    if isinstance(up_rec, tuple) and len(up_rec) == 2 and isinstance(up_rec[1], tuple) else None
                if chain is not None:
                    kw["chain"] = chain
                    gpre_ready = True
            # two consecutive relu ResNet blocks: block i forms its input gradient already multiplied by relu' of block i-1's
            # output (its own input), so block i-1 skips that elementwise pass (ClassicResBlock._bwd)
            res_kw = {}
            if hasattr(self.blocks[i], "relu_output_saved"):
                if premasked:
                    res_kw["gy_premasked"] = True
                premasked = False
                if (i > 0 and self.chain_reverse and hasattr(self.blocks[i - 1], "relu_output_saved")
                        and self.blocks[i].relu_output_saved(tape.stack[-1])
                        and self.blocks[i - 1].relu_output_saved(tape.stack[-2])):
                    res_kw["mask_gx"] = True
                    premasked = True
            if kw:
                g = self.blocks[i]._bwd_ex(g, tape, True, **kw)
            elif res_kw:
                g = self.blocks[i]._bwd(g, tape, True, **res_kw)
            else:
                g = self.blocks[i]._bwd(g, tape, True)
            if projected_fwd and i == nb - 1:
                hook("projection.")
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
