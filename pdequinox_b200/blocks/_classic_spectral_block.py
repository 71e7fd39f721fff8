"""ClassicSpectralBlock: activation(spectral_conv(x) + by_pass_conv(x)) as ONE fused call
(reference: pdequinox/blocks/_classic_spectral_block.py:10-131)."""
from .. import nn, random as jr
from ..conv import PointwiseLinearConv, SpectralConv
from ._base_block import Block, BlockFactory


class ClassicSpectralBlock(Block):
    _fields = ("spectral_conv", "by_pass_conv", "activation")

    def __init__(self, num_spatial_dims, in_channels, out_channels, *, activation=nn.gelu, num_modes=8,
                 use_bias=True, zero_bias_init=False, key):
        self.num_spatial_dims = num_spatial_dims
        k_1, k_2 = jr.split(key)
        self.spectral_conv = SpectralConv(num_spatial_dims=num_spatial_dims, in_channels=in_channels,
                                          out_channels=out_channels, num_modes=num_modes, key=k_1)
        self.by_pass_conv = PointwiseLinearConv(num_spatial_dims=num_spatial_dims, in_channels=in_channels,
                                                out_channels=out_channels, use_bias=use_bias,
                                                zero_bias_init=zero_bias_init, key=k_2)
        self.activation = activation

    @property
    def receptive_field(self):
        return self.spectral_conv.receptive_field

    def _fwd(self, x, tape):
        act = nn.activation_code(self.activation)
        y, saved = self.spectral_conv.apply(x, tape, bypass=self.by_pass_conv, act=act, save=tape is not None)
        if tape is not None:
            tape.push((x, saved))
        return y

    def _bwd(self, gy, tape, need_gx=True):
        x, saved = tape.pop()
        return self.spectral_conv.apply_bwd(x, gy, saved, tape, bypass=self.by_pass_conv,
                                            act=nn.activation_code(self.activation), need_gx=need_gx)

    # ---- rank-one cotangent (the block sits right before a C -> 1 pointwise projection) ----
    def rank1_ok(self, tape):
        """True if the reverse pass of the forward call on top of `tape` can take its cotangent as gy_w[o] * gy1[b, pixel]."""
        from .. import _lib
        if not tape.stack:
            return False
        _, saved = tape.stack[-1]
        return bool(_lib.lib().pdeqx_spectral_block_rank1_ok(saved[0].handle, 1, nn.activation_code(self.activation)))

    def lifted_ok(self, u):
        return self.spectral_conv.lifted_ok(u, nn.activation_code(self.activation))

    def projected_ok(self, x):
        return self.spectral_conv.projected_ok(x, nn.activation_code(self.activation))

    def _fwd_ex(self, x, tape, lifting=None, proj=None):
        """Forward pass with the architecture's pointwise ends folded in: `lifting` (a 1 -> C PointwiseLinearConv; `x` is then
        its single-channel input and the lifted tensor is never materialised) and / or `proj` (a C -> 1 PointwiseLinearConv
        behind this block: the projected field is returned and this block's output is never written)."""
        act = nn.activation_code(self.activation)
        y, saved = self.spectral_conv.apply_ex(x, tape, bypass=self.by_pass_conv, act=act, save=tape is not None,
                                               lifting=(lifting.weight, lifting.bias) if lifting is not None else None,
                                               proj=(proj.weight, proj.bias) if proj is not None else None)
        if tape is not None:
            tape.push((None if lifting is not None else x, saved))  # x = None marks the lifted forward pass
        return y

    def _fwd_lifted(self, u, lifting, tape):
        return self._fwd_ex(u, tape, lifting=lifting)

    def _bwd_ex(self, gy, tape, need_gx=True, gy_w=None, lift=None, lifted=None, proj_gw=None, gy_is_gpre=False, chain=None):
        """gy_w: `gy` is the single-channel factor of a rank-one cotangent; lift = (u, gw, gb): accumulate the gradients of
        the 1 -> C lifting that produced this block's input from `u` instead of returning the input gradient; proj_gw: the
        forward pass was projected (`_fwd_ex(proj=...)`): also accumulate the projection's weight gradient; gy_is_gpre / chain:
        see `chain_with` (two consecutive blocks share the pass between them)."""
        x, saved = tape.pop()
        return self.spectral_conv.apply_bwd(x, gy, saved, tape, bypass=self.by_pass_conv,
                                            act=nn.activation_code(self.activation), need_gx=need_gx, gy_w=gy_w, lift=lift,
                                            lifted=lifted if x is None else None, proj_gw=proj_gw, gy_is_gpre=gy_is_gpre,
                                            chain=chain)

    def chain_with(self, up, my_record, up_record, lifting=None, u=None):
        """Description of the block `up` IN FRONT of this one (pdequinox/_sequential.py:111-116) for a chained reverse pass,
        or None if the two do not qualify: this block's data-gradient kernel then also applies act'(pre) of `up`, so the
        tensor between the two blocks' reverse passes is never written or read. `*_record` = what the blocks' forward passes
        pushed on the tape; `lifting` / `u`: `up` ran with a folded lifting of the field `u` (its record holds x = None)."""
        from .. import _lib
        if not isinstance(up, ClassicSpectralBlock) or up.by_pass_conv is None or self.by_pass_conv is None:
            return None
        (_, (plan, _, _)), (up_x, up_saved) = my_record, up_record
        act_up = nn.activation_code(up.activation)
        if up_x is None and (lifting is None or u is None):
            return None
        if not _lib.lib().pdeqx_spectral_block_chain_ok(plan.handle, up_saved[0].handle, act_up):
            return None
        return dict(conv=up.spectral_conv, x=up_x, saved=up_saved, bypass=up.by_pass_conv, act=act_up, u=u,
                    lifted=(lifting.weight, lifting.bias) if up_x is None else None)


class ClassicSpectralBlockFactory(BlockFactory):
    def __init__(self, *, num_modes=8, use_bias: bool = True, zero_bias_init: bool = False):
        self.num_modes = num_modes
        self.use_bias = use_bias
        self.zero_bias_init = zero_bias_init

    def __call__(self, num_spatial_dims, in_channels, out_channels, *, activation, boundary_mode, key):
        return ClassicSpectralBlock(num_spatial_dims=num_spatial_dims, in_channels=in_channels,
                                    out_channels=out_channels, num_modes=self.num_modes, activation=activation,
                                    key=key, use_bias=self.use_bias, zero_bias_init=self.zero_bias_init)
