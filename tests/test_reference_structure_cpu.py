"""The reference's structural tests for the widened components, restated against this package (CPU only, construction does not
need a GPU): tests/test_instantiate.py:64-91 (every architecture builds with its defaults for every boundary mode),
tests/test_count_parameters.py:845-946 (ModernResBlock parameter formula, block == factory) and :620-700 (LinearConvBlock)."""
import itertools

import pytest

import pdequinox_b200 as pdeqx

ARCHS = [pdeqx.arch.ClassicFNO, pdeqx.arch.ClassicUNet, pdeqx.arch.ClassicResNet, pdeqx.arch.ConvNet,
         pdeqx.arch.DilatedResNet, pdeqx.arch.ModernResNet, pdeqx.arch.ModernUNet]


@pytest.mark.parametrize("D,arch,mode", [(D, a, m) for D in (1, 2) for a in ARCHS for m in ("periodic", "dirichlet", "neumann")])
def test_default_config_instantiates(D, arch, mode):
    kw = {} if arch is pdeqx.arch.ClassicFNO else {"boundary_mode": mode}
    model = arch(D, 2, 5, key=pdeqx.random.PRNGKey(0), **kw)
    assert pdeqx.count_parameters(model) > 0
    assert len(model.receptive_field) == D


GRID = list(itertools.product((1, 2, 3), (1, 2, 5), (1, 2, 5), (2, 3, 4), (True, False), (True, False), ("periodic", "neumann")))


@pytest.mark.parametrize("D,ci,co,k,use_norm,use_bias,mode", GRID[::7])  # every 7th point of the reference's grid
def test_modern_res_block_parameter_formula(D, ci, co, k, use_norm, use_bias, mode):
    block = pdeqx.blocks.ModernResBlock(D, ci, co, kernel_size=k, use_norm=use_norm, use_bias=use_bias, boundary_mode=mode,
                                        key=pdeqx.random.PRNGKey(0))
    expect = (2 * ci if use_norm else 0) + ci * co * k ** D + (co if use_bias else 0)       # norm_1, conv_1
    expect += (2 * co if use_norm else 0) + co * co * k ** D + (co if use_bias else 0)      # norm_2, conv_2
    if ci != co:
        expect += (2 * ci if use_norm else 0) + ci * co  # bypass norm (on the INPUT channels) + bias-free 1x1 conv
    assert pdeqx.count_parameters(block) == expect
    factory = pdeqx.blocks.ModernResBlockFactory(kernel_size=k, use_norm=use_norm, use_bias=use_bias)
    twin = factory(num_spatial_dims=D, in_channels=ci, out_channels=co, activation=pdeqx.nn.relu, boundary_mode=mode,
                   key=pdeqx.random.PRNGKey(0))
    assert pdeqx.count_parameters(twin) == expect


@pytest.mark.parametrize("D,ci,co,k,use_bias", [(1, 1, 5, 3, True), (2, 2, 2, 4, False), (3, 5, 1, 2, True)])
def test_linear_conv_block_parameter_formula(D, ci, co, k, use_bias):
    block = pdeqx.blocks.LinearConvBlock(D, ci, co, kernel_size=k, use_bias=use_bias, boundary_mode="dirichlet", key=0)
    assert pdeqx.count_parameters(block) == ci * co * k ** D + (co if use_bias else 0)
    twin = pdeqx.blocks.LinearConvBlockFactory(kernel_size=k, use_bias=use_bias)(
        num_spatial_dims=D, in_channels=ci, out_channels=co, activation=None, boundary_mode="dirichlet", key=0)
    assert pdeqx.count_parameters(twin) == pdeqx.count_parameters(block)


def test_conv_net_layer_count_and_depth_zero():
    net = pdeqx.ConvNet(2, 3, 4, hidden_channels=7, depth=4, boundary_mode="periodic", key=0)
    assert [(c.in_channels, c.out_channels) for c in net.layers] == [(3, 7), (7, 7), (7, 7), (7, 7), (7, 4)]
    assert net.receptive_field == ((5.0, 5.0), (5.0, 5.0))
    lin = pdeqx.ConvNet(1, 3, 4, depth=0, use_final_bias=False, boundary_mode="periodic", key=0)
    assert len(lin.layers) == 1 and lin.layers[0].bias is None
