#!/usr/bin/env python
"""Generates tests/golden/hotpath_golden.npz by executing the REFERENCE'S OWN SOURCE
(/root/reference/pdequinox, unmodified, read-only) on top of the NumPy stand-in for jax/equinox in
tests/golden/_jaxshim (neither package is installable in this image). Run it in the build container only:

    python tests/golden/make_golden.py

The fixtures pin everything the reference's Python code decides -- corner-slice order and weight
interpretation of SpectralConv, SAME-padding arithmetic and padding modes of PhysicsConv, the
transposed-conv index arithmetic, block wiring (conv/norm/activation/bypass order), Sequential /
Hierarchical plumbing -- in float64. Third-party numerics enter through the shim: FFTs via numpy's
pocketfft (what JAX's CPU backend uses), lax.conv_general_dilated via torch's CPU convolution (independent
of the oracle's tap loop), GroupNorm / gelu from their documented definitions.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_jaxshim"))
sys.path.insert(0, "/root/reference")

import jax  # noqa: E402  (the shim)
import pdequinox as ref  # noqa: E402  (the reference, unmodified)

rng = np.random.default_rng(20261019)
arrays, manifest = {}, []


def named_leaves(obj, prefix=""):
    """{path: array} walking Modules / lists / tuples by attribute name (the reference pytree's paths)."""
    out = {}
    if isinstance(obj, np.ndarray):
        out[prefix.rstrip(".")] = obj
    elif isinstance(obj, (list, tuple)):
        for i, o in enumerate(obj):
            out.update(named_leaves(o, f"{prefix}{i}."))
    elif hasattr(obj, "__dict__") and not isinstance(obj, type):
        for k in field_order(obj):
            out.update(named_leaves(vars(obj)[k], f"{prefix}{k}."))
    return out


def field_order(obj):
    """Attribute names in dataclass-FIELD order (class annotations, base classes first): the order in which equinox flattens a
    Module, i.e. the reference's leaf order -- not the order in which __init__ happens to assign them."""
    names = []
    for cls in reversed(type(obj).__mro__):
        for k in getattr(cls, "__annotations__", {}):
            if k not in names:
                names.append(k)
    have = vars(obj)
    return [k for k in names if k in have] + [k for k in have if k not in names]


def randomise(obj):
    """Overwrite every parameter with float64 values of the same shape (gives non-trivial biases / norm affines)."""
    if isinstance(obj, (list, tuple)):
        for o in obj:
            randomise(o)
        return
    if not hasattr(obj, "__dict__") or isinstance(obj, type):
        return
    for k, v in list(vars(obj).items()):
        if isinstance(v, np.ndarray) and v.dtype.kind == "f":
            scale = max(float(np.abs(v).max()), 0.05) if v.size else 1.0
            new = rng.uniform(-scale, scale, v.shape)
            if k == "weight" and v.ndim == 1:  # GroupNorm scale around 1
                new = 1.0 + 0.3 * rng.standard_normal(v.shape)
            object.__setattr__(obj, k, new.astype(np.float64))
        else:
            randomise(v)


def record(name, kind, ctor_kwargs, module, x, extra=None):
    randomise(module)
    y = np.asarray(module(x))
    assert y.dtype == np.float64, (name, y.dtype)
    arrays[f"{name}/x"] = x
    arrays[f"{name}/y"] = y
    for path, a in named_leaves(module).items():
        arrays[f"{name}/p/{path}"] = np.asarray(a)
    manifest.append({"name": name, "kind": kind, "kwargs": ctor_kwargs, **(extra or {})})


key = jax.random.PRNGKey(0)

# ---- SpectralConv: allocation (G, in, out, *modes), call-time (G, o, i, *modes) -> use in == out or overwrite ----
for name, D, ci, co, spatial, modes in [
        ("spectral_1d", 1, 3, 3, (20,), (7,)), ("spectral_1d_nyquist", 1, 2, 2, (8,), (5,)),
        ("spectral_1d_odd", 1, 2, 2, (9,), (5,)), ("spectral_2d", 2, 3, 3, (12, 10), (4, 3)),
        ("spectral_2d_overlap", 2, 2, 2, (6, 16), (4, 3)), ("spectral_3d", 3, 2, 2, (8, 6, 10), (3, 2, 4)),
        ("spectral_2d_tc_shape", 2, 32, 32, (8, 128), (3, 16))]:
    m = ref.conv.SpectralConv(D, ci, co, modes, key=key)
    record(name, "SpectralConv", dict(num_spatial_dims=D, in_channels=ci, out_channels=co, num_modes=list(modes)), m,
           rng.standard_normal((ci,) + spatial))
# in != out through overwritten weights (what tests/test_spectral_conv.py:180-201 of the reference does)
m = ref.conv.SpectralConv(2, 2, 5, (3, 4), key=key)
m.weights_real = rng.standard_normal((2, 5, 2, 3, 4))
m.weights_imag = rng.standard_normal((2, 5, 2, 3, 4))
x = rng.standard_normal((2, 10, 9))
arrays["spectral_2d_in_ne_out/x"], arrays["spectral_2d_in_ne_out/y"] = x, np.asarray(m(x))
arrays["spectral_2d_in_ne_out/p/weights_real"], arrays["spectral_2d_in_ne_out/p/weights_imag"] = m.weights_real, m.weights_imag
manifest.append({"name": "spectral_2d_in_ne_out", "kind": "SpectralConvRaw", "kwargs": dict(num_modes=[3, 4])})

# ---- PhysicsConv ----
i = 0
for D, ci, co, spatial, k, stride, dil, groups in [
        (1, 3, 4, (17,), 3, 1, 1, 1), (1, 4, 6, (16,), 3, 2, 1, 2), (1, 2, 2, (12,), 4, 1, 2, 1), (1, 2, 3, (9,), 5, 1, 3, 1),
        (2, 3, 5, (10, 12), 3, 1, 1, 1), (2, 4, 4, (16, 16), 3, 1, 4, 1), (2, 2, 2, (9, 8), (3, 2), (2, 1), 1, 1),
        (2, 4, 4, (12, 12), 3, 1, 8, 1), (3, 2, 3, (6, 7, 8), 3, 1, 1, 1), (3, 2, 2, (8, 8, 8), 3, 2, 1, 1)]:
    for mode in ("periodic", "dirichlet", "neumann"):
        kw = dict(num_spatial_dims=D, in_channels=ci, out_channels=co, kernel_size=k, stride=stride, dilation=dil,
                  groups=groups, boundary_mode=mode)
        m = ref.conv.PhysicsConv(**kw, key=key)
        record(f"conv_{i}_{mode}", "PhysicsConv", kw, m, rng.standard_normal((ci,) + spatial))
    i += 1

# ---- PhysicsConvTranspose (next tier; pins the oracle's restatement of _conv.py:378-471) ----
i = 0
for D, ci, co, spatial, k, stride, op in [(1, 2, 3, (8,), 3, 2, 1), (2, 2, 2, (6, 5), 3, 2, 1), (1, 3, 2, (7,), 3, 1, 0),
                                          (3, 1, 2, (4, 4, 4), 3, 2, 1)]:
    for mode in ("periodic", "dirichlet", "neumann"):
        kw = dict(num_spatial_dims=D, in_channels=ci, out_channels=co, kernel_size=k, stride=stride, output_padding=op,
                  boundary_mode=mode)
        m = ref.conv.PhysicsConvTranspose(**kw, key=key)
        record(f"convT_{i}_{mode}", "PhysicsConvTranspose", kw, m, rng.standard_normal((ci,) + spatial))
    i += 1

# ---- PointwiseLinearConv ----
for D, ci, co, spatial in [(1, 1, 8, (16,)), (2, 8, 1, (6, 7)), (3, 4, 5, (3, 4, 5))]:
    kw = dict(num_spatial_dims=D, in_channels=ci, out_channels=co)
    record(f"pointwise_{D}d", "PointwiseLinearConv", kw, ref.conv.PointwiseLinearConv(**kw, key=key),
           rng.standard_normal((ci,) + spatial))

# ---- blocks ----
kw = dict(num_spatial_dims=2, in_channels=4, out_channels=4, num_modes=3)
record("block_spectral_gelu", "ClassicSpectralBlock", kw, ref.blocks.ClassicSpectralBlock(**kw, key=key),
       rng.standard_normal((4, 10, 8)), {"activation": "gelu"})
for mode in ("periodic", "dirichlet", "neumann"):
    for use_norm in (False, True):
        kw = dict(num_spatial_dims=2, in_channels=4, out_channels=4, boundary_mode=mode, use_norm=use_norm)
        record(f"block_res_{mode}_{int(use_norm)}", "ClassicResBlock", kw, ref.blocks.ClassicResBlock(**kw, key=key),
               rng.standard_normal((4, 9, 10)), {"activation": "relu"})
    kw = dict(num_spatial_dims=2, in_channels=4, out_channels=4, boundary_mode=mode)
    record(f"block_dilated_{mode}", "DilatedResBlock", kw, ref.blocks.DilatedResBlock(**kw, key=key),
           rng.standard_normal((4, 20, 18)), {"activation": "relu"})
kw = dict(num_spatial_dims=1, in_channels=3, out_channels=5, boundary_mode="periodic", use_norm=True)
record("block_res_in_ne_out", "ClassicResBlock", kw, ref.blocks.ClassicResBlock(**kw, key=key), rng.standard_normal((3, 12)),
       {"activation": "relu"})

# ---- architectures (BASELINE.json configs, scaled down) ----
kw = dict(num_spatial_dims=1, in_channels=1, out_channels=1)
record("arch_fno_1d_default", "ClassicFNO", kw, ref.arch.ClassicFNO(**kw, key=key), rng.standard_normal((1, 64)))
kw = dict(num_spatial_dims=2, in_channels=1, out_channels=1, hidden_channels=8, num_modes=4, num_blocks=3)
record("arch_fno_2d", "ClassicFNO", kw, ref.arch.ClassicFNO(**kw, key=key), rng.standard_normal((1, 16, 12)))
kw = dict(num_spatial_dims=3, in_channels=2, out_channels=1, hidden_channels=4, num_modes=3, num_blocks=2)
record("arch_fno_3d", "ClassicFNO", kw, ref.arch.ClassicFNO(**kw, key=key), rng.standard_normal((2, 8, 8, 8)))
for mode in ("periodic", "dirichlet", "neumann"):
    kw = dict(num_spatial_dims=2, in_channels=1, out_channels=1, hidden_channels=6, num_blocks=2, boundary_mode=mode)
    record(f"arch_resnet_{mode}", "ClassicResNet", kw, ref.arch.ClassicResNet(**kw, key=key), rng.standard_normal((1, 12, 10)))
    kw = dict(num_spatial_dims=2, in_channels=1, out_channels=1, hidden_channels=4, num_blocks=1, boundary_mode=mode)
    record(f"arch_dilated_{mode}", "DilatedResNet", kw, ref.arch.DilatedResNet(**kw, key=key), rng.standard_normal((1, 24, 20)))
kw = dict(num_spatial_dims=1, in_channels=1, out_channels=1, hidden_channels=4, num_levels=2, boundary_mode="dirichlet")
record("arch_unet_1d_dirichlet", "ClassicUNet", kw, ref.arch.ClassicUNet(**kw, key=key), rng.standard_normal((1, 32)))

# ---- round-1 widening: UNet path (appended after the original cases so that their random streams are unchanged) ----
i = 4
for D, ci, co, spatial, k, stride, op, dil in [(1, 2, 2, (8,), 4, 2, 1, 1), (1, 2, 3, (10,), 5, 2, 1, 1), (2, 2, 3, (6, 6), 3, 2, 0, 1),
                                               (1, 2, 2, (9,), 3, 2, 1, 2), (1, 3, 2, (12,), 2, 3, 2, 1), (2, 2, 2, (5, 6), (3, 4), (2, 1), (1, 0), 1)]:
    for mode in ("periodic", "dirichlet", "neumann"):
        kw = dict(num_spatial_dims=D, in_channels=ci, out_channels=co, kernel_size=k, stride=stride, output_padding=op,
                  dilation=dil, boundary_mode=mode)
        m = ref.conv.PhysicsConvTranspose(**kw, key=key)
        record(f"convT_{i}_{mode}", "PhysicsConvTranspose", kw, m, rng.standard_normal((ci,) + spatial))
    i += 1
for mode in ("periodic", "dirichlet", "neumann"):
    for use_norm in (False, True):
        kw = dict(num_spatial_dims=2, in_channels=3, out_channels=4, boundary_mode=mode, use_norm=use_norm)
        record(f"block_double_{mode}_{int(use_norm)}", "ClassicDoubleConvBlock", kw,
               ref.blocks.ClassicDoubleConvBlock(**kw, key=key), rng.standard_normal((3, 9, 10)), {"activation": "relu"})
    kw = dict(num_spatial_dims=1, in_channels=3, out_channels=3, boundary_mode=mode)
    record(f"block_down_{mode}", "LinearConvDownBlock", kw, ref.blocks.LinearConvDownBlock(**kw, key=key), rng.standard_normal((3, 16)))
    kw = dict(num_spatial_dims=1, in_channels=4, out_channels=2, boundary_mode=mode)
    record(f"block_up_{mode}", "LinearConvUpBlock", kw, ref.blocks.LinearConvUpBlock(**kw, key=key), rng.standard_normal((4, 8)))
for mode in ("periodic", "neumann"):
    kw = dict(num_spatial_dims=1, in_channels=1, out_channels=1, hidden_channels=4, num_levels=2, boundary_mode=mode)
    record(f"arch_unet_1d_{mode}", "ClassicUNet", kw, ref.arch.ClassicUNet(**kw, key=key), rng.standard_normal((1, 32)))
for mode in ("periodic", "dirichlet", "neumann"):
    kw = dict(num_spatial_dims=2, in_channels=2, out_channels=1, hidden_channels=4, num_levels=2, boundary_mode=mode)
    record(f"arch_unet_2d_{mode}", "ClassicUNet", kw, ref.arch.ClassicUNet(**kw, key=key), rng.standard_normal((2, 16, 12)))
kw = dict(num_spatial_dims=1, in_channels=1, out_channels=1, boundary_mode="dirichlet")  # BASELINE.json configs[1]: README defaults
record("arch_unet_1d_readme", "ClassicUNet", kw, ref.arch.ClassicUNet(**kw, key=key), rng.standard_normal((1, 64)))
kw = dict(num_spatial_dims=1, in_channels=1, out_channels=1, hidden_channels=4, num_levels=1, use_norm=False, boundary_mode="dirichlet")
record("arch_unet_1d_nonorm", "ClassicUNet", kw, ref.arch.ClassicUNet(**kw, key=key), rng.standard_normal((1, 16)))

# ---- the pre-activation ResNet and the plain ConvNet (appended: earlier random streams unchanged) ----
for mode in ("periodic", "dirichlet", "neumann"):
    for use_norm in (False, True):
        kw = dict(num_spatial_dims=2, in_channels=4, out_channels=4, boundary_mode=mode, use_norm=use_norm)
        record(f"block_modern_{mode}_{int(use_norm)}", "ModernResBlock", kw, ref.blocks.ModernResBlock(**kw, key=key),
               rng.standard_normal((4, 9, 10)), {"activation": "relu"})
    kw = dict(num_spatial_dims=2, in_channels=1, out_channels=1, hidden_channels=6, num_blocks=2, boundary_mode=mode)
    record(f"arch_modern_{mode}", "ModernResNet", kw, ref.arch.ModernResNet(**kw, key=key), rng.standard_normal((1, 12, 10)))
    kw = dict(num_spatial_dims=2, in_channels=2, out_channels=1, hidden_channels=5, depth=3, boundary_mode=mode)
    record(f"arch_convnet_{mode}", "ConvNet", kw, ref.arch.ConvNet(**kw, key=key), rng.standard_normal((2, 11, 9)))
for use_norm in (False, True):
    kw = dict(num_spatial_dims=1, in_channels=3, out_channels=5, boundary_mode="periodic", use_norm=use_norm)
    record(f"block_modern_in_ne_out_{int(use_norm)}", "ModernResBlock", kw, ref.blocks.ModernResBlock(**kw, key=key),
           rng.standard_normal((3, 12)), {"activation": "relu"})
kw = dict(num_spatial_dims=1, in_channels=1, out_channels=2, depth=0, boundary_mode="dirichlet")
record("arch_convnet_depth0", "ConvNet", kw, ref.arch.ConvNet(**kw, key=key), rng.standard_normal((1, 16)))
kw = dict(num_spatial_dims=3, in_channels=1, out_channels=1, hidden_channels=3, depth=2, boundary_mode="periodic")
record("arch_convnet_3d", "ConvNet", kw, ref.arch.ConvNet(**kw, key=key), rng.standard_normal((1, 6, 5, 7)))

for mode in ("periodic", "dirichlet", "neumann"):
    kw = dict(num_spatial_dims=1, in_channels=1, out_channels=1, hidden_channels=4, num_levels=2, boundary_mode=mode)
    record(f"arch_modern_unet_1d_{mode}", "ModernUNet", kw, ref.arch.ModernUNet(**kw, key=key), rng.standard_normal((1, 32)))
    kw = dict(num_spatial_dims=2, in_channels=3, out_channels=4, boundary_mode=mode)
    record(f"block_linear_{mode}", "LinearConvBlock", kw, ref.blocks.LinearConvBlock(**kw, key=key), rng.standard_normal((3, 7, 9)))
kw = dict(num_spatial_dims=2, in_channels=2, out_channels=1, hidden_channels=4, num_levels=2, num_blocks=1, use_norm=False,
          channel_multipliers=[2, 3], boundary_mode="periodic")
record("arch_modern_unet_2d_multipliers", "ModernUNet", kw, ref.arch.ModernUNet(**kw, key=key), rng.standard_normal((2, 16, 12)))

# ---- parameter counts computed by the reference's own count_parameters ----
counts = {}
for nm, ctor in [("ClassicFNO", ref.arch.ClassicFNO), ("ClassicResNet", ref.arch.ClassicResNet),
                 ("DilatedResNet", ref.arch.DilatedResNet), ("ClassicUNet", ref.arch.ClassicUNet),
                 ("ModernResNet", ref.arch.ModernResNet), ("ConvNet", ref.arch.ConvNet),
                 ("ModernUNet", ref.arch.ModernUNet)]:
    for D in (1, 2, 3):
        counts[f"{nm}_{D}d"] = int(ref.count_parameters(ctor(D, 1, 1, key=key)))
counts["ClassicFNO_2d_h64_m16"] = int(ref.count_parameters(ref.arch.ClassicFNO(2, 1, 1, hidden_channels=64, num_modes=16, key=key)))

np.savez_compressed(os.path.join(HERE, "hotpath_golden.npz"), **arrays)
json.dump({"cases": manifest, "parameter_counts": counts,
           "generated_by": "tests/golden/make_golden.py (reference source on the NumPy jax/equinox stand-in)"},
          open(os.path.join(HERE, "hotpath_golden.json"), "w"), indent=1)
print(f"{len(manifest)} cases, {len(arrays)} arrays, {sum(a.nbytes for a in arrays.values()) / 1e6:.2f} MB raw")
print(counts)
