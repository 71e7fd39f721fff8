"""The CUDA path against the golden vectors of the reference's own source (fp32 mode, rel-L2 <= 1e-5).
Weights are loaded by pytree PATH, so this also pins that our modules expose the reference's leaf names."""
import numpy as np
import pytest
import torch

import golden_cases as gc
import pdequinox_b200 as pdeqx
from conftest import rel_l2

pytestmark = pytest.mark.gpu

SUPPORTED = {"SpectralConv", "SpectralConvRaw", "PhysicsConv", "PointwiseLinearConv", "ClassicSpectralBlock",
             "ClassicResBlock", "DilatedResBlock", "ClassicFNO", "ClassicResNet", "DilatedResNet", "PhysicsConvTranspose",
             "ClassicDoubleConvBlock", "LinearConvDownBlock", "LinearConvUpBlock", "ClassicUNet", "ModernResBlock",
             "ModernResNet", "ConvNet", "ModernUNet", "LinearConvBlock"}
CASES = [c for c in gc.CASES if c["kind"] in SUPPORTED]


def build(case, p):
    kind, kw = case["kind"], dict(case["kwargs"])
    if "num_modes" in kw and isinstance(kw["num_modes"], list):
        kw["num_modes"] = tuple(kw["num_modes"])
    if kind == "SpectralConvRaw":
        wr = p["weights_real"]
        # overwritten weights carry the call-time shape (G, o, i, *modes) == an allocation with in=o, out=i
        return pdeqx.SpectralConv(wr.ndim - 3, wr.shape[1], wr.shape[2], kw["num_modes"], key=0)
    cls = {"SpectralConv": pdeqx.SpectralConv, "PhysicsConv": pdeqx.PhysicsConv, "PointwiseLinearConv": pdeqx.PointwiseLinearConv,
           "ClassicSpectralBlock": pdeqx.blocks.ClassicSpectralBlock, "ClassicResBlock": pdeqx.blocks.ClassicResBlock,
           "DilatedResBlock": pdeqx.blocks.DilatedResBlock, "ClassicFNO": pdeqx.arch.ClassicFNO,
           "ClassicResNet": pdeqx.arch.ClassicResNet, "DilatedResNet": pdeqx.arch.DilatedResNet,
           "PhysicsConvTranspose": pdeqx.PhysicsConvTranspose, "ClassicDoubleConvBlock": pdeqx.blocks.ClassicDoubleConvBlock,
           "LinearConvDownBlock": pdeqx.blocks.LinearConvDownBlock, "LinearConvUpBlock": pdeqx.blocks.LinearConvUpBlock,
           "ClassicUNet": pdeqx.arch.ClassicUNet, "ModernResBlock": pdeqx.blocks.ModernResBlock,
           "ModernResNet": pdeqx.arch.ModernResNet, "ConvNet": pdeqx.arch.ConvNet, "ModernUNet": pdeqx.arch.ModernUNet,
           "LinearConvBlock": pdeqx.blocks.LinearConvBlock}[kind]
    if isinstance(kw.get("kernel_size"), list):
        kw["kernel_size"] = tuple(kw["kernel_size"])
    for name in ("stride", "output_padding", "channel_multipliers"):
        if isinstance(kw.get(name), list):
            kw[name] = tuple(kw[name])
    return cls(**kw, key=0)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_path_reproduces_reference_outputs(cuda, case):
    pdeqx.set_precision("fp32")
    x, y, p = gc.arrays(case["name"])
    m = build(case, p)
    leaves = {path: (owner, field) for path, owner, field in m.named_leaves()}
    assert set(leaves) == set(p), (sorted(leaves), sorted(p))  # same pytree paths as the reference
    for path, (owner, field) in leaves.items():
        assert tuple(getattr(owner, field).shape) == p[path].shape, path
        setattr(owner, field, torch.from_numpy(np.ascontiguousarray(p[path], dtype=np.float32)).cuda())
    out = m(torch.from_numpy(x.astype(np.float32)).cuda()).cpu().numpy()
    assert out.shape == y.shape
    assert rel_l2(out, y) <= 1e-5
