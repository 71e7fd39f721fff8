"""Registration of the XLA FFI targets and the trace-time size queries (ctypes into the same shared libraries)."""
from __future__ import annotations

import ctypes as C
import os

import jax
import numpy as np

from .. import _lib as _core  # ctypes prototypes of libpdeqx_b200.so; importing it does not import torch

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_PATH = os.environ.get("PDEQX_XLA_FFI_LIB") or os.path.join(_HERE, "libpdeqx_xla_ffi.so")
_shim = None
_precision = _core.PREC_FP32

ACT = {"identity": _core.ACT_IDENTITY, "relu": _core.ACT_RELU, "gelu": _core.ACT_GELU_TANH}
BOUNDARY = {"zeros": _core.PAD_ZEROS, "circular": _core.PAD_WRAP, "reflect": _core.PAD_REFLECT, "replicate": _core.PAD_EDGE}


def set_precision(mode: str):
    """'fp32' (parity mode, rel-L2 <= 1e-5 against the reference on the JAX CPU backend) or 'tf32' (single-pass TF32 on
    tcgen05, <= 2e-3). Read when an op is TRACED (a static attribute of the custom call), like jax's own
    default_matmul_precision."""
    global _precision
    _precision = {"fp32": _core.PREC_FP32, "tf32": _core.PREC_TF32}[mode]


def get_precision() -> int:
    return _precision


def register():
    """Load libpdeqx_xla_ffi.so and register every handler as an XLA FFI target named `pdeqx_<op>` (platform CUDA).
    The spectral handlers also run in the INITIALIZE stage, where they build their plan (twiddle tables: cudaMalloc +
    cudaMemcpy, not allowed while XLA captures a command buffer)."""
    global _shim
    if _shim is not None:
        return _shim
    if not os.path.exists(SHIM_PATH):
        raise RuntimeError(f"{SHIM_PATH} is missing: build it with `make xla_ffi` (there is no CPU fallback)")
    _core.lib()  # libpdeqx_b200.so first (the shim links against it; rpath $ORIGIN)
    shim = C.CDLL(SHIM_PATH)
    shim.pdeqx_ffi_handler_names.restype = C.POINTER(C.c_char_p)
    n = C.c_int()
    names = shim.pdeqx_ffi_handler_names(C.byref(n))
    for i in range(n.value):
        sym = names[i].decode()
        capsule = jax.ffi.pycapsule(getattr(shim, sym))
        target = "pdeqx_" + sym[len("pdeqx_ffi_"):]
        stages = {"initialize": capsule, "execute": capsule} if "spectral" in sym else capsule
        jax.ffi.register_ffi_target(target, stages, platform="CUDA")
    shim.pdeqx_ffi_spectral_sizes.argtypes = [C.POINTER(_core.SpectralDesc), C.POINTER(C.c_size_t * 3)]
    shim.pdeqx_ffi_spectral_sizes.restype = C.c_int
    _shim = shim
    return shim


def spectral_sizes(ndim, batch, c_in, c_out, spatial, modes, precision):
    """(floats of x_hat, floats of z, bytes of workspace) for one plan geometry -- the result shapes of the custom call."""
    d = _core.SpectralDesc()
    d.ndim, d.batch, d.c_in, d.c_out = ndim, batch, c_in, c_out
    d.spatial, d.modes, d.precision = _core.ivec(spatial, 1), _core.ivec(modes, 1), precision
    out = (C.c_size_t * 3)()
    rc = register().pdeqx_ffi_spectral_sizes(C.byref(d), C.byref(out))
    if rc != 0:
        raise ValueError(_core.lib().pdeqx_last_error().decode())  # e.g. num_modes larger than the grid allows
    return int(out[0]), int(out[1]), int(out[2])


def conv_desc(batch, c_in, c_out, spatial, kernel, cfg):
    d = _core.ConvDesc()
    d.ndim, d.batch, d.c_in, d.c_out, d.groups = len(spatial), batch, c_in, c_out, cfg.groups
    d.spatial, d.kernel = _core.ivec(spatial, 1), _core.ivec(kernel, 1)
    d.stride, d.dilation = _core.ivec(cfg.stride, 1), _core.ivec(cfg.dilation, 1)
    d.pad_lo, d.pad_hi = _core.ivec(cfg.pad_lo, 0), _core.ivec(cfg.pad_hi, 0)
    d.boundary, d.precision = cfg.boundary, cfg.precision
    return d


def conv_out_and_workspace(batch, c_in, c_out, spatial, kernel, cfg):
    """(output spatial shape, workspace floats) of a PhysicsConv call."""
    d = conv_desc(batch, c_in, c_out, spatial, kernel, cfg)
    out = (C.c_int32 * _core.MAX_DIMS)()
    _core.check(_core.lib().pdeqx_conv_out_spatial(d, out))
    nws = _core.lib().pdeqx_conv_workspace_bytes(d)
    return tuple(out[a] for a in range(len(spatial))), nws // 4 + 1


def i32(v):
    return np.int32(v)


def ints(v):
    return np.asarray(v, dtype=np.int64)
