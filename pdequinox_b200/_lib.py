"""ctypes binding of libpdeqx_b200.so (the C ABI declared in include/pdeqx_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
Device memory, streams and process groups come from torch (plumbing); every arithmetic kernel on
the hot path lives in the shared library.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_DIMS = 3
ACT_IDENTITY, ACT_RELU, ACT_GELU_TANH = 0, 1, 2
PAD_ZEROS, PAD_WRAP, PAD_REFLECT, PAD_EDGE = 0, 1, 2, 3
PREC_FP32, PREC_TF32 = 0, 1
K_SPECTRAL_SLAB, K_SPECTRAL_R2C, K_SPECTRAL_MODES, K_POINTWISE_WGRAD, K_CONV_FWD, K_CONV_BWD_DATA, K_CONV_BWD_FILTER, K_DP_EXCHANGE = range(1, 9)

LIB_PATH = os.environ.get("PDEQX_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpdeqx_b200.so")


class PdeqxError(RuntimeError):
    pass


class SpectralDesc(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("batch", C.c_int32), ("c_in", C.c_int32), ("c_out", C.c_int32),
                ("spatial", C.c_int32 * MAX_DIMS), ("modes", C.c_int32 * MAX_DIMS), ("precision", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("batch", C.c_int32), ("c_in", C.c_int32), ("c_out", C.c_int32),
                ("groups", C.c_int32), ("spatial", C.c_int32 * MAX_DIMS), ("kernel", C.c_int32 * MAX_DIMS),
                ("stride", C.c_int32 * MAX_DIMS), ("dilation", C.c_int32 * MAX_DIMS),
                ("pad_lo", C.c_int32 * MAX_DIMS), ("pad_hi", C.c_int32 * MAX_DIMS), ("boundary", C.c_int32),
                ("precision", C.c_int32)]


class ConvTDesc(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("batch", C.c_int32), ("c_in", C.c_int32), ("c_out", C.c_int32),
                ("groups", C.c_int32), ("spatial", C.c_int32 * MAX_DIMS), ("kernel", C.c_int32 * MAX_DIMS),
                ("stride", C.c_int32 * MAX_DIMS), ("dilation", C.c_int32 * MAX_DIMS),
                ("pad_lo", C.c_int32 * MAX_DIMS), ("pad_hi", C.c_int32 * MAX_DIMS),
                ("output_padding", C.c_int32 * MAX_DIMS), ("boundary", C.c_int32)]


_P = C.c_void_p
_I32, _I64, _F = C.c_int32, C.c_int64, C.c_float


class SpectralBwdArgs(C.Structure):
    """pdeqx_spectral_bwd_args (include/pdeqx_b200.h): every argument of the spectral block's VJP, named."""
    _fields_ = [("x", _P), ("gy", _P), ("gy_w", _P), ("lift_u", _P), ("lift_gw", _P), ("lift_gb", _P), ("lift_w", _P),
                ("lift_b", _P), ("proj_gw", _P), ("w_re", _P), ("w_im", _P), ("bypass_w", _P), ("bypass_b", _P),
                ("activation", C.c_int32), ("saved_xhat", _P), ("saved_z", _P), ("gx", _P), ("gw_re", _P), ("gw_im", _P),
                ("g_bypass_w", _P), ("g_bypass_b", _P), ("scratch_full", _P), ("workspace", _P),
                ("workspace_bytes", C.c_size_t), ("stream", _P), ("gy_is_gpre", C.c_int32), ("up_plan", _P), ("up_x", _P),
                ("up_saved_z", _P), ("up_bypass_w", _P), ("up_bypass_b", _P), ("up_activation", C.c_int32),
                ("up_lift_u", _P), ("up_lift_w", _P), ("up_lift_b", _P), ("up_gpre", _P)]

# name -> (restype, argtypes); every symbol include/pdeqx_b200.h declares
PROTOTYPES = {
    "pdeqx_last_error": (C.c_char_p, []),
    "pdeqx_version": (C.c_char_p, []),
    "pdeqx_launch_count": (C.c_uint64, []),
    "pdeqx_has_tcgen05": (C.c_int, []),
    "pdeqx_current_device": (C.c_int, []),
    "pdeqx_set_fast_paths": (None, [C.c_int]),
    "pdeqx_timing_enable": (None, [C.c_int]),
    "pdeqx_timing_read": (C.c_int, [_I32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "pdeqx_timing_read_ex": (C.c_int, [_I32, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "pdeqx_spectral_plan_create": (C.c_int, [C.POINTER(SpectralDesc), C.POINTER(_P)]),
    "pdeqx_spectral_plan_destroy": (None, [_P]),
    "pdeqx_spectral_workspace_bytes": (C.c_size_t, [_P]),
    "pdeqx_spectral_xhat_floats": (C.c_size_t, [_P]),
    "pdeqx_spectral_z_floats": (C.c_size_t, [_P]),
    "pdeqx_spectral_plan_tcgen05_stages": (C.c_int, [_P]),
    "pdeqx_spectral_block_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, C.c_size_t, _P]),
    "pdeqx_spectral_block_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                           C.c_size_t, _P]),
    "pdeqx_spectral_block_rank1_ok": (C.c_int, [_P, _I32, _I32]),
    "pdeqx_spectral_block_bwd_ex": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P,
                                              _P, _P, _P, C.c_size_t, _P]),
    "pdeqx_spectral_block_bwd_args": (C.c_int, [_P, C.POINTER(SpectralBwdArgs)]),
    "pdeqx_spectral_block_chain_ok": (C.c_int, [_P, _P, _I32]),
    "pdeqx_spectral_block_fwd_ex": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, C.c_size_t, _P]),
    "pdeqx_pointwise_fwd": (C.c_int, [_I32, _I32, _I32, _I64, _P, _P, _P, _P, _I32, _P, _P]),
    "pdeqx_pointwise_bwd": (C.c_int, [_I32, _I32, _I32, _I64, _P, _P, _P, _P, _P, _P, _P]),
    "pdeqx_conv_workspace_bytes": (C.c_size_t, [C.POINTER(ConvDesc)]),
    "pdeqx_conv_tcgen05_passes": (C.c_int, [C.POINTER(ConvDesc)]),
    "pdeqx_conv_pair_launches": (C.c_int64, []),
    "pdeqx_conv_out_spatial": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(C.c_int32 * MAX_DIMS)]),
    "pdeqx_conv_fwd": (C.c_int, [C.POINTER(ConvDesc), _P, _P, _P, _P, _I32, _P, _P, C.c_size_t, _P]),
    "pdeqx_conv_bwd_data": (C.c_int, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "pdeqx_conv_bwd_filter": (C.c_int, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "pdeqx_convT_out_spatial": (C.c_int, [C.POINTER(ConvTDesc), C.POINTER(C.c_int32 * MAX_DIMS)]),
    "pdeqx_convT_fwd": (C.c_int, [C.POINTER(ConvTDesc), _P, _P, _P, _P, _P]),
    "pdeqx_convT_bwd_data": (C.c_int, [C.POINTER(ConvTDesc), _P, _P, _P, _P]),
    "pdeqx_convT_bwd_filter": (C.c_int, [C.POINTER(ConvTDesc), _P, _P, _P, _P, _P]),
    "pdeqx_concat_channels": (C.c_int, [_I32, _I32, _I32, _I64, _P, _P, _P, _P]),
    "pdeqx_split_channels": (C.c_int, [_I32, _I32, _I32, _I64, _P, _P, _P, _P]),
    "pdeqx_groupnorm_workspace_bytes": (C.c_size_t, [_I32, _I32]),
    "pdeqx_groupnorm_fwd": (C.c_int, [_I32, _I32, _I32, _I64, _P, _P, _P, _F, _P, _P, _P, _P]),
    "pdeqx_groupnorm_bwd": (C.c_int, [_I32, _I32, _I32, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pdeqx_activation_fwd": (C.c_int, [_I64, _P, _I32, _P, _P]),
    "pdeqx_conv_gn_fusable": (C.c_int, [C.POINTER(ConvDesc)]),
    "pdeqx_conv_gn_slots": (C.c_int, [C.POINTER(ConvDesc)]),
    "pdeqx_conv_gn_fwd": (C.c_int, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, C.c_size_t, _P]),
    "pdeqx_conv_gn_bwd_filter": (C.c_int, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "pdeqx_conv_gn_bwd_data": (C.c_int, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "pdeqx_groupnorm_stats": (C.c_int, [_I32, _I32, _I32, _I64, _P, _F, _P, _P, _P]),
    "pdeqx_groupnorm_stats_from_parts": (C.c_int, [_I32, _I32, _I64, _P, _F, _P, _P]),
    "pdeqx_add_parts": (C.c_int, [_I32, _I64, _I32, _P, _P, _P, _P, _P]),
    "pdeqx_groupnorm_bwd_fused": (C.c_int, [_I32, _I32, _I64, _P, _P, _P, _P, _P, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "pdeqx_add": (C.c_int, [_I64, _P, _P, _P, _P]),
    "pdeqx_activation_bwd": (C.c_int, [_I64, _P, _P, _I32, _P, _P]),
    "pdeqx_mse_loss": (C.c_int, [_I64, _P, _P, _F, _P, _P, _P, _P]),
    "pdeqx_adam_step": (C.c_int, [_I64, _P, _P, _P, _P, _F, _F, _F, _F, _I32, _F, _P]),
    "pdeqx_adam_step_dev": (C.c_int, [_I64, _P, _P, _P, _P, _F, _F, _F, _F, _P, _F, _P]),
}

MAX_RANKS = 8


class DpPeers(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("grads", _P * MAX_RANKS), ("params", _P * MAX_RANKS),
                ("pads", _P * MAX_RANKS), ("mc_grads", _P), ("mc_params", _P)]


PROTOTYPES.update({
    "pdeqx_dp_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "pdeqx_dp_free": (C.c_int, [_P]),
    "pdeqx_dp_export": (C.c_int, [_P, C.c_char * 64]),
    "pdeqx_dp_import": (C.c_int, [C.c_char * 64, C.POINTER(_P)]),
    "pdeqx_dp_unimport": (C.c_int, [_P]),
    "pdeqx_dp_pad_bytes": (C.c_size_t, []),
    "pdeqx_dp_shard_floats": (C.c_int64, [_I64, _I32]),
    "pdeqx_dp_adam_step": (C.c_int, [C.POINTER(DpPeers), _I64, _P, _P, _F, _F, _F, _F, _P, _P, _P]),
    "pdeqx_dp_check": (C.c_int, [C.POINTER(DpPeers), _P]),
})

_lib = None


def lib():
    """The loaded shared library (loaded on first use). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PdeqxError(
                f"{LIB_PATH} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "pdequinox_b200 has no CPU / PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        msg = lib().pdeqx_last_error().decode()
        kind = {-1: "invalid argument", -2: "unsupported", -3: "CUDA error", -4: "workspace"}.get(rc, str(rc))
        if rc == -1:
            raise ValueError(f"pdeqx: {msg}")
        raise PdeqxError(f"pdeqx ({kind}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL). The tensor must be CUDA float32 contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise PdeqxError("pdequinox_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous() or t.dtype.itemsize != 4:
        raise PdeqxError("tensor must be contiguous float32")
    return C.c_void_p(t.data_ptr())


def stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ivec(vals, fill):
    arr = (C.c_int32 * MAX_DIMS)()
    for a in range(MAX_DIMS):
        arr[a] = int(vals[a]) if a < len(vals) else fill
    return arr
