"""The XLA FFI shim (pdequinox_b200/csrc/xla_ffi/pdeqx_xla_ffi.cc -> libpdeqx_xla_ffi.so) on a box without a GPU: it loads,
exports one handler per entry point the reference's modules bind, answers XLA's metadata query (API version + traits)
and rejects malformed call frames through XLA_FFI_Error_Create before touching the device. The compute paths of the same
handlers are driven on the GPU by tests/test_gpu_xla_ffi_shim.py."""
import os

import pytest

import xla_ffi_frame as xf

EXPECTED = ["pdeqx_ffi_spectral_block_fwd", "pdeqx_ffi_spectral_block_bwd", "pdeqx_ffi_conv_fwd", "pdeqx_ffi_conv_bwd",
            "pdeqx_ffi_convT_fwd", "pdeqx_ffi_convT_bwd", "pdeqx_ffi_pointwise_fwd", "pdeqx_ffi_pointwise_bwd",
            "pdeqx_ffi_groupnorm_fwd", "pdeqx_ffi_groupnorm_bwd", "pdeqx_ffi_adam", "pdeqx_ffi_norm_conv_fwd",
            "pdeqx_ffi_norm_conv_bwd"]


@pytest.fixture(scope="module")
def shim():
    if not os.path.exists(xf.SHIM):
        pytest.fail(f"{xf.SHIM} is missing: run `make` (or __graft_entry__.build())")
    return xf.load_shim()


def test_shim_exports_every_handler(shim):
    _, handlers = shim
    assert sorted(handlers) == sorted(EXPECTED)


@pytest.mark.parametrize("name", EXPECTED)
def test_metadata_query(shim, name):
    """XLA validates a handler at registration by calling it with the metadata extension only."""
    rc, major, minor, traits = xf.query_metadata(shim[1][name])
    assert rc is None and (major, minor) == (0, 1)
    assert traits & 1  # command-buffer compatible: the handlers only enqueue on the given stream


def test_malformed_frames_are_rejected_without_a_device(shim):
    _, h = shim
    call = xf.Caller()
    with pytest.raises(xf.FfiError) as e:  # wrong operand count
        call(h["pdeqx_ffi_pointwise_fwd"], [(None, (1, 2, 4))], [(None, (1, 2, 4))], {"activation": 0}, stage=xf.STAGE_INITIALIZE)
    assert e.value.code == 3 and "expected 4 operands / 1 results" in e.value.message
    with pytest.raises(xf.FfiError) as e:  # missing static attribute
        call(h["pdeqx_ffi_pointwise_fwd"], [(None, (1, 2, 4)), (None, (3, 2, 1)), (None, (0,)), (None, (0,))], [(None, (1, 3, 4))], {},
             stage=xf.STAGE_INITIALIZE)
    assert e.value.code == 3 and "'activation' is missing" in e.value.message
    with pytest.raises(xf.FfiError) as e:  # channel mismatch between x and w
        call(h["pdeqx_ffi_pointwise_fwd"], [(None, (1, 2, 4)), (None, (3, 5, 1)), (None, (0,)), (None, (0,))], [(None, (1, 3, 4))],
             {"activation": 0}, stage=xf.STAGE_INITIALIZE)
    assert e.value.code == 3
    with pytest.raises(xf.FfiError) as e:  # conv without its static attributes
        call(h["pdeqx_ffi_conv_fwd"], [(None, (1, 2, 8)), (None, (2, 2, 3)), (None, (0,)), (None, (0,))],
             [(None, (1, 2, 8)), (None, (1,))], {"activation": 0}, stage=xf.STAGE_INITIALIZE)
    assert "static attributes" in e.value.message
    # a well-formed frame passes validation in a non-execute stage without needing a stream or a device
    call(h["pdeqx_ffi_conv_fwd"], [(None, (1, 2, 8)), (None, (2, 2, 3)), (None, (0,)), (None, (0,))], [(None, (1, 2, 8)), (None, (1,))],
         {"activation": 0, "groups": 1, "boundary": 1, "precision": 0, "stride": [1], "dilation": [1], "pad_lo": [1], "pad_hi": [1]},
         stage=xf.STAGE_INITIALIZE)
    with pytest.raises(xf.FfiError) as e:  # Adam must run in place
        call(h["pdeqx_ffi_adam"], [(0x1000, (8,)), (0x2000, (8,)), (0x3000, (8,)), (0x4000, (8,)), (0x5000, (4,))],
             [(0x9000, (8,)), (0x3000, (8,)), (0x4000, (8,)), (0x5000, (4,))],
             {"lr": 3e-4, "b1": 0.9, "b2": 0.999, "eps": 1e-8, "grad_scale": 1.0}, stage=xf.STAGE_INITIALIZE)
    assert "must alias" in e.value.message


def test_jax_side_package_is_guarded_and_compiles():
    """pdequinox_b200.jax_ffi (registration, custom_vjp / custom_vmap ops, the reference's modules rebound) needs jax and
    equinox, which this image does not have: importing it must fail with a clear ImportError and must never be triggered
    by the torch-hosted package; its sources must at least compile."""
    import importlib
    import py_compile
    import sys
    pkg = os.path.join(xf.ROOT, "pdequinox_b200", "jax_ffi")
    for f in sorted(os.listdir(pkg)):
        if f.endswith(".py"):
            py_compile.compile(os.path.join(pkg, f), doraise=True)
    import pdequinox_b200  # noqa: F401
    assert "pdequinox_b200.jax_ffi" not in sys.modules
    try:
        import jax  # noqa: F401
        import equinox  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="needs jax"):
            importlib.import_module("pdequinox_b200.jax_ffi")
