"""JAX side of the drop-in: the reference's own modules (`pdequinox.conv.SpectralConv`, `PhysicsConv`,
`PointwiseLinearConv`, `pdequinox.blocks.ClassicSpectralBlock`) with `__call__` rebound to the sm_100a kernels through XLA
FFI custom calls (`jax.ffi`), `jax.custom_vjp` for the reverse pass and `jax.custom_batching.custom_vmap` so that the
single-sample call convention under `jax.vmap` reaches the batch-native kernels.

    import pdequinox as pdeqx
    from pdequinox_b200 import jax_ffi
    model = jax_ffi.convert(pdeqx.arch.ClassicFNO(2, 1, 1, hidden_channels=64, num_modes=16, key=key))   # same pytree
    loss, grads = eqx.filter_value_and_grad(lambda m, x, y: jnp.mean((jax.vmap(m)(x) - y) ** 2))(model, x, y)

Needs jax (>= 0.4.38: `jax.ffi`), equinox and pdequinox itself, plus `pdequinox_b200/libpdeqx_xla_ffi.so` (`make xla_ffi`)
on a machine with an sm_100 GPU. None of jax / equinox is installable in the image this repository is built in, so this
subpackage is NOT exercised by the test suite here (DESIGN.md "Host language"): what IS tested is every handler it calls,
through hand-built XLA call frames (tests/test_gpu_xla_ffi_shim.py), and the C ABI underneath (the whole GPU suite).
The torch-hosted package (`pdequinox_b200.*`) never imports this subpackage.
"""
try:
    import jax  # noqa: F401
    import jax.ffi  # noqa: F401
    import equinox  # noqa: F401
except ImportError as exc:  # pragma: no cover - exercised only where JAX is absent
    raise ImportError(
        "pdequinox_b200.jax_ffi needs jax (with jax.ffi), equinox and pdequinox; the torch-hosted modules in "
        "pdequinox_b200 itself have no such dependency") from exc

from ._ffi import register, set_precision, get_precision  # noqa: E402,F401
from ._ops import norm_conv_act, physics_conv, pointwise_conv, spectral_block  # noqa: E402,F401
from ._modules import ClassicSpectralBlock, PhysicsConv, PointwiseLinearConv, SpectralConv, convert  # noqa: E402,F401
from ._train import data_parallel_step  # noqa: E402,F401
