"""Batch-sharded train step under JAX: the counterpart of pdequinox_b200/train.py (which issues the same collective through
torch.distributed): `shard_map` over a 1-D device mesh, every device runs the FFI kernels on its shard of the batch, one
`lax.pmean` (XLA's NCCL all-reduce over NVLink) on the gradient pytree. The reference has no multi-device code
(SURVEY.md 2.2); the loss / optimiser lines are README.md:74-86."""
from __future__ import annotations

import equinox as eqx
import jax
import jax.numpy as jnp
from jax.experimental.shard_map import shard_map
from jax.sharding import Mesh, PartitionSpec as P


def data_parallel_step(optimizer, devices=None):
    """Returns `step(model, opt_state, x, y) -> (model, opt_state, loss)`; x, y: (B, C, *S) with B divisible by the number
    of devices. `optimizer`: an optax GradientTransformation (the README uses optax.adam(3e-4))."""
    devices = devices if devices is not None else jax.devices()
    mesh = Mesh(devices, ("data",))

    def local(model, x, y):
        def loss_fn(m):
            return jnp.mean((jax.vmap(m)(x) - y) ** 2)

        loss, grads = eqx.filter_value_and_grad(loss_fn)(model)
        return jax.lax.pmean(loss, "data"), jax.lax.pmean(grads, "data")

    sharded = shard_map(local, mesh=mesh, in_specs=(P(), P("data"), P("data")), out_specs=(P(), P()), check_rep=False)

    @eqx.filter_jit
    def step(model, opt_state, x, y):
        loss, grads = sharded(model, x, y)
        updates, opt_state = optimizer.update(grads, opt_state, eqx.filter(model, eqx.is_array))
        return eqx.apply_updates(model, updates), opt_state, loss

    return step
