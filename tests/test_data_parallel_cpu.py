"""The N>1 host logic on CPU: world_size 2 over gloo. Each rank takes its shard of the global batch
(pdequinox_b200.train.shard_batch), computes the shard's mean-loss gradient (oracle arithmetic stands in for the
CUDA kernels here), the flat arena is summed with the same all-reduce call the GPU path uses, and grad/world equals
the full-batch gradient."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nn as onn
from pdequinox_b200.train import allreduce_gradients_, shard_batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    rng = np.random.default_rng(0)
    p = {k: v.astype(np.float64) for k, v in onn.init_fno(rng, 1, 1, 1, hidden_channels=4, num_modes=3, num_blocks=2).items()}
    x = rng.standard_normal((8, 1, 16))
    y = rng.standard_normal((8, 1, 16))
    return p, x, y


def _flat(g, keys):
    return np.concatenate([np.asarray(g[k], dtype=np.float64).ravel() for k in keys])


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p, x, y = _problem()
    xs, ys = shard_batch(x, rank, world), shard_batch(y, rank, world)
    loss, g = onn.fno_loss_and_grad(p, xs, ys, (3,), 2)
    arena = torch.from_numpy(_flat(g, list(p)))
    loss_t = torch.tensor([loss], dtype=torch.float64)
    allreduce_gradients_(arena, loss_t, world)
    if rank == 0:
        out.put((arena.numpy() / world, float(loss_t)))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    grads, loss = out.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p, x, y = _problem()
    loss_full, g_full = onn.fno_loss_and_grad(p, x, y, (3,), 2)
    assert loss == pytest.approx(loss_full, rel=1e-12)
    assert np.allclose(grads, _flat(g_full, list(p)), rtol=1e-10, atol=1e-14)


def test_shard_batch_partitions_by_sample():
    x = np.arange(24).reshape(8, 3)
    parts = [shard_batch(x, r, 4) for r in range(4)]
    assert np.array_equal(np.concatenate(parts), x)
    with pytest.raises(ValueError):
        shard_batch(x, 0, 3)
