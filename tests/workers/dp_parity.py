"""Worker of tests/test_gpu_data_parallel.py (launched under torch.distributed.run with N ranks): N-rank batch-sharded
training -- eager, bucketed + overlapped all-reduce, and CUDA-graph replay -- against the single-process step on the
full batch. SURVEY.md 8(e) bar: parameters after three Adam steps agree to rel-L2 <= 1e-5 in fp32.

BACKEND=nccl needs one GPU per rank; BACKEND=gloo lets every rank compute on the same GPU (the exchange then goes through
host memory), which is how the single-GPU test box covers the N > 1 path of the CUDA kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist

import pdequinox_b200 as pdeqx
from pdequinox_b200.train import GraphedTrainStep, TrainStep, shard_batch

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
backend = os.environ.get("BACKEND", "nccl")
torch.cuda.set_device(local % torch.cuda.device_count())
if backend == "nccl":
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
else:
    dist.init_process_group("gloo")
pdeqx.set_precision(os.environ.get("PREC", "fp32"))
tol = 1e-5 if pdeqx.get_precision() == "fp32" else 2e-3
rng = np.random.default_rng(5)
B = 4 * world
xs = [rng.standard_normal((B, 1, 64, 128)).astype(np.float32) for _ in range(3)]
ys = [rng.standard_normal((B, 1, 64, 128)).astype(np.float32) for _ in range(3)]


def make():
    return pdeqx.ClassicFNO(2, 1, 1, hidden_channels=32, num_modes=16, num_blocks=2, key=4)


def rel(a, b):
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))


ref = TrainStep(make(), lr=1e-2, distributed=False)  # single process, full batch
for x, y in zip(xs, ys):
    ref(x, y)
errs = {}
# exchange paths: NCCL ranks default to the peer-memory kernel (reduce-scatter + Adam + all-gather over NVLink); the
# all-reduce path (the only one gloo has) and its bucketed, overlapped variant are kept covered explicitly
cases = [("allreduce", dict(exchange="allreduce"))]
if backend == "nccl":
    cases += [("peer", dict(exchange="peer")), ("allreduce+overlap", dict(overlap=True))]
for name, kw in cases:
    dp = TrainStep(make(), lr=1e-2, **kw)
    assert (dp.peer is not None) == (name == "peer"), name
    for x, y in zip(xs, ys):
        loss = dp(shard_batch(x, rank, world), shard_batch(y, rank, world))
    if dp.peer is not None:
        dp.peer.check()
    errs[name] = rel(dp.arena.params, ref.arena.params)
    assert abs(loss.item() - ref.loss.item()) <= 10 * tol * abs(ref.loss.item()), (name, loss.item(), ref.loss.item())
if backend == "nccl":  # graph capture of a step needs a capturable exchange: the peer-memory kernels (default) are plain kernels
    gp = TrainStep(make(), lr=1e-2)
    assert gp.peer is not None
    g = GraphedTrainStep(gp, torch.from_numpy(shard_batch(xs[0], rank, world)), torch.from_numpy(shard_batch(ys[0], rank, world)),
                         warmup=1)
    for x, y in zip(xs, ys):
        g(torch.from_numpy(shard_batch(x, rank, world)).cuda(), torch.from_numpy(shard_batch(y, rank, world)).cuda())
    gp.peer.check()
    errs["graph+peer"] = rel(gp.arena.params, ref.arena.params)
    g.graph.reset()
torch.cuda.synchronize()
if rank == 0:
    print("dp_parity world=%d backend=%s: %s" % (world, backend, ", ".join(f"{k} rel-L2 {v:.2e}" for k, v in errs.items())), flush=True)
bad = {k: v for k, v in errs.items() if not v <= tol}
sys.stdout.flush()
os._exit(1 if bad else 0)
