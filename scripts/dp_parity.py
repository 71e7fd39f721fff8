"""torchrun --nproc-per-node N scripts/dp_parity.py : N-rank batch-sharded training (bucketed, overlapped all-reduce; eager
and CUDA-graph replay) against the single-process step on the full batch. SURVEY.md 8e bar: rel-L2 <= 1e-5 in fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import pdequinox_b200 as pdeqx
from pdequinox_b200.train import TrainStep, GraphedTrainStep, shard_batch

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pdeqx.set_precision(os.environ.get("PREC", "fp32"))
rng = np.random.default_rng(5)
B = 8 * world
xs = [rng.standard_normal((B, 1, 64, 128)).astype(np.float32) for _ in range(3)]
ys = [rng.standard_normal((B, 1, 64, 128)).astype(np.float32) for _ in range(3)]
make = lambda: pdeqx.ClassicFNO(2, 1, 1, hidden_channels=32, num_modes=16, num_blocks=2, key=4)

ref = TrainStep(make(), lr=1e-2, distributed=False)          # single process, full batch
for x, y in zip(xs, ys):
    ref(x, y)
dp = TrainStep(make(), lr=1e-2, overlap=True)                               # this rank's shard, bucketed all-reduce
losses = [dp(shard_batch(x, rank, world), shard_batch(y, rank, world)).item() for x, y in zip(xs, ys)]
rel = lambda a, b: float(torch.linalg.norm(a - b) / torch.linalg.norm(b))
e1 = rel(dp.arena.params, ref.arena.params)

gp = TrainStep(make(), lr=1e-2, overlap=bool(int(os.environ.get("OVERLAP", "0"))))
snap = gp.arena.params.clone()
g = GraphedTrainStep(gp, torch.from_numpy(shard_batch(xs[0], rank, world)), torch.from_numpy(shard_batch(ys[0], rank, world)), warmup=1)
with torch.no_grad():
    gp.arena.params.copy_(snap); gp.mu.zero_(); gp.nu.zero_(); gp.adam_state.zero_()
for x, y in zip(xs, ys):
    g(torch.from_numpy(shard_batch(x, rank, world)).cuda(), torch.from_numpy(shard_batch(y, rank, world)).cuda())
e2 = rel(gp.arena.params, ref.arena.params)
torch.cuda.synchronize()
if rank == 0:
    print(f"dp_parity world={world}: eager rel-L2 {e1:.2e}, graph rel-L2 {e2:.2e}, loss {losses[-1]:.6f} vs {ref.loss.item():.6f}", flush=True)
    assert e1 <= 1e-5 and e2 <= 1e-5, (e1, e2)
g.graph.reset()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
