"""Multi-rank parity of the CUDA path: a 2-rank batch-sharded job must land on the same parameters as the
single-process step on the full batch. With two GPUs the ranks exchange gradients over NCCL (eager, overlapped buckets
and CUDA-graph replay); on a one-GPU box both ranks compute on that GPU and exchange through gloo."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_two_rank_training_matches_single_process(cuda, precision):
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    env = dict(os.environ, BACKEND=backend, PREC=precision)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "workers", "dp_parity.py")]
    out = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "dp_parity world=2" in out.stdout
