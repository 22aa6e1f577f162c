"""Multi-GPU parity: one process per GPU over NCCL (tests/mgpu_worker.py).  The full-width run needs >= 2 GPUs; the
single-rank run drives the SAME protocol code (NCCL process group, gradient all-reduce, table-sharded step through the
owner-side peer reduce, global-negative InfoNCE, sharded top-k merge) on any one-GPU box."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_multi_gpu_paths_match_oracle():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "MULTIGPU_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_distributed_protocol_single_rank_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
    assert res.returncode == 0 and "MULTIGPU_OK world=1" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
