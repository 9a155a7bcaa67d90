"""On-hardware N > 1 correctness (runs when >= 2 GPUs are visible): two NCCL ranks, sharded == unsharded for the
rendered images (bit for bit), the training iteration's losses and the CNN weights after two optimiser steps."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_nccl_ranks_match_the_unsharded_run(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_nccl_worker.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    ranks = "\n".join(line for line in (out.stdout + out.stderr).splitlines() if line.startswith("[rank"))
    assert out.returncode == 0 and "NCCL_WORKER_OK" in out.stdout, ranks[-3000:] or (out.stdout + out.stderr)[-3000:]
