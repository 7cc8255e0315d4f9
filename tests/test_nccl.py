"""The NCCL paths against the single-GPU result, on >= 2 real GPUs: tests/nccl_worker.py under torchrun.
Skipped on a one-GPU box (the gloo tests in test_distributed.py cover the host logic there)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4])
def test_nccl_paths_match_single_gpu(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "nccl_worker.py")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-6000:]
    assert res.stdout.count("NCCL-PARITY-OK") == world, res.stdout[-6000:]
