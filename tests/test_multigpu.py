"""Runs tests/multigpu_parity.py on 2 GPUs when the box has them (skipped on single-GPU boxes)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["nccl", "p2p"])
@pytest.mark.parametrize("decomp", ["slab", "random"])
def test_two_gpu_parity(decomp, transport):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "multigpu_parity.py"), "--decomp", decomp]
    env = dict(os.environ, SSAM_HALO=transport)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("-- OK") == 2
    assert r.stdout.count("halo transport " + {"nccl": "nccl", "p2p": "peer-memory"}[transport]) == 2
