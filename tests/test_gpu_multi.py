"""Multi-GPU parity: a world-2 NCCL run of the data-parallel PPO update (environment shards, gradient
all-reduce per minibatch issued from C inside cirs_ppo_learn) must reproduce the single-process update on the union
of the shards.  Needs two visible GPUs (``gpurun --gpus 2``); on a single-GPU box the test is skipped -- the log of
the two-GPU run is kept under profiles/."""
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


@pytest.mark.parametrize("world", [2])
def test_nccl_update_matches_single_process(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "DIST_OK" in p.stdout, p.stdout[-3000:] + "\n" + p.stderr[-3000:]
