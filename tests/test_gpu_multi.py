"""-m gpu, needs >= 2 GPUs (skipped otherwise): the NCCL path of the sequence-sharded training step."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.timeout(600)
def test_two_rank_nccl_training_step_equals_whole_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(HERE, "dist_train_step.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=os.path.dirname(HERE))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "OK" in r.stdout
