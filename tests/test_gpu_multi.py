"""Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus 2`): scripts/multi_rank_check.py under torchrun -- read
shards per rank, NCCL all-gather of the junction / deletion / insertion / fusion records, every rank equals the oracle."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_two_ranks_allgather_equals_oracle():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29511", os.path.join(ROOT, "scripts", "multi_rank_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "MULTI-RANK OK" in r.stdout, (r.stdout[-3000:] + "\n" + r.stderr[-3000:])
