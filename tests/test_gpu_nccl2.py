"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`): world_size-2 NCCL run
of the batch-sharded classifier forward — the all-gathered logits of the two ranks (through
torch.distributed and through the C-ABI sonet_allgather) equal the single-GPU forward of the full
batch BIT FOR BIT (parity definition 5). Launched like bench.py: torch.distributed.run."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_nccl_world2_gathered_logits_equal_single_gpu_bitwise():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "_nccl2_run.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "NCCL2_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
