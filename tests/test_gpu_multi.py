"""GPU, >= 2 devices: the multi-GPU path (NCCL all-reduce of the charge density per step, of the Gram matrix
once, DEIM arg-max reduction) launched as torchrun does it -- one process per GPU.  Skipped on a single GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_ranks_nccl():
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(ngpu, 4)}",
           "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(here, "dist_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multi-GPU checks ok" in out.stdout
