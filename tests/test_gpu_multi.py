"""GPU (>= 2 devices): one box split over the GPUs with the NCCL transport equals the single-GPU run (tests/mgpu_check.py under torchrun);
all the GPUs of the box are used (2, 4 or 8).  On a one-GPU box the same decomposition is covered by tests/test_gpu_slab.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_slab_decomposition_over_nccl_matches_single_gpu():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs (the decomposition itself is tested on one GPU in test_gpu_slab.py)")
    world = 8 if ngpu >= 8 else 4 if ngpu >= 4 else 2
    pr = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                         "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_check.py")], capture_output=True, text=True, timeout=900)
    assert pr.returncode == 0 and "MGPU_OK" in pr.stdout, pr.stdout[-2000:] + pr.stderr[-4000:]
