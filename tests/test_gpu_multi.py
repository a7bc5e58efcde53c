"""GPU (>= 2 devices): one box split over the GPUs with the NCCL all-reduce ghost exchange equals the single-GPU run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_slab_decomposition_matches_single_gpu():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    pr = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                         "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_check.py")], capture_output=True, text=True, timeout=900)
    assert pr.returncode == 0 and "MGPU_OK" in pr.stdout, pr.stdout[-2000:] + pr.stderr[-4000:]
