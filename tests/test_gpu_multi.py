"""Spawns the multi-rank parity check when at least 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_equals_torus_on_2_gpus():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_single_rank_slab_paths():
    """world == 1 degenerates to local ghost copies: same result as the torus kernels."""
    import pyca_b200
    from pyca_b200.slab import GolSlabCUDA

    slab = GolSlabCUDA(300, 1000, 0, 1, ghost=5, device="cuda")
    slab.random_fill(seed=3)
    slab.run(17)
    ref = pyca_b200.PackedWorld(300, 1000, "cuda")
    ref.random_fill(seed=3)
    ref.step(0b1100, 0b1000, 17)
    assert torch.equal(slab.owned()[:, :ref.wpr], ref.bits[:, :ref.wpr])
