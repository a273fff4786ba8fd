"""Slab-decomposed run on >= 2 GPUs equals the single-GPU run (skipped on a 1-GPU box)."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,n_grid,n_side", [(2, 64, 32), (4, 128, 64), (8, 128, 64)])
def test_slab_decomposition_matches_single_gpu(world, n_grid, n_side):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29400 + world), str(ROOT / "tests" / "multi_gpu_worker.py"), str(n_grid), str(n_side)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
