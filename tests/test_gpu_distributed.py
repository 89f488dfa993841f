"""Multi-GPU parity (NCCL): skipped unless at least two GPUs are visible."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_distributed_residual_matches_oracle(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tests", "dist_parity_worker.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "DIST PARITY OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
