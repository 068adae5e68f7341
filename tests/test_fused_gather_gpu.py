"""Fused normalise + all-gather over NVLink peer memory (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_normalize_allgather_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577",
                          os.path.join(ROOT, "scripts", "test_fused_allgather.py")], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert "FUSED_ALLGATHER_OK=True" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
