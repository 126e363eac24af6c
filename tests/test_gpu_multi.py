"""Multi-GPU parity (partition + exchange + fold on 2 ranks).  Needs >= 2 GPUs: -m gpu."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_rank_exchange_parity():
    n = 2
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
         "--master-port", "29517", os.path.join(ROOT, "tests", "multi_gpu_worker.py")],
        capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_PARITY OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
