"""world_size-2 gloo run of the sharded path on CPU (host-side contract of the exchange)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_gloo_sharded_fold():
    from bytewax_b200 import _native as N

    N.build()
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29533", os.path.join(ROOT, "tests", "gloo_worker.py")],
        capture_output=True, text=True, timeout=300)
    assert "GLOO_PARITY OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
