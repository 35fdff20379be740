"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): see tests/mgpu_worker.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_tile_sharding_and_update_broadcast_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29571",
                        os.path.join(root, "tests", "mgpu_worker.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:]
