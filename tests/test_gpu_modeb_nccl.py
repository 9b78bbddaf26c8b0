"""Mode B over NCCL on two real GPUs (skipped on a one-GPU box): tests/modeb_nccl_worker.py under
torch.distributed.run. The one-GPU stand-in for the same data flow is tests/test_gpu_modeb.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_mode_b_over_nccl_equals_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(here, "modeb_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "0 mismatching fields" in r.stdout
