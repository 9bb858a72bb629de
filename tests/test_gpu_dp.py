"""N ranks == 1 rank on the real nets over NCCL (needs >= 2 GPUs: `gpurun --gpus 2`); see tools/dp_check.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_equal_one_rank_on_real_nets():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dp_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DP_CHECK_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
