"""Point-sharded ICP across 2 GPUs (fused peer exchange vs NCCL baseline vs the single-GPU answer). Needs >= 2 GPUs on
the box (gpurun --gpus 2); skipped otherwise."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.gpu
def test_sharded_icp_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(ROOT, "tools", "sharded_icp_check.py"), "--queries", "50000,1000000", "--voxels", "300000", "--iters", "15", "--reps", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 2 and all(l["ok"] and l["bit_identical_across_ranks"] for l in lines)
