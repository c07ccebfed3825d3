"""Rollout sharding across GPUs: needs >= 2 GPUs on the box, skipped otherwise. One process per GPU under
torchrun, NCCL all-gather inside mppib_solve; results must match the single-GPU solve (tools/mgpu_check.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_solve_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(ROOT, "tools", "mgpu_check.py")], capture_output=True, text=True, timeout=900)
    assert "MGPU_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
