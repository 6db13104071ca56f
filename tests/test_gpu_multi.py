"""GPU, >= 2 devices: sharded solvers over NCCL against the unsharded oracle (skipped on single-GPU boxes)."""

import pathlib
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_two_gpu_sharded_steps_match_oracle(cuda_device):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", str(ROOT / "scripts" / "multigpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert res.returncode == 0 and "MULTIGPU_CHECK_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
