"""Data-parallel equivalence on real GPUs: 2 ranks over NCCL against the single-process full-batch gradient
(tests/tools/dp_equiv.py).  Needs two visible GPUs (`gpurun --gpus 2`); skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_two_rank_cuda_gradients_equal_full_batch(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, SN_DP_MODE=mode, SN_DP_SIZE="256", SN_DP_PER_RANK="2")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "tools", "dp_equiv.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("DP_EQUIV")]
    from conftest import record

    record(f"dp2_cuda_equivalence[{mode}]", line[-1] if line else f"rc={r.returncode}")
    assert r.returncode == 0 and line and " OK " in line[-1], (r.stdout[-2000:], r.stderr[-3000:])
