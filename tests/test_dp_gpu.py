"""T4 on real GPUs (needs >= 2): the 2-rank data-parallel step (NCCL) equals the shard-wise single-process
expectation and leaves the replicas bit-identical.  Skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_two_rank_data_parallel_step_matches_shardwise_expectation():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "dp_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DP CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
