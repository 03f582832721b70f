"""Frame-parallel exchange on real GPUs (needs >= 2 on the box; skipped otherwise): the library's own exchange over
NVLink / NVSwitch on a symmetric buffer -- two-shot (fsgs_exchange_rows + expansion) and one-shot (the rank sum folded
into the expansion kernel, fsgs_compact_grad_expand_peers) -- inside the fused backward must give the gradients
ncclAllReduce gives and the sum of the frames' full gradients, bit-identical on every rank."""
import json
import os
import re
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("P", [20003, 1000])
def test_nvlink_exchange_matches_nccl_and_the_local_sum(P):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs on one box")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29533 + P % 89), os.path.join(ROOT, "tests", "mgpu_worker.py"), str(P)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    # (the ranks print concurrently: two records can end up on one line)
    lines = [json.loads(m) for m in re.findall(r"MGPU (\{.*?\})", p.stdout)]
    assert p.returncode == 0 and len(lines) == world, p.stdout[-1500:] + p.stderr[-3000:]
    for r in lines:
        assert r["nvlink_vs_local_sum"] < 1e-5 and r["nccl_vs_local_sum"] < 1e-5, r
        assert r["nvlink_two_shot_vs_local_sum"] < 1e-5 and r["nvlink_one_shot_vs_local_sum"] < 1e-5, r
        assert r["nvlink_pull_gather_vs_local_sum"] < 1e-5, r
        assert r["nvlink_vs_nccl"] < 1e-5, r                    # (two separate backward runs: float atomics order)
        assert r["nvlink_pose_local"] < 1e-5, r                 # pose gradients stay local
        assert r["bit_identical_across_ranks"], r
