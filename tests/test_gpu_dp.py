"""Data-parallel map training with two ranks across densification steps (replicas must not diverge when the
map is densified).  Two processes share cuda:0 and talk through gloo, so the test runs on a single-GPU box; the
exchange kernels and the host logic are those of the NCCL path."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dp_two_ranks_cross_densification():
    env = dict(os.environ, DP_SAME_GPU="1", MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(ROOT, "tests", "tools", "train_dp_check.py")],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert {l["mode"] for l in lines if "mode" in l} == {"sparse", "dense", "auto"} and all(l["ok"] for l in lines), lines
