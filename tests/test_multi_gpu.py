"""-m gpu, needs >= 2 GPUs (skipped otherwise): trajectory sharding over NCCL gives the same training as one GPU."""
import json
import os
import subprocess
import sys

import pytest
import torch as pt

from conftest import ROOT

pytestmark = pytest.mark.gpu


def run(world, loss):
    script = os.path.join(ROOT, "tools", "ddp_check.py")
    if world == 1:
        cmd = [sys.executable, script, loss]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), script, loss]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("DDP_RESULT ")][-1]
    return json.loads(line[len("DDP_RESULT "):])


@pytest.mark.parametrize("loss", ["log-variance", "relative_entropy"])
def test_two_ranks_match_one_rank(loss):
    if pt.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    a, b = run(1, loss), run(2, loss)
    assert b["world"] == 2 and b["K_local"] == 500
    for x, y in zip(a["loss_log"], b["loss_log"]):
        assert abs(x - y) <= 2e-5 * abs(x), (a, b)      # identical trajectories; only summation order differs
    assert abs(a["theta_sum"] - b["theta_sum"]) < 1e-3 * max(1.0, abs(a["theta_sum"]))
