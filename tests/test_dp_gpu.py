"""Data parallelism on hardware: the k-rank product path -- per-rank shards, SyncBN statistics exchanged
over NVLink peer memory, NCCL gradient all-reduce inside the captured step -- equals the single-process step on the
concatenated batch.  Needs >= 2 GPUs on the box (skipped on a 1-GPU lease; tests/test_dp_gloo.py covers the same claim
with the oracle on CPU, world size 2)."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import REPO

pytestmark = pytest.mark.gpu


def _torchrun(script, nproc, env=None, timeout=600):
    e = dict(os.environ, **(env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(REPO, script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("steps", ["1", "3"])
def test_dp2_fixmatch_equals_single_process(steps):
    r = _torchrun("tools/dp_check.py", 2, {"DP_STEPS": steps})
    assert r.returncode == 0 and "DP equivalence OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp2_cps_and_eval_equal_single_process():
    r = _torchrun("tools/dp_check_semi.py", 2)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
