"""N > 1 on real GPUs (skipped on a one-GPU box): tests/mp_worker.py under torch.distributed.run, two ranks."""
import json
import os
import subprocess
import sys

import pytest

from tests import util

pytestmark = pytest.mark.gpu


def test_two_gpus_nccl_posterior_tempering_and_shard_invariance():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(util.ROOT, "tests", "mp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=util.ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.split("\n") if ln.startswith("MPRESULT ")][0]
    res = json.loads(line[len("MPRESULT "):])
    assert len(res) == 2
    for rank, ok in enumerate(res):
        bad = [k for k, v in ok.items() if not v]
        assert not bad, (rank, bad)
    assert "shard_invariant" in res[0]
