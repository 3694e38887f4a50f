"""Desynchronised stepping (mq_step, csrc/chain.cu: step_desync): chains run through their cheap proposals at their own
pace and park when they draw one that needs travel-time tables; the tables of all parked chains are built in one full
launch.  Every chain draws from its own counter-based stream, so its trajectory must be the lock-step one
(MCMCEQ_DESYNC=0) BIT FOR BIT: counters, likelihoods, models, hypocentres, station corrections, sigmas and the decimated
records.  Each mode runs in its own subprocess because the switch is read once per process."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

SCRIPT = r"""
import sys, tempfile, numpy as np
sys.path.insert(0, %r)
import mcmc_eq_b200 as mq
from tests import inputs
out = {}
for name, n, tria in (("example2", 97, 0), ("example2", 8, 1), ("example", 40, 0)):
    d = tempfile.mkdtemp(prefix="mqds_")
    cfgp, pkp = inputs.materialise(name, d, j_max_start=30, j_max_main=100000, deci=7, tria=tria)
    cfg, pk = mq.read_config(cfgp), mq.Picks.read(pkp)
    smp = mq.Sampler(cfg, pk, n, 0, 21)
    smp.init_chains()
    recs = []
    for chunk in (7, 5, 7, 7, 4, 6):           # config strings: start phase, then the mixed main phase
        smp.step(chunk)
        r, lost = smp.drain()
        assert lost == 0
        recs += r
    smp.step(6, "QVRPBDMN")
    smp.step(5, "P")
    c, ll, rms = smp.stats()
    m = smp.get_models()
    key = f"{name}_{n}_{tria}_"
    out[key + "c"] = c; out[key + "ll"] = ll; out[key + "rms"] = rms
    for f in ("dim", "z", "vp", "vpvs", "eq", "pres", "sres", "noise", "origin"):
        out[key + f] = getattr(m, f)
    out[key + "rec_number"] = np.array([r["number"] for r in recs]); out[key + "rec_rms"] = np.array([r["rms"] for r in recs])
    out[key + "rec_chain"] = np.array([r["chain"] for r in recs])
    out[key + "launches"] = np.array(mq.lib().mq_launch_count())
    smp.close()
np.savez(sys.argv[1], **out)
"""


def _run(path, desync):
    env = dict(os.environ, MCMCEQ_DESYNC="1" if desync else "0")
    r = subprocess.run([sys.executable, "-c", SCRIPT % util.ROOT, path], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return dict(np.load(path))


def test_desynchronised_chains_follow_the_lock_step_trajectories(tmp_path):
    a = _run(str(tmp_path / "lockstep.npz"), False)
    b = _run(str(tmp_path / "desync.npz"), True)
    assert set(a) == set(b)
    for k in a:
        if k.endswith("launches"):
            continue
        assert np.array_equal(a[k], b[k], equal_nan=True), (k, int((a[k] != b[k]).sum()), a[k].size)
    # every iteration was accepted or rejected, and model proposals occurred
    c = b["example2_97_0_c"]
    assert (c[:, 17] + c[:, 18] == 7 + 5 + 7 + 7 + 4 + 6 + 6 + 5).all() and c[:, 1:17].reshape(97, 8, 2).sum((0, 2)).min() > 0
