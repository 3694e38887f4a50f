"""Posterior / ensemble level of parity (north_star level 3): free-running chains of this library (Philox
streams) cannot reproduce a reference chain sample by sample, but an ensemble of them must be statistically
indistinguishable from an ensemble of reference chains run with the same configuration.

Fixture: tests/golden/ensemble_ref_example2.npz -- 32 independent chains of the UNMODIFIED reference on Example2 (1500 + 2500
accepted models, every 100th written; tools/make_golden.py: ensemble_ref).  Here: 128 chains on the GPU, same config.
Compared at 1000 / 2000 / 3000 / 4000 accepted models: RMS, model dimension, mean hypocentre depth, mean Vp of the nuclei,
sigmas, size of the station corrections; and over the whole run: proposals tested, accept / reject counts per proposal
kind.  Tolerance: the difference of the ensemble means must be below 4.5 standard errors (both ensembles' spread, i.e. the
reference's own chain-to-chain variability sets the scale) or below 2 % of the value, whichever is larger."""
import os
import tempfile

import numpy as np
import pytest

from tests import inputs, util

pytestmark = pytest.mark.gpu

KINDS = ("noise", "P-vel", "Vp/Vs", "quake", "resid", "move", "birth", "death")


def _close(name, a, b, rel=0.02, nsig=4.5):
    ma, mb = a.mean(0), b.mean(0)
    se = np.sqrt(a.var(0, ddof=1) / len(a) + b.var(0, ddof=1) / len(b))
    tol = np.maximum(nsig * se, rel * np.abs(ma))
    bad = np.abs(ma - mb) > tol
    assert not np.any(bad), (name, ma, mb, se)
    return float(np.max(np.abs(ma - mb) / np.maximum(se, 1e-12)))


def test_gpu_ensemble_matches_reference_ensemble():
    import mcmc_eq_b200 as mq
    ref = dict(np.load(os.path.join(util.GOLDEN, "ensemble_ref_example2.npz")))
    d = tempfile.mkdtemp(prefix="mqens_")
    cfgp, pkp = inputs.materialise("example2", d, j_max_start=1500, j_max_main=2500, deci=100, true_random=1)
    cfg, pk = mq.read_config(cfgp), mq.Picks.read(pkp)
    n = 128
    smp = mq.Sampler(cfg, pk, n, 0, 4242)
    smp.init_chains()
    recs = [[] for _ in range(n)]
    for _ in range(400):
        smp.step(50)
        out, lost = smp.drain()
        assert lost == 0
        for r in out:
            recs[r["chain"]].append(r)
        counts, _ll, _rms = smp.stats()
        if (counts[:, 17] >= 4000).all():
            break
    counts, _ll, _rms = smp.stats()
    smp.close()
    assert (counts[:, 17] == 4000).all()                        # every chain ran to j_max_start + j_max_main accepted models
    assert all(len(r) == 40 for r in recs)
    idx = [9, 19, 29, 39]

    def traj(f):
        return np.array([[f(recs[c][i]) for i in idx] for c in range(n)], float)

    worst = {}
    worst["rms"] = _close("rms", ref["rms"][:, idx], traj(lambda r: r["rms"]))
    worst["dim"] = _close("dim", ref["dim"][:, idx].astype(float), traj(lambda r: r["dim"]), rel=0.05)
    worst["zmean"] = _close("zmean", ref["zmean"][:, idx], traj(lambda r: r["eq"][:, 2].mean()))
    worst["vp_mean"] = _close("vp_mean", ref["vp_mean"][:, idx], traj(lambda r: r["vp"].mean()))
    worst["res_rms"] = _close("res_rms", ref["res_rms"][:, idx][:, 1:],
                              traj(lambda r: np.sqrt(np.mean(np.square(np.stack([r["pres"], r["sres"]], 1)))))[:, 1:])
    # sigmas: the record holds them as 2*class+phase, the reference file as p0 p1 p2 p3 s0 s1 s2 s3
    order = [0, 2, 4, 6, 1, 3, 5, 7]
    for k in range(8):
        worst[f"noise{k}"] = _close(f"noise{k}", ref["noise"][:, idx, k], traj(lambda r: r["noise"][order[k]]))
    # bookkeeping of the whole run: counts[:, 1 + 2*slot] accepted, [2 + 2*slot] rejected, slots in the order of the cnt lines
    acc = counts[:, 1:17:2].astype(float)
    rej = counts[:, 2:17:2].astype(float)
    worst["tested"] = _close("tested", ref["tested"].astype(float)[:, None], counts[:, 0:1].astype(float))
    for k, name in enumerate(KINDS):
        worst["acc_" + name] = _close("acc_" + name, ref["acc"][:, k:k + 1].astype(float), acc[:, k:k + 1], rel=0.03)
        worst["rej_" + name] = _close("rej_" + name, ref["rej"][:, k:k + 1].astype(float), rej[:, k:k + 1], rel=0.03)
    print("ensemble: largest |difference of means| in standard errors:", {k: round(v, 2) for k, v in worst.items()})
    print("proposals per kind (accepted + rejected), reference vs here:",
          {name: (round(float((ref["acc"][:, k] + ref["rej"][:, k]).mean()), 1), round(float((acc[:, k] + rej[:, k]).mean()), 1))
           for k, name in enumerate(KINDS)})
