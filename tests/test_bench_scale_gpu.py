"""Parity of the BENCHED path at bench scale, directly against the CPU oracle: 1024 chains of the
synthetic workloads bench.py times -- synth.workload(200, 50) = BASELINE configs[2] and a 1024-chain shard of
synth.workload(2000, 100) = configs[3] -- through mq_forward_host(calct = 3), the call bench.py's e2e figure goes through.
A launch of 1024 chains x 2 phases x 62 source depths = 3968 warp-tasks takes the pipelined kernel (eik_pipe_kernel: box
phase on shared-memory slices, march in tensor memory), which the smaller fixtures never reach; mq_profile_kernels proves
which kernel ran.  Checked on >= 32 sampled chains: every pick's prediction (1e-4 s), the eight class sums (2e-5
relative), origin times (1e-4 s) and every node of the stored receiver rows (max(1e-4 s, 2e-6 T)), i.e. the reference's
setup_table_new output (src/misfit.c:270-289) and cal_fit_newx (src/misfit.c:83-153) on identical models."""
import numpy as np
import pytest

from tests import fwd_helpers as fh
from tests import util

pytestmark = pytest.mark.gpu


def _chain_states(mq, synth, cfg, pk, n, seed):
    """Valid chain states the way a run produces them: start models + a few mixed iterations, then half of the chains
    replaced by posterior-like random models (9-20 layers: the regime where 94 % of the solves re-discretise the source
    region)."""
    smp = mq.Sampler(cfg, pk, n, 0, seed)
    smp.init_chains()
    smp.step(6, "QVRPBDMN")
    m = smp.get_models()
    rng = np.random.default_rng(seed)
    rnd = fh.random_states(rng, cfg, pk, n // 2, "posterior", 20)
    for c, s in enumerate(rnd):
        d = len(s["z"])
        m.dim[2 * c] = d
        m.z[2 * c, :d], m.vp[2 * c, :d], m.vpvs[2 * c, :d] = s["z"], s["vp"], s["vpvs"]
    return smp, m


def _check(cfg, pk, smp, m, mf, origin, chains):
    worst = dict(dt=0.0, rel=0.0, row=0.0, org=0.0)
    for c in chains:
        d = int(m.dim[c])
        rmf, rorg, _res, rpred, tabs = fh.oracle_forward(cfg, pk, m.z[c, :d], m.vp[c, :d], m.vpvs[c, :d], m.eq[c], m.pres[c],
                                                          m.sres[c], want_tables=True)
        _r, tpred = smp.predictions(c)
        ok = np.abs(rpred) < 1e20                       # picks inside the table (1e30 sentinel of src/interpol.c:64-65 otherwise)
        assert ok.any() and (np.abs(tpred[ok] - rpred[ok]) <= 1e-4).all(), (c, np.abs(tpred[ok] - rpred[ok]).max())
        assert ((np.abs(tpred) > 1e20) == ~ok).all()
        worst["dt"] = max(worst["dt"], float(np.abs(tpred[ok] - rpred[ok]).max()))
        fin = np.isfinite(rmf)
        assert (np.isfinite(mf[c]) == fin).all()
        assert np.allclose(mf[c][fin], rmf[fin], rtol=2e-5, atol=1e-6), (c, mf[c], rmf)
        nz_ = fin & (rmf > 0)
        if nz_.any():
            worst["rel"] = max(worst["rel"], float((np.abs(mf[c][nz_] - rmf[nz_]) / rmf[nz_]).max()))
        if ok.all():
            assert np.abs(origin[c] - rorg).max() <= 1e-4
            worst["org"] = max(worst["org"], float(np.abs(origin[c] - rorg).max()))
        for ph in (1, 2):
            rows, idx = smp.rows(c, ph)
            ref = tabs[ph - 1][idx]
            err = np.abs(rows - ref)
            assert (err <= util.eikonal_tol(ref)).all(), (c, ph, float(err.max()))
            worst["row"] = max(worst["row"], float(err.max()))
    return worst


@pytest.mark.parametrize("events,stations", [(200, 50), (2000, 100)])
def test_pipelined_kernel_at_bench_scale_matches_the_oracle(events, stations):
    import mcmc_eq_b200 as mq
    from mcmc_eq_b200 import synth
    n = 1024
    cfg, pk, truth = synth.workload(events, stations, 33, 0, j_max_start=0, j_max_main=2**30, deci=2**30)
    # the synthetic picks (made by this library) agree with the oracle's prediction, which bench.py's reference arm uses
    ref_pred = fh.oracle_forward(cfg, pk, truth["z"], truth["vp"], truth["vpvs"], truth["eq"], truth["pres"], truth["sres"])[3]
    assert np.abs(ref_pred - truth["tpred"]).max() <= 1e-4
    smp, m = _chain_states(mq, synth, cfg, pk, n, 5)
    smp.profile(True)
    mf, origin = smp.forward_host(m, 3)
    _ms, launches, per_launch = smp.profile(False)
    kernels = smp.profile_kernels()
    assert launches == 1 and per_launch == 2 * n * cfg.grid.nz
    assert list(kernels) == ["eik_pipe_kernel"], kernels     # the benched kernel, not the fused one
    rng = np.random.default_rng(9)
    chains = sorted(int(c) for c in rng.choice(n, size=32 if events <= 200 else 12, replace=False))
    chains += [0, 1, n - 2, n - 1]                           # both kinds of state, first and last warp-tasks
    worst = _check(cfg, pk, smp, m, mf, origin, chains)
    print(f"bench-scale parity ({events} events x {stations} stations, {len(chains)} chains): max |dT| pick {worst['dt']:.2e} s, "
          f"row {worst['row']:.2e} s, class sum {worst['rel']:.2e} rel., origin {worst['org']:.2e} s")
    smp.close()
