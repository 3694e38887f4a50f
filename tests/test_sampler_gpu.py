"""The batched Metropolis-Hastings loop (mq_init_chains / mq_step): invariants that hold for any RNG
stream -- bookkeeping, prior bounds, and that the incrementally maintained likelihood (single-event
updates, buffer flips) equals a from-scratch forward of the final state."""
import tempfile

import numpy as np
import pytest

from tests import inputs
from tests import fwd_helpers as fh

pytestmark = pytest.mark.gpu


def _sampler(name, n, seed=3, **over):
    import mcmc_eq_b200 as mq
    d = tempfile.mkdtemp(prefix="mqs_")
    cfgp, pkp = inputs.materialise(name, d, **over)
    cfg, pk = mq.read_config(cfgp), mq.Picks.read(pkp)
    return mq, cfg, pk, mq.Sampler(cfg, pk, n, 0, seed)


def test_start_models_are_valid_and_scored(oracle):
    mq, cfg, pk, smp = _sampler("example2", 16)
    smp.init_chains()
    m = smp.get_models()
    g = cfg.grid
    zmin, zmax = g.z0, g.z0 + (g.nz - 1) * g.h
    from tests.util import ptr
    for c in range(16):
        d = m.dim[c]
        assert 1 <= d <= g.nz
        assert (m.z[c, :d] >= zmin).all() and (m.z[c, :d] <= zmax).all()
        assert (m.vp[c, :d] > cfg.vpmin).all() and (m.vp[c, :d] < cfg.vpmax).all()
        inv = -abs(cfg.inv_control) if cfg.inv_control > 0 else cfg.inv_control
        assert oracle.ch_model_valid(int(d), ptr(m.z[c]), ptr(m.vp[c]), ptr(m.vpvs[c]), g.h, zmin, zmax, inv) == 0
    counts, ll, rms = smp.stats()
    assert np.isfinite(ll).all() and (rms > 0).all() and len(set(np.round(rms, 6))) > 8   # chains differ
    # likelihood of chain 0 against the oracle
    rmf, *_ = fh.oracle_forward(cfg, pk, m.z[0, :m.dim[0]], m.vp[0, :m.dim[0]], m.vpvs[0, :m.dim[0]], m.eq[0], m.pres[0], m.sres[0])
    assert abs(rms[0] - np.sqrt(rmf.sum() / pk.n_picks)) < 1e-4 * rms[0]
    st = smp.snapshot(0, 0)
    assert st["dim"] == m.dim[0] and st["eq"].shape == (pk.n_events, 3)
    smp.close()


@pytest.mark.parametrize("name", ["example2", "example"])
def test_incremental_state_equals_fresh_forward(name):
    mq, cfg, pk, smp = _sampler(name, 24, seed=5, deci=25)
    smp.init_chains()
    n_it = 120
    recs_all = []
    for _ in range(n_it // 20):
        smp.step(20)
        recs, lost = smp.drain()
        assert lost == 0
        recs_all += recs
    counts, ll, rms = smp.stats()
    assert (counts[:, 17] + counts[:, 18] == n_it).all()             # every iteration accepted or rejected
    assert (counts[:, 1:17:2].sum(1) == counts[:, 17]).all() and (counts[:, 2:17:2].sum(1) == counts[:, 18]).all()
    assert (counts[:, 0] <= n_it).all() and counts[:, 17].sum() > 0
    kinds = counts[:, 1:17].reshape(24, 8, 2).sum((0, 2))
    assert kinds[3] > 0 and kinds[0] > 0                               # start phase string is "QN"
    mf_inc = None
    # from-scratch forward of the final states must reproduce the maintained likelihood
    m = smp.get_models()
    mf, origin = smp.forward(3)
    counts2, ll2, rms2 = smp.stats()
    assert np.allclose(ll2, ll, rtol=2e-5, atol=1e-3), np.abs(ll2 - ll).max()
    assert np.allclose(rms2, rms, rtol=1e-5)
    assert np.allclose(origin, m.origin, atol=2e-5)
    for r in recs_all:
        assert r["code"] in "QN" and (r["number"] + 1) % 25 == 0 and r["kind"] == 0
    best = smp.snapshot(0, 1)
    assert best["rms"] <= rms[0] + 1e-9
    smp.close()


def test_model_proposals_flip_buffers_consistently():
    """All eight proposal kinds, main-phase mix; then again the from-scratch check."""
    mq, cfg, pk, smp = _sampler("example2", 32, seed=9, j_max_start=0, j_max_main=100000)
    smp.init_chains()
    smp.step(60, "QVRPBDMN")
    counts, ll, rms = smp.stats()
    tried = counts[:, 1:17].reshape(32, 8, 2).sum((0, 2))
    assert (tried > 0).all()                                           # N P V Q R M B D all drawn
    m = smp.get_models()
    assert (m.dim >= 1).all() and (m.dim < cfg.max_dim).all()
    _mf, _ = smp.forward(3)
    _c, ll2, rms2 = smp.stats()
    assert np.allclose(ll2, ll, rtol=2e-5, atol=1e-3) and np.allclose(rms2, rms, rtol=1e-5)
    # station corrections keep zero mean under scor_flag = 0 (src/mcmc_eq.c:910-916)
    assert np.abs(m.pres.mean(1)).max() < 1e-4 and np.abs(m.sres.mean(1)).max() < 1e-4
    smp.close()


def test_same_seed_same_chains():
    mq, cfg, pk, a = _sampler("example2", 8, seed=11)
    a.init_chains(); a.step(40)
    ra = a.stats()
    _mq, _cfg, _pk, b = _sampler("example2", 8, seed=11)
    b.init_chains(); b.step(40)
    rb = b.stats()
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
    a.close(); b.close()


def test_linear_gradient_models_keep_their_end_nuclei():
    """Config line 29 = 1: start models carry two extra nuclei at the top and the bottom of the model
    (src/mcmc_eq.c:577-588) which the move and death arms never touch (:990-998, :1059-1067); the maintained
    likelihood still equals a from-scratch forward."""
    mq, cfg, pk, smp = _sampler("example2", 32, seed=13, j_max_start=0, j_max_main=100000, tria=1)
    g = cfg.grid
    zmin, zmax = np.float32(g.z0), np.float32(g.z0 + (g.nz - 1) * g.h)
    smp.init_chains()
    m = smp.get_models()
    assert (m.dim >= 3).all() and (m.z[:, 0] == zmin).all() and (m.z[:, 1] == zmax).all()
    smp.step(80, "QVRPBDMN")
    counts, ll, rms = smp.stats()
    tried = counts[:, 1:17].reshape(32, 8, 2).sum((0, 2))
    assert (tried > 0).all() and counts[:, 17].sum() > 0
    m = smp.get_models()
    assert (m.dim >= 2).all() and (m.z[:, 0] == zmin).all() and (m.z[:, 1] == zmax).all()
    _mf, _ = smp.forward(3)
    _c, ll2, rms2 = smp.stats()
    assert np.allclose(ll2, ll, rtol=2e-5, atol=1e-3) and np.allclose(rms2, rms, rtol=1e-5)
    smp.close()
