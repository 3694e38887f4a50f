"""Parity of the batched forward (mq_forward: rasterise -> tables -> lookup -> residual sums) with the
oracle and with cal_fit_newx of the compiled reference (fixtures).  Tolerances: per-pick prediction
1e-4 s absolute, class sums 2e-5 relative, origin time 1e-4 s."""
import os
import tempfile

import numpy as np
import pytest

from tests import util, inputs
from tests import fwd_helpers as fh

pytestmark = pytest.mark.gpu


def _setup(name, n_chains, **over):
    import mcmc_eq_b200 as mq
    d = tempfile.mkdtemp(prefix="mqfw_")
    cfgp, pkp = inputs.materialise(name, d, **over)
    cfg, pk = mq.read_config(cfgp), mq.Picks.read(pkp)
    return mq, cfg, pk, mq.Sampler(cfg, pk, n_chains, 0, 1)


def _check_sums(mf, ref, rel=2e-5):
    assert np.allclose(mf, ref, rtol=rel, atol=1e-6), (mf, ref, np.abs(mf - ref) / np.maximum(ref, 1e-9))


@pytest.mark.parametrize("name", ["example2", "example"])
def test_forward_matches_reference_fixtures(name):
    d = np.load(os.path.join(util.GOLDEN, "forward_ref.npz"))
    n = int(d[f"{name}_n"])
    mq, cfg, pk, smp = _setup(name, n)
    states = [{k: d[f"{name}_{i}_{k}"] for k in ("z", "vp", "vpvs", "eq", "pres", "sres", "noise")} for i in range(n)]
    m = fh.fill_models(smp.new_models(32), states)
    mf, origin = smp.forward_host(m, 3)
    for i in range(n):
        _check_sums(mf[i], d[f"{name}_{i}_mf"])
        assert np.abs(origin[i] - d[f"{name}_{i}_origin"]).max() < 1e-4
    # receiver rows 1,2 of the P table of state 0 against the reference's own table
    tab = smp.table(0, 1)
    ref = d[f"{name}_0_tabP_rows12"]
    assert (np.abs(tab[1:3] - ref) <= util.eikonal_tol(ref)).all()
    # the rows the device keeps (mq_get_rows) are rows of that table
    rows, idx = smp.rows(0, 1)
    assert len(idx) >= 2 and (np.diff(idx) > 0).all()
    assert (np.abs(rows - tab[idx]) <= util.eikonal_tol(tab[idx])).all()
    smp.close()


def test_tria_forward_matches_reference_fixtures(oracle):
    """Config line 29 = 1 (linear gradients between the nuclei, src/misfit.c:217-250) against cal_fit_newx of the compiled
    reference with TRIA = 1 (tests/golden/forward_ref_tria.npz) and per pick against the oracle."""
    d = np.load(os.path.join(util.GOLDEN, "forward_ref_tria.npz"))
    n = int(d["n"])
    mq, cfg, pk, smp = _setup("example2", n, tria=1)
    states = [{k: d[f"{i}_{k}"] for k in ("z", "vp", "vpvs", "eq", "pres", "sres", "noise")} for i in range(n)]
    mf, origin = smp.forward_host(fh.fill_models(smp.new_models(32), states), 3)
    for i, s in enumerate(states):
        _check_sums(mf[i], d[f"{i}_mf"])
        assert np.abs(origin[i] - d[f"{i}_origin"]).max() < 1e-4
        _rmf, _rorg, rres, rtp = fh.oracle_forward(cfg, pk, s["z"], s["vp"], s["vpvs"], s["eq"], s["pres"], s["sres"])
        res, tp = smp.predictions(i)
        assert np.abs(tp - rtp).max() < 1e-4 and np.abs(res - rres).max() < 1e-4
    for ps, key in ((1, "0_tabP_rows12"), (2, "0_tabS_rows12")):
        tab = smp.table(0, ps)
        assert (np.abs(tab[1:3] - d[key]) <= util.eikonal_tol(d[key])).all()
    smp.close()


def test_forward_matches_oracle_per_pick(oracle):
    mq, cfg, pk, smp = _setup("example2", 6)
    rng = np.random.default_rng(21)
    states = fh.random_states(rng, cfg, pk, 6, "posterior") 
    states[1] = fh.random_states(rng, cfg, pk, 1, "contrast", 4)[0]
    states[2] = fh.random_states(rng, cfg, pk, 1, "gradient", 1)[0]     # single layer: homogeneous half space
    smp.set_models(fh.fill_models(smp.new_models(32), states))
    mf, origin = smp.forward(3)
    for c, s in enumerate(states):
        rmf, rorg, rres, rtp = fh.oracle_forward(cfg, pk, s["z"], s["vp"], s["vpvs"], s["eq"], s["pres"], s["sres"])
        _check_sums(mf[c], rmf)
        assert np.abs(origin[c] - rorg).max() < 1e-4
        res, tp = smp.predictions(c)
        assert np.abs(tp - rtp).max() < 1e-4 and np.abs(res - rres).max() < 1e-4
    # calct = 0 re-uses the tables: same sums; calct = 1/2 rebuild one phase only: still the same model -> same sums
    for calct in (0, 1, 2):
        mf2, _ = smp.forward(calct)
        assert np.array_equal(mf2, mf)
    smp.close()


def test_full_table_matches_oracle(oracle):
    mq, cfg, pk, smp = _setup("example2", 2)
    rng = np.random.default_rng(22)
    states = fh.random_states(rng, cfg, pk, 2, "posterior")
    smp.set_models(fh.fill_models(smp.new_models(32), states))
    s = states[1]
    _mf, _o, _r, _t, tabs = fh.oracle_forward(cfg, pk, s["z"], s["vp"], s["vpvs"], s["eq"], s["pres"], s["sres"], True)
    for phase in (1, 2):
        tab = smp.table(1, phase)          # ttt[j][iz][i], the reference's layout
        assert tab.shape == tabs[phase - 1].shape
        assert (np.abs(tab - tabs[phase - 1]) <= util.eikonal_tol(tabs[phase - 1])).all()
    smp.close()


def test_out_of_table_and_bad_station_correction(oracle):
    """Events outside the table give the reference's 1e30 sentinel (src/interpol.c:64-65); a pick that points
    to an invalid station correction is an error code, not an exit(0) (src/misfit.c:93)."""
    mq, cfg, pk, smp = _setup("example2", 1)
    rng = np.random.default_rng(23)
    s = fh.random_states(rng, cfg, pk, 1)[0]
    g = cfg.grid
    s["eq"][0, 2] = g.z0 + (g.nz - 1) * g.h          # deepest node: iz1 >= nz-1 -> 1e30
    smp.set_models(fh.fill_models(smp.new_models(32), [s]))
    mf, origin = smp.forward(3)
    rmf, rorg, *_ = fh.oracle_forward(cfg, pk, s["z"], s["vp"], s["vpvs"], s["eq"], s["pres"], s["sres"])
    # every pick of event 0 predicts 1e30: its origin time is -1e30 and its de-meaned residuals vanish
    assert origin[0, 0] < -9e29 and rorg[0] < -9e29
    assert np.abs(origin[0, 1:] - rorg[1:]).max() < 1e-4
    # ... in exact arithmetic; in the reference's FP32 the squares overflow: infinite misfit in the event's classes
    assert np.isinf(rmf).any() and np.array_equal(np.isinf(mf[0]), np.isinf(rmf))
    fin = np.isfinite(rmf)
    _check_sums(mf[0][fin], rmf[fin])
    s["eq"][0, 2] = 5.0
    s["pres"][3] = -99999.0
    m = fh.fill_models(smp.new_models(32), [s])
    with pytest.raises(mq.MqError) as e:
        smp.forward_host(m, 0)
    assert e.value.code == -5
    smp.close()


def test_straight_ray_branch(oracle):
    """eikonal = 0 (config line 32): half-space straight rays, src/misfit.c:90,108."""
    mq, cfg, pk, smp = _setup("example2", 3, eikonal=0)
    rng = np.random.default_rng(24)
    states = fh.random_states(rng, cfg, pk, 3)
    smp.set_models(fh.fill_models(smp.new_models(32), states))
    mf, origin = smp.forward(3)
    for c, s in enumerate(states):
        rmf, rorg, *_ = fh.oracle_forward(cfg, pk, s["z"], s["vp"], s["vpvs"], s["eq"], s["pres"], s["sres"])
        _check_sums(mf[c], rmf, 5e-6)
        assert np.abs(origin[c] - rorg).max() < 2e-5
    smp.close()


def test_host_driven_loop_with_table_backup(oracle):
    """The function seam used the way the reference's main uses cal_fit_newx: tables are those of
    the last call with calct != 0; mq_tables_save / mq_tables_restore stand for the reference's backup and restore of them
    (src/mcmc_eq.c:856,1161,1171).  Sequence: model A (calct 3) -> save -> model B (calct 3, "rejected") -> restore ->
    hypocentre change on A (calct 0) must score against A's tables."""
    import mcmc_eq_b200 as mq
    rng = np.random.default_rng(21)
    d = tempfile.mkdtemp(prefix="mqf_")
    cfgp, pkp = inputs.materialise("example2", d)
    cfg, pk = mq.read_config(cfgp), mq.Picks.read(pkp)
    smp = mq.Sampler(cfg, pk, 1, 0, 1)
    a, b = fh.random_states(rng, cfg, pk, 2)
    mfa, _ = smp.forward_host(fh.fill_models(smp.new_models(32), [a]), 3)
    smp.tables_save()
    smp.forward_host(fh.fill_models(smp.new_models(32), [b]), 3)
    smp.tables_restore()
    a2 = dict(a)
    a2["eq"] = a["eq"].copy()
    a2["eq"][5] += np.float32([0.7, -0.4, 1.1])
    mf2, org2 = smp.forward_host(fh.fill_models(smp.new_models(32), [a2]), 0)
    ref, rorg, *_ = fh.oracle_forward(cfg, pk, a2["z"], a2["vp"], a2["vpvs"], a2["eq"], a2["pres"], a2["sres"])
    assert np.allclose(mf2[0], ref, rtol=2e-5) and np.allclose(org2[0], rorg, atol=2e-5)
    assert not np.allclose(mf2[0], mfa[0], rtol=1e-6)
    smp.close()


def test_fine_grid_takes_the_generic_kernel(oracle):
    """A 0.1 km grid (601 depth nodes: the plane no longer fits the fast kernel's shared memory) runs through the generic
    one-lane-per-solve kernel with its time fields in a memory-capped global scratch; same parity bar."""
    import mcmc_eq_b200 as mq
    rng = np.random.default_rng(8)
    d = tempfile.mkdtemp(prefix="mqfine_")
    cfgp, pkp = inputs.materialise("example2", d, h=0.1, nx=481, ny=481, nz=601)
    cfg, pk = mq.read_config(cfgp), mq.Picks.read(pkp)
    assert cfg.grid.nz == 601 and abs(cfg.grid.h - 0.1) < 1e-6
    smp = mq.Sampler(cfg, pk, 1, 0, 1)
    st = fh.random_states(rng, cfg, pk, 1, max_layers=8)
    mf, org = smp.forward_host(fh.fill_models(smp.new_models(32), st), 3)
    _res, tp = smp.predictions(0)
    ref, rorg, _r, rtp = fh.oracle_forward(cfg, pk, st[0]["z"], st[0]["vp"], st[0]["vpvs"], st[0]["eq"], st[0]["pres"], st[0]["sres"])
    assert np.abs(tp - rtp).max() < 1e-4
    assert np.allclose(mf[0], ref, rtol=2e-5) and np.allclose(org[0], rorg, atol=2e-5)
    smp.close()
