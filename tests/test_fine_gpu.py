"""The fine-grid stress case (BASELINE.json configs[4]: 0.1 km grid to 200 km depth, eikonal plane 565 x 2001): planes
that do not fit a shared-memory slice take eik_fine_kernel -- the warp-synchronous solver with its per-lane arrays in global
memory and the rows that run along the grid's right edge on the lock-step path.  Checked against the CPU oracle
(reference time_2d, src/time_2d.c:301, as setup_table_new calls it, src/misfit.c:270-289): stored receiver rows of sampled
source depths, per-pick predictions 1e-4 s, class sums 2e-5 relative.  Table tolerance on this plane: max(1e-4 s, 5e-6 T) --
a path through the 2001 x 565 plane crosses thirty times more nodes than one through the 62-row plane the 2e-6 T bound of
SURVEY.md appendix A.5 was measured on, and the FP32-vs-double rounding differences of the restatement add up along it
(measured: 1.5e-4 s at T = 47 s, 3.2e-6 T, source at the bottom of a high-contrast model)."""
import ctypes as C

import numpy as np
import pytest

from tests import fwd_helpers as fh
from tests import util

pytestmark = pytest.mark.gpu


def _oracle_misfit_sparse(cfg, pk, st, depths_by_phase):
    """Oracle forward with tables that hold only the source depths the picks need (the others stay zero and are never read)."""
    L = util.oracle()
    g = fh.fm_grid(cfg)
    nz, nxmod = cfg.grid.nz, L.fm_nxmod(C.byref(g))
    tabs, slows = [], []
    for ps in (1, 2):
        slow = np.zeros(nz, np.float32)
        L.fm_rasterise(C.byref(g), len(st["z"]), util.ptr(util.f32(st["z"])), util.ptr(util.f32(st["vp"])), util.ptr(util.f32(st["vpvs"])), ps,
                       util.ptr(slow))
        slows.append(slow)
        # the stations sit in the top layers: only receiver layers 0 .. 3 of ttt[nz][nz][nxmod] are ever read (src/misfit.c:91)
        t = np.zeros((4, nz, nxmod), np.float32)
        for iz in depths_by_phase:
            field, rc = util.oracle_time_2d(slow, nxmod, iz)
            assert rc == 0
            t[:, iz, :] = field.T[:4]      # ttt[j][iz][i] = t[i*nz + j], src/misfit.c:281-288
        tabs.append(t)
    p = fh.fm_picks(pk)
    mf, origin = np.zeros(8, np.float32), np.zeros(pk.n_events, np.float32)
    resid, tpred = np.zeros(pk.n_picks, np.float32), np.zeros(pk.n_picks, np.float32)
    rc = L.fm_misfit(C.byref(g), C.byref(p), util.ptr(util.f32(st["eq"])), util.ptr(util.f32(st["pres"])), util.ptr(util.f32(st["sres"])),
                     util.ptr(tabs[0]), util.ptr(tabs[1]), 1, len(st["z"]), util.ptr(util.f32(st["z"])), util.ptr(util.f32(st["vp"])),
                     util.ptr(util.f32(st["vpvs"])), util.ptr(mf), util.ptr(origin), util.ptr(resid), util.ptr(tpred))
    assert rc == 0
    return mf, origin, tpred, tabs


def test_fine_grid_forward_matches_the_oracle():
    import mcmc_eq_b200 as mq
    from mcmc_eq_b200 import synth
    # a small array on the fine grid: 3 events x 6 stations keep the oracle's share of the test to a few dozen solves
    cfg, pk, truth = synth.workload(3, 6, 7, 0, fine=True)
    assert cfg.grid.nz == 2001 and int(np.sqrt(cfg.grid.nx ** 2 + cfg.grid.ny ** 2)) == 565
    n = 2
    smp = mq.Sampler(cfg, pk, n, 0, 3)
    rng = np.random.default_rng(4)
    st = fh.random_states(rng, cfg, pk, n, "posterior", 14)
    st[1]["z"], st[1]["vp"], st[1]["vpvs"] = util.voronoi_model(rng, 7, 0.0, 200.0, "contrast")      # head waves, reverse propagation
    for s in st:
        s["eq"] = truth["eq"] + rng.normal(0, 0.3, truth["eq"].shape).astype(np.float32)
        s["eq"][:, 2] = np.clip(s["eq"][:, 2], 0.5, 199.0)
    smp.profile(True)
    mf, origin = smp.forward_host(fh.fill_models(smp.new_models(), st), 3)
    smp.profile(False)
    assert list(smp.profile_kernels()) == ["eik_fine_kernel"]
    h, z0 = cfg.grid.h, cfg.grid.z0
    for c in range(n):
        iz1 = [int(np.float32(z - z0) / np.float32(h)) for z in st[c]["eq"][:, 2]]
        depths = sorted(set(iz1) | set(i + 1 for i in iz1) | ({0, 1, 1000, 2000} if c == 0 else {13, 1999}))
        rmf, rorg, rpred, tabs = _oracle_misfit_sparse(cfg, pk, st[c], depths)
        _r, tpred = smp.predictions(c)
        assert np.abs(tpred - rpred).max() <= 1e-4, float(np.abs(tpred - rpred).max())
        assert np.allclose(mf[c], rmf, rtol=2e-5, atol=1e-7), (mf[c], rmf)
        assert np.abs(origin[c] - rorg).max() <= 1e-4
        for ph in (1, 2):
            rows, idx = smp.rows(c, ph)
            assert idx.max() < 4
            for iz in depths:
                ref = tabs[ph - 1][idx, iz, :]
                err = np.abs(rows[:, iz, :] - ref)
                assert (err <= np.maximum(1e-4, 5e-6 * np.abs(ref))).all(), (c, ph, iz, float(err.max()))
    smp.close()


def test_fine_kernel_full_fields_match_the_oracle():
    """Whole fields through mq_eikonal_batch on a tall narrow plane (the box reaches the right edge long before the top and
    the bottom: every later row sweep runs along the masked dummy column), with a high-contrast model that raises head waves."""
    import mcmc_eq_b200 as mq
    rng = np.random.default_rng(6)
    nx, nz, h = 90, 700, 0.25          # 700 depth nodes: 2109 floats per lane, no shared-memory slice
    for kind in ("posterior", "contrast", "lvz"):
        z, vp, vpvs = util.voronoi_model(rng, 9, 0.0, (nz - 1) * h, kind)
        s = util.rasterise_np(z, vp, vpvs, h, 0.0, nz, 1)
        izs = np.array([0, 5, 11, 350, 688, 699], np.int32)
        t = mq.eikonal_batch(np.tile(s, (len(izs), 1)), izs, nx)
        for k, iz in enumerate(izs):
            tref, rc = util.oracle_time_2d(s, nx, int(iz))
            assert rc == 0
            err = np.abs(t[k] - tref)
            assert (err <= util.eikonal_tol(tref)).all(), (kind, int(iz), float(err.max()))
