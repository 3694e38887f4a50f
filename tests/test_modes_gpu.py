"""Secondary modes and edge cases of the same path:
prior sampling (aflag == 1: the likelihood is never evaluated and every proposal is accepted, src/misfit.c:61,
src/mcmc_eq.c:1135) and the Voronoi tie rule (a depth node equidistant from two nuclei belongs to the one with the
HIGHER index, find_in_cell's `<=`, src/mod_grd.c:102) through the device rasteriser."""
import tempfile

import numpy as np
import pytest

from tests import inputs, util
from tests import fwd_helpers as fh

pytestmark = pytest.mark.gpu


def _sampler(n, seed=3, **over):
    import mcmc_eq_b200 as mq
    d = tempfile.mkdtemp(prefix="mqm_")
    cfgp, pkp = inputs.materialise("example2", d, **over)
    cfg, pk = mq.read_config(cfgp), mq.Picks.read(pkp)
    return mq, cfg, pk, mq.Sampler(cfg, pk, n, 0, seed)


def test_prior_sampling_accepts_everything_and_never_evaluates(oracle):
    n, iters = 16, 40
    mq, cfg, pk, smp = _sampler(n, aflag=1, j_max_start=0, j_max_main=10**6, deci=10)
    assert cfg.aflag == 1
    smp.profile(True)
    smp.init_chains()
    counts, ll, rms = smp.stats()
    assert (ll == 0).all() and (rms == 0).all()                 # cal_fit_newx returns zero sums (src/misfit.c:61)
    smp.step(iters, "QVRPBDMN")
    _ms, eik_launches, _ = smp.profile(False)
    assert eik_launches == 0                                     # no table was built, no pick looked up
    counts, ll, rms = smp.stats()
    assert (counts[:, 17] == iters).all() and (counts[:, 18] == 0).all()   # alpha12 = 1 for every arm, eligible or not (:1135)
    assert (counts[:, 1:17].reshape(n, 8, 2).sum((0, 2)) > 0).all()         # all eight arms drawn
    assert (ll == 0).all() and (rms == 0).all()
    # the chain samples the prior: states stay inside the prior box and valid under model_valid (src/mcmc_eq.c:180-229)
    m = smp.get_models()
    g = cfg.grid
    zmin, zmax = g.z0, g.z0 + (g.nz - 1) * g.h
    inv = -abs(cfg.inv_control) if cfg.inv_control > 0 else cfg.inv_control
    assert len(set(int(d) for d in m.dim)) > 1                   # births and deaths were accepted
    for c in range(n):
        d = int(m.dim[c])
        assert (m.vp[c, :d] > cfg.vpmin).all() and (m.vp[c, :d] < cfg.vpmax).all()
        assert (m.z[c, :d] >= zmin).all() and (m.z[c, :d] <= zmax).all()
        assert oracle.ch_model_valid(d, util.ptr(m.z[c]), util.ptr(m.vp[c]), util.ptr(m.vpvs[c]), g.h, zmin, zmax, inv) == 0
    assert (m.noise > cfg.noise_min).all() and (m.noise < cfg.noise_max).all()
    recs, lost = smp.drain()
    assert lost == 0 and len(recs) == n * min(iters // 10, 4) and all(r["rms"] == 0 for r in recs)
    smp.close()


def test_depth_node_equidistant_from_two_nuclei_takes_the_higher_index():
    """Example2 grid: nodes at -2 + 0.5 k km.  Nuclei at 1.0 and 2.0 km leave the node at 1.5 km exactly equidistant
    (0.25 km^2 both ways in float).  Listed as (1.0, 2.0) the node takes the velocity of the 2.0 km nucleus, listed as
    (2.0, 1.0) that of the 1.0 km nucleus: two different slowness columns, each checked against the oracle."""
    mq, cfg, pk, smp = _sampler(2)
    g = cfg.grid
    rng = np.random.default_rng(2)
    st = fh.random_states(rng, cfg, pk, 2)
    vp = {-1.0: 3.0, 1.0: 4.0, 2.0: 6.5, 9.0: 7.5}
    for c, order in enumerate(([-1.0, 1.0, 2.0, 9.0], [-1.0, 2.0, 1.0, 9.0])):
        st[c]["z"] = np.float32(order)
        st[c]["vp"] = np.float32([vp[z] for z in order])
        st[c]["vpvs"] = np.float32([1.7, 1.75, 1.8, 1.85])
        st[c]["eq"], st[c]["pres"], st[c]["sres"] = st[0]["eq"], st[0]["pres"], st[0]["sres"]
    k = int(round((1.5 - g.z0) / g.h))
    assert g.z0 + k * g.h == 1.5
    s0 = util.rasterise_np(st[0]["z"], st[0]["vp"], st[0]["vpvs"], g.h, g.z0, g.nz, 1)
    s1 = util.rasterise_np(st[1]["z"], st[1]["vp"], st[1]["vpvs"], g.h, g.z0, g.nz, 1)
    assert s0[k] == np.float32(g.h / 6.5) and s1[k] == np.float32(g.h / 4.0) and (np.delete(s0, k) == np.delete(s1, k)).all()
    mf, origin = smp.forward_host(fh.fill_models(smp.new_models(32), st), 3)
    tabs = []
    for c in range(2):
        rmf, rorg, _r, _t, rtab = fh.oracle_forward(cfg, pk, st[c]["z"], st[c]["vp"], st[c]["vpvs"], st[c]["eq"], st[c]["pres"],
                                                     st[c]["sres"], want_tables=True)
        assert np.allclose(mf[c], rmf, rtol=2e-5, atol=1e-6)
        for ph in (1, 2):
            t = smp.table(c, ph)
            assert (np.abs(t - rtab[ph - 1]) <= util.eikonal_tol(rtab[ph - 1])).all()
            tabs.append(t)
    assert np.abs(tabs[0] - tabs[2]).max() > 1e-2       # the tie decides a whole cell's velocity: the two tables differ
    smp.close()
