"""On-device posterior accumulation (mq_posterior_*): pass 1 of the reference's analyse_eq (src/analyse_eq.c:496-643)
applied on the GPU to every decimated model.  Checked (1) exactly against a numpy statement of the same binning on the
drained records, (2) against the reference's own analyse_eq run on the text of those records (oracle/_ref, when present)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from tests import inputs, util

pytestmark = pytest.mark.gpu


def _nearest(z, zq):
    d2 = ((z.astype(np.float32) - np.float32(zq)) ** 2).astype(np.float32)
    return len(z) - 1 - int(np.argmin(d2[::-1]))          # ties -> highest index (src/mod_grd.c:93-110)


def _tria_profile(z, v, nz, h, z0):
    """Linear interpolation between the depth-sorted nuclei in the reference's float arithmetic; a depth node that no
    segment holds keeps the segment of the node above (src/analyse_eq.c:573-580)."""
    order = np.argsort(z, kind="stable")
    zs, vs = z[order].astype(np.float32), v[order].astype(np.float32)
    out, k = np.zeros(nz, np.float32), 0
    for i in range(nz):
        zz = np.float32(np.float32(i) * np.float32(h) + np.float32(z0))
        for si in range(len(zs) - 1):
            if zs[si] <= zz < zs[si + 1]:
                k = si
        a = np.float32(np.float32(vs[k + 1] - vs[k]) / np.float32(zs[k + 1] - zs[k]))
        b = np.float32(vs[k] - np.float32(a * zs[k]))
        out[i] = np.float32(np.float32(a * zz) + b)
    return out


@pytest.mark.parametrize("tria", [0, 1])
def test_posterior_matches_numpy_and_analyse_eq(tria):
    import mcmc_eq_b200 as mq
    from mcmc_eq_b200.io import format_record
    d = tempfile.mkdtemp(prefix="mqp_")
    cfgp, pkp = inputs.materialise("example2", d, j_max_start=40, j_max_main=160, deci=10, true_random=5, tria=tria)
    cfg, pk = mq.read_config(cfgp), mq.Picks.read(pkp)
    n, burn, dv, dvs = 6, 50, np.float32(0.1), np.float32(0.02)
    smp = mq.Sampler(cfg, pk, n, 0, 5)
    dims = smp.posterior_begin(float(dv), float(dvs), burn)
    smp.init_chains()
    recs = []
    for _ in range(60):
        smp.step(10)
        r, lost = smp.drain()
        assert lost == 0
        recs += r
    post = smp.posterior_get()
    smp.close()
    used = [r for r in recs if r["number"] > burn]
    assert post["n_models"] == len(used) > 20
    g = cfg.grid
    assert (dims.ndv, dims.ndvpvs, dims.nz) == (int((cfg.vpmax - cfg.vpmin) / dv) + 1, int((cfg.vpvsmax - cfg.vpvsmin) / dvs) + 1, g.nz)
    # ---- (1) numpy statement of the binning
    hp, hs, bd = np.zeros((dims.ndv, g.nz), np.int32), np.zeros((dims.ndvpvs, g.nz), np.int32), np.zeros(g.nz, np.int32)
    vsum = np.zeros((g.nz, 4))
    for r in used:
        if tria:
            assert r["z"][0] == np.float32(g.z0) and r["z"][1] == np.float32(g.z0 + (g.nz - 1) * g.h)
            pv, pr = _tria_profile(r["z"], r["vp"], g.nz, g.h, g.z0), _tria_profile(r["z"], r["vpvs"], g.nz, g.h, g.z0)
        for i in range(g.nz):
            zz = np.float32(np.float32(i) * np.float32(g.h) + np.float32(g.z0))
            if tria:
                vv, rr0 = pv[i], pr[i]
            else:
                k = _nearest(r["z"], zz)
                vv, rr0 = np.float32(r["vp"][k]), np.float32(r["vpvs"][k])
                bd[i] += vv != np.float32(r["vp"][_nearest(r["z"], np.float32(zz - np.float32(g.h)))])
            vv = min(max(vv, np.float32(cfg.vpmin)), np.float32(cfg.vpmax))
            hp[int(np.float32(vv - np.float32(cfg.vpmin)) / dv), i] += 1
            rr = min(max(rr0, np.float32(cfg.vpvsmin)), np.float32(cfg.vpvsmax))
            hs[int(np.float32(rr - np.float32(cfg.vpvsmin)) / dvs), i] += 1
            vsum[i] += [float(vv), float(vv) ** 2, float(rr), float(rr) ** 2]
    assert np.array_equal(post["hist_vp"], hp) and np.array_equal(post["hist_vpvs"], hs) and np.array_equal(post["boundary"], bd)
    assert np.allclose(post["vsum"], vsum, rtol=1e-12)
    eq = np.stack([np.concatenate([r["eq"], r["origin"][:, None]], 1) for r in used]).astype(np.float64)
    assert np.allclose(post["eqsum"][:, :4], eq.sum(0), rtol=1e-10, atol=1e-9) and np.allclose(post["eqsum"][:, 4:], (eq ** 2).sum(0), rtol=1e-10)
    res = np.stack([np.stack([r["pres"], r["sres"]], 1) for r in used]).astype(np.float64)
    assert np.allclose(post["ressum"][:, :2], res.sum(0), atol=1e-9) and np.allclose(post["ressum"][:, 2:], (res ** 2).sum(0), atol=1e-9)
    noise = np.stack([r["noise"] for r in used]).astype(np.float64)
    assert np.allclose(post["noisesum"][:8], noise.sum(0)) and np.allclose(post["noisesum"][8:], (noise ** 2).sum(0))
    # ---- (2) the reference's analyse_eq on the text of the same records
    exe = os.path.join(util.REF_DIR, "analyse_eq")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/analyse_eq not built")
    allp = os.path.join(d, "all.out")
    with open(allp, "w") as f:
        for r in used:
            f.write(format_record(r, pk.reftime))
    out = subprocess.run(f"ulimit -s unlimited; {exe} {cfgp} {allp} {float(dv)} {float(dvs)}", shell=True, cwd=d, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-300:]
    lines = [ln.split() for ln in out.stdout.split("\n") if ln.strip()]
    binp = np.array([int(t[3]) for t in lines if t[0] == "BINP"]).reshape(dims.ndv, g.nz)
    binv = np.array([int(t[3]) for t in lines if t[0] == "BINV"]).reshape(dims.ndvpvs, g.nz)
    # the text carries 6 decimals, so a value within 5e-7 of a bin edge may land in the neighbouring bin
    assert np.abs(binp - post["hist_vp"]).sum() <= 4 and np.abs(binv - post["hist_vpvs"]).sum() <= 4
    assert binp.sum() == post["hist_vp"].sum() == len(used) * g.nz
    m = len(used)
    eqr = np.array([[float(x) for x in t[2:5]] for t in lines if t[0] == "EQ"])
    assert np.allclose(eqr, post["eqsum"][:, :3] / m, atol=1.5e-3)              # printed with %9.3f
    resr = np.array([[float(x) for x in t[2:4]] for t in lines if t[0] == "RES"])
    assert np.allclose(resr, post["ressum"][:, :2] / m, atol=1.5e-3)
    noi = np.array([float(x) for x in [t for t in lines if t[0] == "NOISE"][0][1:9]])
    assert np.allclose(noi, (post["noisesum"][:8] / m)[[0, 2, 4, 6, 1, 3, 5, 7]], atol=1.5e-3)
    stan = np.array([[float(x) for x in t[1:]] for t in lines if t[0] == "STAN"])
    assert np.allclose(stan[:, 1], post["vsum"][:, 0] / m, atol=1.5e-3) and np.allclose(stan[:, 3], post["vsum"][:, 2] / m, atol=1.5e-3)
    assert np.allclose(stan[:, 11], post["boundary"] / m, atol=1e-5)


def test_tempering_single_gpu_matches_host_plan():
    import mcmc_eq_b200 as mq
    from mcmc_eq_b200 import dist as mqd
    d = tempfile.mkdtemp(prefix="mqt_")
    cfgp, pkp = inputs.materialise("example2", d, j_max_start=1000, j_max_main=1000, deci=1000)
    cfg, pk = mq.read_config(cfgp), mq.Picks.read(pkp)
    n = 32
    smp = mq.Sampler(cfg, pk, n, 0, 17)
    smp.init_chains()
    assert np.array_equal(smp.get_beta(), np.ones(n, np.float32))
    assert smp.temper_swap(0) == 0                                  # all temperatures equal: nothing to swap
    ladder = np.float32([1.0, 0.5, 0.25, 0.1])
    beta = np.tile(ladder, n // 4)
    smp.set_beta(beta)
    smp.step(30, "QN")
    total_swaps = 0
    for rnd in range(4):
        _c, ll, _r = smp.stats()
        noise = smp.get_models().noise
        before = smp.get_beta()
        want, k_want = mqd.swap_plan(mqd.full_loglik(ll, noise, pk.n_class), before, rnd, seed=17)
        k = smp.temper_swap(rnd)
        after = smp.get_beta()
        assert np.array_equal(after, want) and k == k_want, (rnd, after, want)
        assert sorted(after) == sorted(before)
        total_swaps += k
        smp.step(10, "QN")
    assert total_swaps > 0
    # hot chains accept more: the tempered acceptance uses beta * delta(ll)
    counts, _, _ = smp.stats()
    acc = counts[:, 17] / (counts[:, 17] + counts[:, 18])
    assert np.isfinite(acc).all()
    with pytest.raises(mq.MqError):
        smp.set_beta(np.zeros(n, np.float32))
    smp.close()
