"""The decimated-record path (print_model_raw, src/mcmc_eq.c:234-248, written every deci-th accepted model at :1163):
device-side ring, lost-record accounting, the asynchronous drain on the copy stream next to running steps, and the
bulk snapshot calls."""
import tempfile
import threading

import numpy as np
import pytest

from tests import inputs

pytestmark = pytest.mark.gpu


def _sampler(n, seed=3, ring=None, **over):
    import mcmc_eq_b200 as mq
    d = tempfile.mkdtemp(prefix="mqr_")
    kw = dict(j_max_start=0, j_max_main=10**6, deci=5)
    kw.update(over)
    cfgp, pkp = inputs.materialise("example2", d, **kw)
    cfg, pk = mq.read_config(cfgp), mq.Picks.read(pkp)
    smp = mq.Sampler(cfg, pk, n, 0, seed)
    if ring:
        smp.set_ring(ring)
    return cfg, pk, smp


def _key(r):
    return (r["chain"], r["number"])


def _same(a, b):
    for k in ("chain", "kind", "code", "number", "dim"):
        assert a[k] == b[k], k
    assert a["rms"] == b["rms"]
    for k in ("noise", "z", "vp", "vpvs", "eq", "origin", "pres", "sres"):
        assert np.array_equal(a[k], b[k]), k


def test_full_ring_drops_records_and_counts_them():
    """Three records' worth of accepted models without a drain on a one-slot ring: the first record of every chain is
    kept, the later ones are dropped and reported -- never silently (lost was always 0)."""
    cfg, pk, smp = _sampler(16, ring=1)
    smp.init_chains()
    smp.step(60, "QN")
    counts, _, _ = smp.stats()
    due = counts[:, 17] // 5                     # records each chain has produced
    assert (due >= 3).sum() > 0
    recs, lost = smp.drain()
    assert lost == int(np.maximum(due - 1, 0).sum()) and lost > 0
    assert sorted(r["chain"] for r in recs) == [c for c in range(16) if due[c] > 0]
    assert all(r["number"] == 4 for r in recs)   # the first decimated model (acce == deci), not a later one
    # after the drain the ring takes records again and nothing is reported twice
    smp.step(10, "QN")
    recs2, lost2 = smp.drain()
    c2, _, _ = smp.stats()
    assert all(r["number"] > 4 for r in recs2)
    assert len(recs2) + lost2 == int((c2[:, 17] // 5 - due).sum())
    smp.close()


def test_ring_keeps_every_record_in_order():
    cfg, pk, smp = _sampler(24, ring=8)
    smp.init_chains()
    got = []
    for _ in range(4):
        smp.step(35, "QRN")                      # at most 7 records per chain between drains
        recs, lost = smp.drain()
        assert lost == 0
        got += recs
    counts, _, _ = smp.stats()
    for c in range(24):
        nums = [r["number"] for r in got if r["chain"] == c]
        assert nums == [5 * (k + 1) - 1 for k in range(int(counts[c, 17]) // 5)]
    smp.close()


def test_asynchronous_drain_equals_synchronous_drain():
    """Same seed, same steps: one sampler is drained synchronously after every chunk, the other starts a drain and
    keeps stepping while a second thread waits for the batch and collects it."""
    _, _, a = _sampler(32, seed=21, ring=4)
    _, _, b = _sampler(32, seed=21, ring=4)
    a.init_chains(); b.init_chains()
    ra, rb, lost_b = [], [], [0]

    def finish(batch):
        recs, lost = b.drain_finish(batch)
        rb.extend(recs)
        lost_b[0] += lost

    threads = []
    for _ in range(6):
        a.step(10, "QRPN")
        recs, lost = a.drain()
        assert lost == 0
        ra += recs
        b.step(10, "QRPN")
        if len(threads) >= 2:                    # the handle owns two batches
            threads.pop(0).join()
        t = threading.Thread(target=finish, args=(b.drain_begin(),))
        t.start()
        threads.append(t)
    for t in threads:
        t.join()
    assert lost_b[0] == 0 and len(ra) == len(rb) > 0
    for x, y in zip(sorted(ra, key=_key), sorted(rb, key=_key)):
        _same(x, y)
    sa, sb = a.stats(), b.stats()
    assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1])
    a.close(); b.close()


def test_third_drain_in_flight_is_refused():
    import mcmc_eq_b200 as mq
    _, _, smp = _sampler(4)
    smp.init_chains()
    b1, b2 = smp.drain_begin(), smp.drain_begin()
    with pytest.raises(mq.MqError) as e:
        smp.drain_begin()
    assert e.value.code == -9
    smp.drain_finish(b1); smp.drain_finish(b2)
    smp.drain_finish(smp.drain_begin())
    smp.close()


def test_bulk_snapshots_equal_single_snapshots():
    _, _, smp = _sampler(12, seed=8)
    smp.init_chains()
    smp.step(30, "QVRPBDMN")
    for which in (0, 1):
        every = smp.snapshot_all(which)
        assert [r["chain"] for r in every] == list(range(12))
        for c in (0, 5, 11):
            _same(every[c], smp.snapshot(c, which))
    m = smp.get_models()
    cur = smp.snapshot_all(0)
    for c in range(12):
        d = m.dim[c]
        assert cur[c]["dim"] == d and np.array_equal(cur[c]["z"], m.z[c, :d]) and np.array_equal(cur[c]["eq"], m.eq[c])
        assert np.array_equal(cur[c]["origin"], m.origin[c]) and cur[c]["code"] == "S"
    smp.close()


def test_asynchronous_output_keeps_up_with_hundreds_of_chains():
    """512 chains, a record every 10 accepted models, drained asynchronously every 40 iterations by a second thread: nothing
    is lost, every due record arrives exactly once, and the sampler's device time stays within 10 % of the same run without
    output (tools/output_bench.py measures 0.3 % at 8192 chains)."""
    import mcmc_eq_b200 as mq
    from mcmc_eq_b200 import synth

    def run(with_output):
        cfg, pk, _ = synth.workload(40, 20, 33, 0, j_max_start=0, j_max_main=2**30, deci=10)
        smp = mq.Sampler(cfg, pk, 512, 0, 77)
        smp.set_ring(8)
        smp.init_chains()
        smp.step(40, None)
        first, lost0 = smp.drain()
        got, lost, threads = [], [0], []

        def finish(batch):
            recs, l = smp.drain_finish(batch)
            got.extend((r["chain"], r["number"]) for r in recs)
            lost[0] += l

        smp.sync()
        smp.timer_start(6)
        for _ in range(5):
            smp.step(40, None)
            if with_output:
                if len(threads) >= 2:
                    threads.pop(0).join()
                t = threading.Thread(target=finish, args=(smp.drain_begin(),))
                t.start()
                threads.append(t)
        ms = smp.timer_stop(6)
        for t in threads:
            t.join()
        counts, _, _ = smp.stats()
        smp.close()
        return ms, got, lost[0] + lost0, counts[:, 17], len(first)

    ms0, _g, _l, acc0, _f = run(False)
    ms1, got, lost, acc1, n_first = run(True)
    assert np.array_equal(acc0, acc1)                         # output does not touch the trajectories
    assert lost == 0
    assert len(got) == len(set(got)) == int((acc1 // 10).sum()) - n_first
    assert ms1 <= 1.10 * ms0, (ms0, ms1)
