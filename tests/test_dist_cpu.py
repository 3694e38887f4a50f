"""Host-side logic of the N > 1 path on two gloo ranks (CPU): sharding, the hand-over of the communicator id, the
reduction of per-rank posterior accumulators and the swap plan of a tempering round (every rank must reach the same
decisions from the same gathered numbers)."""
import os
import socket

import numpy as np
import pytest

from mcmc_eq_b200 import dist as mqd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_chains, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the communicator id made on rank 0 reaches every rank unchanged
        ident = bytes((7 * i + 3) % 256 for i in range(128))
        got = mqd.exchange_unique_id(dist, lambda: ident)
        # 2. shards tile the chain range
        first, count = mqd.shard_range(total_chains, rank, world)
        # 3. per-rank accumulators (what mq_posterior_get returns) sum to the accumulators of the whole job
        rng = np.random.default_rng(100 + rank)
        hist = rng.integers(0, 50, (11, 7)).astype(np.int32)
        t = torch.from_numpy(hist.copy())
        dist.all_reduce(t)
        # 4. tempering: gather (L, beta) of all chains, every rank computes the plan for the whole job
        rngc = np.random.default_rng(5)                       # the job's chains, same on every rank
        L_all = rngc.normal(-4000, 30, total_chains)
        b_all = rngc.choice([1.0, 0.7, 0.5, 0.35], total_chains).astype(np.float32)
        mine = torch.tensor(np.stack([L_all[first:first + count], b_all[first:first + count].astype(np.float64)], 1))
        sizes = [mqd.shard_range(total_chains, r, world)[1] for r in range(world)]
        parts = [torch.zeros((s, 2), dtype=torch.float64) for s in sizes]
        dist.all_gather(parts, mine) if len(set(sizes)) == 1 else dist.all_gather_object(parts, mine)
        allv = torch.cat([torch.as_tensor(p) for p in parts]).numpy()
        plans = [mqd.swap_plan(allv[:, 0], allv[:, 1], r, seed=99) for r in range(4)]
        q.put((rank, got == ident, first, count, hist, t.numpy().copy(), [p[0].tobytes() for p in plans], [p[1] for p in plans],
               np.array_equal(allv[:, 0], L_all)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 65])
def test_two_ranks_agree(total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ok0, f0, c0, h0, s0, plan0, sw0, g0), (r1, ok1, f1, c1, h1, s1, plan1, sw1, g1) = res
    assert ok0 and ok1                                       # id handed over
    assert f0 == 0 and f1 == c0 and c0 + c1 == total and abs(c0 - c1) <= 1
    assert np.array_equal(s0, h0 + h1) and np.array_equal(s1, s0)
    assert g0 and g1 and plan0 == plan1 and sw0 == sw1       # same gathered numbers, same swap decisions
    assert sum(sw0) > 0


def test_swap_plan_properties():
    rng = np.random.default_rng(1)
    n = 200
    L = rng.normal(-4000, 50, n)
    b = rng.choice([1.0, 0.8, 0.6, 0.4], n).astype(np.float32)
    for rnd in range(6):
        nb, k = mqd.swap_plan(L, b, rnd, seed=7)
        assert sorted(nb) == sorted(b)                                   # temperatures are permuted, never created
        par = rnd & 1
        changed = np.nonzero(nb != b)[0]
        for i in changed:                                                # only partners of this round's pairing exchange
            j = ((i - par) ^ 1) + par
            assert 0 <= j < n and nb[i] == b[j] and nb[j] == b[i]
        assert np.array_equal(nb, mqd.swap_plan(L, b, rnd, seed=7)[0])   # deterministic
        # a swap that moves the better-fitting state to the colder chain is always accepted
        for lo in range(par, n - 1, 2):
            if (b[lo] - b[lo + 1]) * (L[lo + 1] - L[lo]) >= 0 and b[lo] != b[lo + 1]:
                assert nb[lo] == b[lo + 1]
    # all temperatures equal: nothing can change
    assert np.array_equal(mqd.swap_plan(L, np.ones(n, np.float32), 0, 7)[0], np.ones(n, np.float32))
    assert 0.0 < mqd.swap_uniform(7, 0, 0) <= 1.0


def test_shard_and_names():
    for total, world in ((1024, 8), (10, 4), (3, 8)):
        spans = [mqd.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        assert all(spans[r][0] + spans[r][1] == spans[r + 1][0] for r in range(world - 1))
    assert mqd.out_name("rjx-%03d.out", 7) == "rjx-007.out"
    assert mqd.out_name("run/rjx.out", 12) == "run/rjx-012.out"
    assert mqd.out_name("rjx", 3) == "rjx-003"
