"""Worker of tests/test_multigpu_gpu.py: one process per GPU under torch.distributed.run.  Checks, on real NCCL:
the library's own communicator (id handed over by torch.distributed), the posterior all-reduce, a tempering round across
ranks, and that a chain's trajectory does not depend on which GPU runs it (global chain numbering of the random streams)."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import mcmc_eq_b200 as mq
    from mcmc_eq_b200 import dist as mqd
    from mcmc_eq_b200._lib import comm_unique_id
    from tests import inputs
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = tempfile.mkdtemp(prefix=f"mqmp{rank}_")
    cfgp, pkp = inputs.materialise("example2", d, j_max_start=30, j_max_main=100000, deci=5, true_random=3)
    cfg, pk = mq.read_config(cfgp), mq.Picks.read(pkp)
    n, seed = 8, 31
    smp = mq.Sampler(cfg, pk, n, local, seed)
    ident = mqd.exchange_unique_id(dist, comm_unique_id, device=f"cuda:{local}")
    smp.comm_init(ident, rank, world)            # also sets the chain offset to rank * n
    smp.posterior_begin(0.1, 0.02, 10)
    smp.init_chains()
    for _ in range(12):
        smp.step(5)
        smp.drain()
    local_post = smp.posterior_get()
    counts, ll, rms = smp.stats()
    smp.posterior_allreduce()
    glob_post = smp.posterior_get()
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(post=local_post, counts=counts, ll=ll, rms=rms))
    ok = {}
    for key in ("hist_vp", "hist_vpvs", "boundary", "vsum", "eqsum", "ressum", "noisesum"):
        want = sum(g["post"][key] for g in gathered)
        ok["allreduce_" + key] = bool(np.allclose(glob_post[key], want, rtol=1e-12, atol=1e-9))
    ok["allreduce_n"] = glob_post["n_models"] == sum(g["post"]["n_models"] for g in gathered) > 0
    # sharding invariance: the same 2n chains in ONE handle (no communicator) on rank 0's GPU
    if rank == 0:
        one = mq.Sampler(cfg, pk, n * world, local, seed)
        one.init_chains()
        for _ in range(12):
            one.step(5)
            one.drain()
        c1, l1, r1 = one.stats()
        one.close()
        ok["shard_invariant"] = bool(np.array_equal(c1, np.concatenate([g["counts"] for g in gathered])) and
                                     np.array_equal(l1, np.concatenate([g["ll"] for g in gathered])))
    # tempering across ranks
    ladder = np.float32([1.0, 0.6, 0.3, 0.15])
    smp.set_beta(np.tile(ladder, n // 4)[::-1].copy() if rank % 2 else np.tile(ladder, n // 4))
    swaps = 0
    for rnd in range(4):
        smp.step(6, "QN")
        _c, ll, _r = smp.stats()
        noise = smp.get_models().noise
        before = smp.get_beta()
        parts = [None] * world
        dist.all_gather_object(parts, dict(L=mqd.full_loglik(ll, noise, pk.n_class), b=before))
        want, k_all = mqd.swap_plan(np.concatenate([p["L"] for p in parts]), np.concatenate([p["b"] for p in parts]), rnd, seed)
        k = smp.temper_swap(rnd)
        after = smp.get_beta()
        ok[f"temper_round{rnd}"] = bool(np.array_equal(after, want[rank * n:(rank + 1) * n]))
        ks = [None] * world
        dist.all_gather_object(ks, k)
        ok[f"temper_count{rnd}"] = sum(ks) == k_all
        swaps += k_all
    ok["temper_some_swaps"] = swaps > 0
    smp.comm_destroy()
    smp.close()
    res = [None] * world
    dist.all_gather_object(res, ok)
    if rank == 0:
        print("MPRESULT " + json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
