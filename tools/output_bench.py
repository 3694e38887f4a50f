"""The asynchronous output path under load: many chains, a record per chain every `deci` accepted
models, drained by mq_drain_begin / a writer thread while the sampler keeps stepping, against the same run without output.

    python tools/output_bench.py [chains] [rounds]

Each round steps `chunk` iterations of the reference's mixed proposal string, then starts a drain; a writer thread waits for
the batch (device-to-host copy of exactly the packed records into pinned memory) and takes the records apart (the text
formatting of print_model_raw, src/mcmc_eq.c:234-248, is the C front end's job: host/mcmc_eq_main.c, tests/test_cli_gpu.py).  Prints one JSON line: device time per round with and without
output, records written, records lost."""
import io
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import mcmc_eq_b200 as mq  # noqa: E402
from mcmc_eq_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 6
chunk, deci = 100, 25          # >= 1 record per chain per 100 iterations at ~50 % acceptance


def fmt(r, out):
    """print_model_raw: one `mod` line, one EQ line per event, one RES line per station"""
    tri = " ".join(f"{a:f} {b:f} {c:f}" for a, b, c in zip(r["z"], r["vp"], r["vpvs"]))
    out.write(f"mod {r['code']}. {r['number']:8d} {r['dim']:3d} {r['rms']:f} " + " ".join(f"{x:f}" for x in r["noise"]) + " " + tri + "\n")
    for e, (q, o) in enumerate(zip(r["eq"], r["origin"])):
        out.write(f"EQ  {r['code']}. {r['number']:8d} {e} {r['rms']:f} {q[0]:f} {q[1]:f} {q[2]:f} {0.0:f} {o:f}\n")
    for k, (a, b) in enumerate(zip(r["pres"], r["sres"])):
        out.write(f"RES {r['code']}. {r['number']:8d} {k} {r['rms']:f} {a:f} {b:f}\n")


def run(with_output):
    cfg, pk, _ = synth.workload(200, 50, 33, 0, j_max_start=0, j_max_main=2**30, deci=deci)
    smp = mq.Sampler(cfg, pk, n, 0, 1000)
    smp.set_ring(8)
    smp.init_chains()
    smp.step(chunk, None)
    smp.drain()
    stats = dict(records=0, lost=0, bytes=0)
    threads = []

    def finish(batch):
        recs, lost = smp.drain_finish(batch)
        if stats["records"] == 0 and recs:        # the text of one record, for the byte count (the C front end formats all of them)
            buf = io.StringIO()
            fmt(recs[0], buf)
            stats["bytes"] = buf.tell()
        stats["records"] += len(recs); stats["lost"] += lost

    smp.sync()
    smp.timer_start(5)
    for _ in range(rounds):
        smp.step(chunk, None)
        if with_output:
            if len(threads) >= 2:
                threads.pop(0).join()
            t = threading.Thread(target=finish, args=(smp.drain_begin(),))
            t.start()
            threads.append(t)
    ms = smp.timer_stop(5)
    t0 = time.perf_counter()
    for t in threads:
        t.join()
    tail = time.perf_counter() - t0
    counts, _ll, _rms = smp.stats()
    smp.close()
    return ms / rounds, stats, tail, int(counts[:, 17].sum())


base_ms, _s, _t, _a = run(False)
out_ms, st, tail, accepted = run(True)
print(json.dumps({"chains": n, "iterations_per_round": chunk, "deci": deci, "rounds": rounds, "ms_per_round_no_output": base_ms,
                  "ms_per_round_with_output": out_ms, "overhead": out_ms / base_ms - 1.0, "records": st["records"], "lost": st["lost"],
                  "text_bytes_per_record": st["bytes"], "writer_tail_s": tail, "records_per_chain_per_100_iterations": st["records"] / n / (rounds * chunk / 100.0)}))
