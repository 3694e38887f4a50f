"""Raw receiver-row tables of a few chains, pipelined against fused eikonal kernel: where do they differ?"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = r"""
import sys, tempfile, numpy as np
sys.path.insert(0, %r)
import mcmc_eq_b200 as mq
from tests import inputs, fwd_helpers as fh
d = tempfile.mkdtemp(prefix="mqdr_")
import os
name = os.environ.get("DIAG_SET", "example2")
cfgp, pkp = inputs.materialise(name, d)
cfg, pk = mq.read_config(cfgp), mq.Picks.read(pkp)
n = 1230
smp = mq.Sampler(cfg, pk, n, 0, 1)
rng = np.random.default_rng(4)
st = fh.random_states(rng, cfg, pk, n, kind="lvz" if name == "example2" else "posterior")
smp.forward_host(fh.fill_models(smp.new_models(32), st), 3)
out = {}
for c in (0, 1, 4, 700):
    for ph in (1, 2):
        t, idx = smp.rows(c, ph)
        out[f"t{c}_{ph}"] = t; out["idx"] = idx
np.savez(sys.argv[1], **out)
"""
res = {}
for pipe in ("0", "1"):
    path = os.path.join(tempfile.mkdtemp(), "o.npz")
    r = subprocess.run([sys.executable, "-c", SCRIPT % ROOT, path], env=dict(os.environ, MCMCEQ_EIKONAL_PIPE=pipe), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    res[pipe] = dict(np.load(path))
a, b = res["0"], res["1"]
print("row index", a["idx"])
for k in sorted(a):
    if k == "idx":
        continue
    ta, tb = a[k], b[k]
    bad = ta != tb
    print(k, "differing entries", int(bad.sum()), "of", bad.size)
    if bad.any():
        r, iz, x = np.nonzero(bad)
        print("   rows:", np.bincount(r, minlength=ta.shape[0]), " source depths (count per iz):", dict(zip(*np.unique(iz, return_counts=True))))
        print("   x range of differences: min", x.min(), "max", x.max(), " histogram over x//10:", np.bincount(x // 10))
        for j in range(min(5, len(r))):
            print("   sample r,iz,x =", r[j], iz[j], x[j], " fused", ta[r[j], iz[j], x[j]], " pipe", tb[r[j], iz[j], x[j]])
        # for one bad source depth, the whole x profile of the first differing row
        i0 = iz[0]; r0 = r[0]
        xs = np.nonzero(bad[r0, i0])[0]
        print("   iz", i0, "row", r0, "bad x:", xs[:40], " pipe values there:", tb[r0, i0, xs[:8]])
