import os, subprocess, sys, tempfile
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
from tests.test_pipe_gpu import SCRIPT
res = {}
for lc in ("0", "1"):
    path = os.path.join(tempfile.mkdtemp(), "o.npz")
    r = subprocess.run([sys.executable, "-c", SCRIPT % ROOT, path, "1230"], env=dict(os.environ, MCMCEQ_EIKONAL_PIPE="1", MCMCEQ_PIPE_LC=lc), capture_output=True, text=True)
    if r.returncode:
        print("FAILED", r.stderr[-600:]); sys.exit(1)
    res[lc] = dict(np.load(path))
for k in res["0"]:
    a, b = res["0"][k], res["1"][k]
    if not np.array_equal(a, b):
        d = np.abs(a.astype(float) - b.astype(float))
        idx = np.argwhere(a != b)
        print(k, "differs in", len(idx), "of", a.size, "max", np.nanmax(d), "first idx", idx[:6].tolist())
print("done")
