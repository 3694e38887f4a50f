"""Pipelined against fused eikonal kernel on the same inputs (the script of tests/test_pipe_gpu.py): prints what differs."""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_pipe_gpu import SCRIPT  # noqa: E402
res = {}
for pipe in ("0", "1"):
    path = os.path.join(tempfile.mkdtemp(), "o.npz")
    r = subprocess.run([sys.executable, "-c", SCRIPT % ROOT, path, "1230"], env=dict(os.environ, MCMCEQ_EIKONAL_PIPE=pipe), capture_output=True, text=True)
    if r.returncode:
        print("pipe", pipe, "FAILED", r.stderr[-600:]); sys.exit(1)
    res[pipe] = dict(np.load(path))
bad = [k for k in res["0"] if not np.array_equal(res["0"][k], res["1"][k])]
print("pipelined kernel differs from the fused kernel in:", bad if bad else "nothing")
for k in bad:
    a, b = res["0"][k], res["1"][k]
    idx = np.argwhere(a != b)
    print(k, len(idx), "of", a.size, "max |diff|", float(np.nanmax(np.abs(a.astype(float) - b.astype(float)))), "first", idx[:5].tolist())
