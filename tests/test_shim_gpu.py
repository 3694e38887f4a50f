"""The FUNCTION seam: the reference's own chain driver, unmodified, linked against
mcmc_eq_b200/libmcmceq_shim.so (host/gpu_shim.c) instead of its CPU code, run with a fixed libc seed and compared with
the file the all-CPU reference wrote for the same configuration (tests/golden/chain_ref_example2.out: Example2, seed 77,
60 + 140 accepted models, every 20th written).  The driver consumes one libc rand() stream, so a single accept/reject
decision taken differently would shift every later proposal: equal model numbers, proposal letters, dimensions and
accept/reject counters at the end of the file mean that all ~300 decisions were the reference's.

  oracle/_ref/mcmc_eq_shim_t2d   only time_2d.o replaced (src/fdtimes.h:6-7): every eikonal solve runs on the GPU
  oracle/_ref/mcmc_eq_shim       misfit.c and interpol.c replaced as well (src/mcmc_eq.c:77-78 includes them by name; the
                                 driver is compiled through a link next to two empty files of those names): cal_fit_newx
                                 (src/misfit.c:45) is the batched GPU forward, the tables stay on the device

Both binaries are built by oracle/Makefile in the build container (they need the reference sources) and travel to the
GPU box; the test is skipped when they are absent."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from tests import inputs, util

pytestmark = pytest.mark.gpu


def _records(text):
    """[(tag, code, number, dim, rms, floats...)] of the sta / mod / bat lines and the cnt lines of a chain file."""
    recs, cnt = [], []
    for ln in text.strip().split("\n"):
        tok = ln.split()
        if tok[0] in ("sta", "mod", "bat"):
            recs.append((tok[0], tok[1], int(tok[2]), int(tok[3]), np.array([float(x) for x in tok[4:]])))
        elif tok[0] == "cnt":
            cnt.append(ln)
    return recs, cnt


def _run(exe, d):
    cfgp, pkp = inputs.materialise("example2", d, j_max_start=60, j_max_main=140, deci=20, true_random=77)
    out = os.path.join(d, "chain.out")
    r = subprocess.run([exe, cfgp, out, pkp], cwd=d, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-1500:]
    return open(out).read()


@pytest.mark.parametrize("exe_name", ["mcmc_eq_shim", "mcmc_eq_shim_t2d"])
def test_reference_driver_on_the_gpu_seam_takes_the_reference_decisions(exe_name):
    exe = os.path.join(util.REF_DIR, exe_name)
    if not os.path.exists(exe) or not os.path.exists(os.path.join(util.ROOT, "mcmc_eq_b200", "libmcmceq_shim.so")):
        pytest.skip(f"oracle/_ref/{exe_name} not built (needs the reference sources: make -C oracle ref)")
    ref = open(os.path.join(util.GOLDEN, "chain_ref_example2.out")).read()
    with tempfile.TemporaryDirectory() as d:
        got = _run(exe, d)
    r_ref, c_ref = _records(ref)
    r_got, c_got = _records(got)
    assert c_got == c_ref                                   # accepted / rejected per proposal kind, models evaluated
    assert [(t, c, n, k) for t, c, n, k, _ in r_got] == [(t, c, n, k) for t, c, n, k, _ in r_ref]
    for (_t, _c, _n, _k, a), (_t2, _c2, _n2, _k2, b) in zip(r_got, r_ref):
        assert a.shape == b.shape
        assert np.allclose(a, b, rtol=0, atol=2e-4), float(np.abs(a - b).max())     # RMS, sigmas, nuclei (printed with %f)
    # hypocentres, origin times and station corrections of every record
    eq_ref = np.array([[float(x) for x in ln.split()[3:]] for ln in ref.split("\n") if ln.startswith("EQ ")])
    eq_got = np.array([[float(x) for x in ln.split()[3:]] for ln in got.split("\n") if ln.startswith("EQ ")])
    assert eq_ref.shape == eq_got.shape and np.abs(eq_ref - eq_got).max() <= 2e-4
    res_ref = np.array([[float(x) for x in ln.split()[3:]] for ln in ref.split("\n") if ln.startswith("RES ")])
    res_got = np.array([[float(x) for x in ln.split()[3:]] for ln in got.split("\n") if ln.startswith("RES ")])
    assert res_ref.shape == res_got.shape and np.abs(res_ref - res_got).max() <= 2e-4


def test_traveltimet_twin_matches_the_oracle(oracle):
    """mq_traveltimet (reference src/interpol.c:43-83) on a host table against the oracle's lookup, inside and outside."""
    import ctypes as C
    import mcmc_eq_b200 as mq
    g = util.EXAMPLE2_GRID
    nx = util.nxmod_of(g)
    rng = np.random.default_rng(3)
    tab = np.cumsum(rng.uniform(0.01, 0.2, (g["nz"], nx)).astype(np.float32), axis=1)
    rows = (C.POINTER(C.c_float) * g["nz"])(*[tab[k].ctypes.data_as(C.POINTER(C.c_float)) for k in range(g["nz"])])
    L = mq.lib()
    L.mq_traveltimet.argtypes = [C.POINTER(C.POINTER(C.c_float)), C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.POINTER(C.c_float), C.c_int]
    fg = util.FmGrid(g["h"], g["nx"], g["ny"], g["nz"], g["x0"], g["y0"], g["z0"])
    full = np.zeros((g["nz"], g["nz"], nx), np.float32)
    full[5] = tab
    for dist, z in [(0.0, -2.0), (3.3, 1.7), (41.2, 20.3), (67.9, 27.9), (68.1, 3.0), (12.0, 28.1), (200.0, 5.0)]:
        out = C.c_float(0)
        assert L.mq_traveltimet(rows, g["nx"], g["ny"], g["nz"], g["h"], dist, z, g["z0"], C.byref(out), 0) == 0
        want = oracle.fm_traveltime(util.ptr(full[5]), C.byref(fg), dist, z)
        assert out.value == want or abs(out.value - want) <= 1e-6 * abs(want), (dist, z, out.value, want)
