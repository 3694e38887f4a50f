"""The process seam: mcmc_eq_b200/host/mcmc_eq (C host code over the C ABI) run the way the reference executable is run
(src/mcmc_eq.c:332-338), its output files checked against the structure of a file written by the reference itself
(tests/golden/chain_ref_example2.out, same config: 60+140 accepted models, every 20th written) and fed to the
reference's own consumer, analyse_eq, when oracle/_ref holds it."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from tests import inputs, util

pytestmark = pytest.mark.gpu

EXE = os.path.join(util.ROOT, "mcmc_eq_b200", "host", "mcmc_eq")


def _shape(text):
    """Line structure of a chain file: (tag, number of tokens) per line, and the model numbers of the mod records."""
    lines = text.strip().split("\n")
    return [(ln.split()[0], len(ln.split())) for ln in lines], [int(ln.split()[2]) for ln in lines if ln.startswith("mod")]


def _run(d, n_chains, out, **over):
    kw = dict(j_max_start=60, j_max_main=140, deci=20, true_random=77)
    kw.update(over)
    cfgp, pkp = inputs.materialise("example2", d, **kw)
    cmd = [EXE, cfgp, os.path.join(d, out), pkp, "-q"] + (["-n", str(n_chains)] if n_chains > 1 else [])
    r = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return cfgp


def test_single_chain_file_has_the_reference_structure():
    assert os.path.exists(EXE), "host/mcmc_eq not built (python -m mcmc_eq_b200.build)"
    ref = open(os.path.join(util.GOLDEN, "chain_ref_example2.out")).read()
    with tempfile.TemporaryDirectory() as d:
        _run(d, 1, "rjx-000.out")
        got = open(os.path.join(d, "rjx-000.out")).read()
    s_ref, n_ref = _shape(ref)
    s_got, n_got = _shape(got)
    # same records in the same order; the mod/sta/bat lines differ in length only through the model dimension
    assert [t for t, _ in s_got] == [t for t, _ in s_ref]
    assert n_got == n_ref == list(range(19, 200, 20))
    for (t, k), ln in zip(s_got, got.strip().split("\n")):
        tok = ln.split()
        if t in ("sta", "mod", "bat"):
            assert k == 13 + 3 * int(tok[3]) and re.fullmatch(r"ST|BF|[QRPVMBDN]\.", tok[1])
        elif t == "EQ":
            assert k == 10
        elif t == "RES":
            assert k == 7
    # cnt lines: accepted sums to the chain length, tested counts every evaluated proposal
    cnt = {ln[4:12].strip(): [int(x) for x in ln.split()[-2:]] for ln in got.split("\n") if ln.startswith("cnt") and "a/r" in ln}
    assert sum(a for a, _ in cnt.values()) == 200
    tested = int(re.search(r"cnt RMS tested\s+(\d+)", got).group(1))
    assert 200 <= tested <= sum(a + r for a, r in cnt.values())
    # floats are printed with %f like the reference
    assert re.match(r"sta ST        0 +\d+ \d+\.\d{6} ", got)


def test_many_chains_write_rjx_files_that_analyse_eq_reads():
    with tempfile.TemporaryDirectory() as d:
        cfgp = _run(d, 5, "rjx-%03d.out")
        files = sorted(f for f in os.listdir(d) if f.startswith("rjx-"))
        assert files == [f"rjx-{k:03d}.out" for k in range(1, 6)]
        texts = [open(os.path.join(d, f)).read() for f in files]
        rms_end = []
        for t in texts:
            shape, nums = _shape(t)
            assert nums == list(range(19, 200, 20)) and shape[0][0] == "sta" and shape[-1][0] == "cnt"
            rms_end.append(float([ln for ln in t.split("\n") if ln.startswith("bat")][0].split()[4]))
        assert len(set(texts)) == 5                      # chains are independent
        start = [float(t.split("\n")[0].split()[4]) for t in texts]
        assert all(e <= s for e, s in zip(rms_end, start))   # best model is never worse than the start model
        exe = os.path.join(util.REF_DIR, "analyse_eq")
        if os.path.exists(exe):
            # the filter of scriptsV2/disp_m_average_sl.sh:86-91: burn-in by model number, no "BF", no "cnt"
            allp = os.path.join(d, "all.out")
            with open(allp, "w") as f:
                for t in texts:
                    for ln in t.split("\n"):
                        tok = ln.split()
                        if tok and tok[0] != "cnt" and tok[1] != "BF" and int(tok[2]) > 0:
                            f.write(ln + "\n")
            r = subprocess.run(f"ulimit -s unlimited; {exe} {cfgp} {allp} 0.1 0.02", shell=True, cwd=d, capture_output=True,
                               text=True, timeout=120)
            assert r.returncode == 0, r.stderr[-300:]
            tags = [ln.split()[0] for ln in r.stdout.split("\n") if ln.strip()]
            assert tags.count("EZ") == 225 and tags.count("STAN") == 61 and tags.count("RES") == 8 and tags.count("NOISE") == 1
            stan = np.array([[float(x) for x in ln.split()[1:]] for ln in r.stdout.split("\n") if ln.startswith("STAN")])
            assert (stan[:, 1] > 1.0).all() and (stan[:, 1] < 10.0).all()   # mean Vp per depth node inside the prior
            assert np.allclose(stan[:, 0], -2.0 + 0.5 * np.arange(61))


def test_out_name_without_pattern_inserts_chain_number():
    with tempfile.TemporaryDirectory() as d:
        _run(d, 2, "run.out", j_max_main=20)
        assert sorted(f for f in os.listdir(d) if f.startswith("run")) == ["run-001.out", "run-002.out"]
