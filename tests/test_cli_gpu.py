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


def test_fw_mod_prints_what_the_reference_fw_mod_prints():
    """host/fw_mod (C over the C ABI) against the stdout of the reference's own fw_mod for the same chain record
    (tests/golden/fw_mod_example2.txt, written by tools/make_golden.py from oracle/_ref/fw_mod): same lines, residuals and
    predictions within 1e-4 s, distances / depths / observed times as printed."""
    sys_path = os.path.join(util.ROOT, "tools")
    import sys
    sys.path.insert(0, sys_path)
    from make_golden import last_block
    exe = os.path.join(util.ROOT, "mcmc_eq_b200", "host", "fw_mod")
    assert os.path.exists(exe), "host/fw_mod not built"
    ref = open(os.path.join(util.GOLDEN, "fw_mod_example2.txt")).read().strip().split("\n")
    with tempfile.TemporaryDirectory() as d:
        cfgp, pkp = inputs.materialise("example2", d, j_max_start=60, j_max_main=140, deci=20, true_random=77)
        blk = os.path.join(d, "block")
        open(blk, "w").write(last_block(open(os.path.join(util.GOLDEN, "chain_ref_example2.out")).read()))
        r = subprocess.run([exe, cfgp, blk, pkp], cwd=d, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = r.stdout.strip().split("\n")
    assert len(got) == len(ref) - 1 == 225 + 3600
    worst = 0.0
    for a, b in zip(got, ref[:-1]):
        ta, tb = a.split(), b.split()
        assert len(ta) == len(tb) and ta[0] == tb[0] if ta[0] == "EVENT" else ta[-1] == tb[-1]
        if ta[0] == "EVENT":
            assert ta[:6] == tb[:6] and abs(float(ta[6]) - float(tb[6])) < 1e-4        # origin time
        else:
            va, vb = [float(x) for x in ta[:6]], [float(x) for x in tb[:6]]
            assert va[1] == vb[1] and va[2] == vb[2] and va[4] == vb[4]                 # distance, depth, observed time
            assert abs(va[0] - vb[0]) < 1e-4 and abs(va[5] - vb[5]) < 1e-4 and abs(va[3] - vb[3]) < 1e-4
            worst = max(worst, abs(va[5] - vb[5]))
    # stderr: "Start model found with loglikelihood L RMS=R"
    want = ref[-1].split()
    have = r.stderr.strip().split("\n")[-1].split()
    assert abs(float(have[-1].split("=")[1]) - float(want[-1].split("=")[1])) < 2e-6
    assert float(have[-2]) == float(want[-2]) == -0.5      # the reference's constant (cal_fit_newx returns 1.0)


def test_start_from_model_dat():
    """aflag == 3 (config line 34): velocity model, hypocentres, station corrections and sigmas of the start model come from
    model.dat in the working directory, an analyse_eq result file (src/mcmc_eq.c:381,636-731)."""
    rng = np.random.default_rng(12)
    with tempfile.TemporaryDirectory() as d:
        zs = np.array([-1.0, 2.5, 7.0, 15.0])
        vps = np.array([3.9, 5.1, 6.0, 6.8])
        rs = np.array([1.9, 1.78, 1.73, 1.71])
        eq = np.round(np.stack([rng.uniform(-8, 8, 225), rng.uniform(-8, 8, 225), rng.uniform(1, 14, 225)], 1), 3)
        res = np.round(rng.normal(0, 0.1, (8, 2)), 3)
        noise = np.round(rng.uniform(0.05, 0.4, 8), 3)            # file order p0 p1 p2 p3 s0 s1 s2 s3
        with open(os.path.join(d, "model.dat"), "w") as f:
            for z, v, r in zip(zs, vps, rs):
                f.write("STAN %7.3f %7.3f %7.3f %7.3f %7.3f %7.3f %7.3f %7.3f %7.3f %7.3f %7.3f %7.5f\n" % (z, 9, 9, 9, 9, v, 9, r, 9, 9, 9, 0))
            for i, q in enumerate(eq):
                f.write("EQ %4d %9.3f %9.3f %9.3f %9.3f %9.3f %9.3f %14.3f %7.3f %7.3f %9.5f\n" % (i, q[0], q[1], q[2], 0, 0, 0, 0, 0, 0, 0))
            for i, r in enumerate(res):
                f.write("RES %4d %7.3f %7.3f %7.3f %7.3f\n" % (i, r[0], r[1], 0, 0))
            f.write("NOISE " + " ".join("%7.3f" % x for x in list(noise) + [0] * 8) + "\n")
        _run(d, 2, "rjx-%03d.out", aflag=3, inp_model_switch="VQRN", j_max_start=5, j_max_main=5, deci=5)
        for k in (1, 2):
            lines = open(os.path.join(d, f"rjx-{k:03d}.out")).read().split("\n")
            sta = lines[0].split()
            assert sta[:2] == ["sta", "ST"] and int(sta[3]) == 4
            got_noise = np.array([float(x) for x in sta[5:13]])
            assert np.allclose(got_noise, noise, atol=1e-6)
            tri = np.array([float(x) for x in sta[13:13 + 12]]).reshape(4, 3)
            assert np.allclose(tri[:, 0], zs, atol=1e-6) and np.allclose(tri[:, 1], vps, atol=1e-6) and np.allclose(tri[:, 2], rs, atol=1e-6)
            got_eq = np.array([[float(x) for x in ln.split()[5:8]] for ln in lines[1:226]])
            assert np.allclose(got_eq, eq, atol=1e-5)
            got_res = np.array([[float(x) for x in ln.split()[5:7]] for ln in lines[226:234]])
            assert np.allclose(got_res, res, atol=1e-6)
            assert float(sta[4]) > 0


def test_fw_prints_what_the_reference_fw_prints():
    """host/fw (gridded model from an analyse_eq-style result file, src/fw.c) against the reference fw's own output for the
    same res.dat (tests/golden/fw_example2.npz, from oracle/_ref/fw): what Example/make_synthetics turns into synthetic picks."""
    import sys
    sys.path.insert(0, os.path.join(util.ROOT, "tools"))
    from make_golden import write_res_dat, parse_forward_stdout
    exe = os.path.join(util.ROOT, "mcmc_eq_b200", "host", "fw")
    assert os.path.exists(exe), "host/fw not built"
    g = np.load(os.path.join(util.GOLDEN, "fw_example2.npz"))
    with tempfile.TemporaryDirectory() as d:
        cfgp, pkp = inputs.materialise("example2", d)
        res = os.path.join(d, "res.dat")
        write_res_dat(res, g["zn"], g["vpn"], g["rn"], g["eq"], g["pres"], g["sres"])
        r = subprocess.run([exe, cfgp, res, pkp], cwd=d, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    ev, pk = parse_forward_stdout(r.stdout)
    assert ev.shape == g["events"].shape and pk.shape == g["picks"].shape
    assert np.array_equal(ev[:, :4], g["events"][:, :4]) and np.abs(ev[:, 4] - g["events"][:, 4]).max() < 1e-4   # origin times
    assert np.array_equal(pk[:, [1, 2, 4, 6]], g["picks"][:, [1, 2, 4, 6]])      # distance, depth, observed time, phase
    assert np.abs(pk[:, [0, 3, 5]] - g["picks"][:, [0, 3, 5]]).max() < 1e-4      # residual, origin, predicted time
