"""Debug: one forward of the Example set (1230 posterior-like chains) so that a library built with -DEIKF_TRACE=<solve>
(tools/ab_build.py) prints the box phase of that solve: MCMCEQ_LIB=... MCMCEQ_EIKONAL_PIPE=0|1 python tools/trace_run.py"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcmc_eq_b200 as mq
from tests import inputs, fwd_helpers as fh
d = tempfile.mkdtemp(prefix="mqtr_")
name = os.environ.get("DIAG_SET", "example")
cfgp, pkp = inputs.materialise(name, d)
cfg, pk = mq.read_config(cfgp), mq.Picks.read(pkp)
n = 1230
smp = mq.Sampler(cfg, pk, n, 0, 1)
rng = np.random.default_rng(4)
st = fh.random_states(rng, cfg, pk, n, kind="lvz" if name == "example2" else "posterior")
smp.forward_host(fh.fill_models(smp.new_models(32), st), 3)
t, idx = smp.rows(0, 1)
print("rows", idx, "t[0, 19, :4]", t[0, 19, :4], "t[1, 19, :4]", t[1, 19, :4])
