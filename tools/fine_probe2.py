"""forward(3) against step(1, 'P') on the fine-grid plane: same number of solves per launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcmc_eq_b200 as mq
from mcmc_eq_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cfg, pk, _ = synth.workload(20, 10, 33, 0, fine=True, j_max_start=0, j_max_main=2**30, deci=2**30)
smp = mq.Sampler(cfg, pk, n, 0, 2)
smp.init_chains()
for what in ("forward", "step", "forward", "step"):
    smp.profile(True)
    if what == "forward":
        smp.forward(3, want_origin=False)
    else:
        smp.step(1, "P")
        smp.sync()
    smp.profile(False)
    ms, k, per = smp.profile()
    print(f"{what}: {ms / max(k, 1):.1f} ms per launch of {per} solves = {1000 * ms / max(k, 1) / per:.1f} us per solve, launches {k}, {smp.profile_kernels()}")
