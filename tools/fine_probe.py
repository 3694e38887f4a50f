"""One full forward of a few chains on the fine-grid plane (565 x 2001): the launch ncu captures for eik_fine_kernel.
    python tools/fine_probe.py [chains]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcmc_eq_b200 as mq
from mcmc_eq_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg, pk, _ = synth.workload(20, 10, 33, 0, fine=True)
smp = mq.Sampler(cfg, pk, n, 0, 2)
smp.init_chains()
smp.profile(True)
t0 = time.perf_counter()
smp.forward(3, want_origin=False)
smp.profile(False)
ms, k, per = smp.profile()
print(f"{n} chains: {ms / max(k, 1):.1f} ms per launch of {per} solves = {1000 * ms / max(k, 1) / per:.1f} us per solve, kernels {smp.profile_kernels()}")
