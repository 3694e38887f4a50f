"""Time of the misfit kernel alone (CUDA events round its launches, mq_profile_misfit) at a given shape:
    python tools/misfit_probe.py [chains events stations]      (default: the config-4 shard, 1024 x 2000 x 100)
MCMCEQ_MISFIT_SCRATCH=1 selects the round-1 data flow (residuals through a global scratch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcmc_eq_b200 as mq
from mcmc_eq_b200 import synth

n, ne, ns = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1024, 2000, 100)
cfg, pk, _ = synth.workload(ne, ns, 33, 0, j_max_start=0, j_max_main=2**30, deci=2**30)
smp = mq.Sampler(cfg, pk, n, 0, 1)
smp.init_chains()
for _ in range(3):
    smp.forward(0, want_origin=False)
smp.profile(True)
for _ in range(10):
    smp.forward(0, want_origin=False)       # calct = 0: lookup + residuals on the existing tables
smp.profile(False)
k, ms = smp.profile_misfit()
print(f"misfit_kernel {n} chains x {ne} events x {ns} stations ({n * pk.n_picks / 1e6:.0f} M pick look-ups per launch): "
      f"{ms / k:.3f} ms per launch ({k} launches), scratch={os.environ.get('MCMCEQ_MISFIT_SCRATCH', '0')}")
smp.close()
