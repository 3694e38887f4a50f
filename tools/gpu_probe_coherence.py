"""How much of the eikonal kernel's time is lost to lanes of a warp doing different things?  Times the table rebuild
(1024 chains x 2 phases x 62 depths) for: the bench's random start models; 32 distinct models replicated so that every
warp holds 32 DIFFERENT models (same total work as ...); the same 32 models arranged so that every warp holds 32 copies of
ONE model (perfect coherence: the upper bound of any regrouping of solves)."""
import sys, numpy as np
sys.path.insert(0, '.')
import mcmc_eq_b200 as mq
from mcmc_eq_b200 import synth

cfg, pk, truth = synth.workload(200, 50, 33, 0, j_max_start=0, j_max_main=2**30, deci=2**30)
n = 1024
smp = mq.Sampler(cfg, pk, n, 0, 1000)
smp.init_chains()
m0 = smp.get_models()

def timed(m, label):
    smp.set_models(m)
    smp.forward(3)
    smp.profile(True)
    for _ in range(5):
        smp.forward(3)
    ms, k, per = smp.profile(False)
    print(f"{label:55s} eikonal {ms / k:7.3f} ms per launch of {per} solves", flush=True)

def arranged(idx):
    m = smp.new_models()
    for name in ("dim", "z", "vp", "vpvs", "eq", "pres", "sres", "noise"):
        getattr(m, name)[:] = getattr(m0, name)[idx]
    return m

timed(m0, "random start models (bench)")
# items are ordered chain-major within a source depth: item = chain*2 + phase, 32 consecutive items = 16 chains x 2 phases
base = np.arange(n)
timed(arranged(base % 32), "32 distinct models, every warp mixes 16 of them")
timed(arranged((base // 16) % 32), "32 distinct models, every warp holds ONE model (P and S)")
timed(arranged(np.zeros(n, int)), "one model everywhere")
smp.close()
