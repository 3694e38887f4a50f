"""Build a variant of the CUDA library next to the in-tree one for A/B runs on the GPU box:
    python tools/ab_build.py <name> [nvcc flags ...]      ->  mcmc_eq_b200/libmcmceq_b200_<name>.so
    MCMCEQ_LIB=mcmc_eq_b200/libmcmceq_b200_<name>.so python bench.py ...
The variant .so is git-ignored and travels with the gpurun snapshot like the main one."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcmc_eq_b200 import build  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(build.PKG, f"libmcmceq_b200_{name}.so")
cmd = ["nvcc"] + build.NVCC_FLAGS + flags + build.cuda_sources() + ["-o", out, "-lcudart", "-ldl"]
subprocess.run(cmd, check=True, cwd=ROOT)
print("built", out)
