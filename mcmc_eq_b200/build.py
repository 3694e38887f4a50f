"""Build the CUDA library (sm_100a) and the CPU checkers in-tree.

`python -m mcmc_eq_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a
GPU; the resulting .so files are git-ignored but travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libmcmceq_b200.so")
HOST_DIR = os.path.join(PKG, "host")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # keep FP32 products un-fused except where the source says fmaf(): the host build of
    # the solver core (tests/emu) is then bit-identical to the device code
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def cuda_sources() -> list[str]:
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    # host-side readers/writers of the reference's file formats are part of the same library
    srcs.append(os.path.join(HOST_DIR, "mq_io.c"))
    return srcs


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = cuda_sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "mcmceq_b200.h"))
    deps.append(os.path.join(HOST_DIR, "mq_io.h"))
    if not force and _newer(LIB, deps):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + srcs + ["-o", LIB, "-lcudart", "-ldl"]
    subprocess.run(cmd, check=True, cwd=ROOT)
    return LIB


def build_oracle(force: bool = False) -> None:
    """Compile oracle/ (our CPU restatement) and, when /root/reference exists, oracle/_ref."""
    args = ["make", "-C", os.path.join(ROOT, "oracle")]
    if force:
        args.append("-B")
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)


def build_emu(force: bool = False) -> str:
    """Host build of the CUDA solver core, used by the CPU tests only."""
    src = os.path.join(ROOT, "tests", "emu", "eik_emu.cpp")
    out = os.path.join(ROOT, "tests", "emu", "libeik_emu.so")
    deps = [src, os.path.join(CSRC, "eik_core.cuh"), os.path.join(CSRC, "eik_fast.cuh")]
    if force or not _newer(out, deps):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-but-set-variable",
                        src, "-o", out], check=True)
    return out


def build_emu_mt(force: bool = False) -> str:
    """Host build of the warp-synchronous solver with a warp of host threads (tests/emu/host_warp.h), CPU tests only."""
    src = os.path.join(ROOT, "tests", "emu", "eik_emu_mt.cpp")
    out = os.path.join(ROOT, "tests", "emu", "libeik_emu_mt.so")
    deps = [src, os.path.join(ROOT, "tests", "emu", "host_warp.h"), os.path.join(CSRC, "eik_core.cuh"), os.path.join(CSRC, "eik_fast.cuh")]
    if force or not _newer(out, deps):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-Wall", "-Wno-unknown-pragmas",
                        "-Wno-unused-but-set-variable", src, "-o", out], check=True)
    return out


def build_host(force: bool = False) -> None:
    mk = os.path.join(HOST_DIR, "Makefile")
    if os.path.exists(mk):
        subprocess.run(["make", "-C", HOST_DIR] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL)


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_host(force)          # before the oracle: oracle/_ref links the reference's drivers against the function-seam shim
    build_oracle(force)
    build_emu(force)
    build_emu_mt(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB)
