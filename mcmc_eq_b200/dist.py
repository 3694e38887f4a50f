"""Host-side plumbing for more than one GPU: one process per GPU (torchrun), chains sharded contiguously.

torch.distributed is used for the rendez-vous only (hand the NCCL unique id of the library's own communicator to every
rank, barriers, max-over-ranks of timings).  The two collectives of the path itself -- the posterior reduction and the
tempering all-gather -- run inside libmcmceq_b200.so on its own NCCL communicator (csrc/comm.cu).

`swap_plan` is the numpy statement of the library's swap rule (csrc/comm.cu: temper_swap_kernel); the CPU tests run it
on two gloo ranks to check that every rank reaches the same decisions from the same gathered numbers, and the GPU tests
compare the device against it.
"""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous chain range [first, first + count) of `rank`; counts differ by at most one."""
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def out_name(pattern: str, chain1: int) -> str:
    """File of global chain number `chain1` (1-based, like the SLURM array index of run/srun_mcmc_eq.sh:13):
    a printf pattern with one integer conversion, or the number inserted before the extension."""
    if "%" in pattern:
        return pattern % chain1
    stem, dot, ext = pattern.rpartition(".")
    return f"{stem}-{chain1:03d}.{ext}" if dot and "/" not in ext else f"{pattern}-{chain1:03d}"


def exchange_unique_id(dist, make_id, device=None) -> bytes:
    """Rank 0 creates the communicator id (`make_id()` -> 128 bytes), everybody receives it."""
    import torch
    rank = dist.get_rank()
    t = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        t = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


# ---- swap rule of the tempering round, on the host ------------------------------------------------------------------
def _philox_word(seed: int, a: int, b: int) -> int:
    """First output word of Philox4x32-10 keyed by `seed` with counter (b_lo, b_hi, a, 'swap')."""
    m0, m1, w0, w1, mask = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c0, c1, c2, c3 = b & mask, (b >> 32) & mask, a & mask, 0x73776170
    k0, k1 = seed & mask, (seed >> 32) & mask
    for _ in range(10):
        p0, p1 = m0 * c0, m1 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & mask, p1 >> 32, p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + w0) & mask, (k1 + w1) & mask
    return c0


def swap_uniform(seed: int, lo_chain: int, round_: int) -> np.float32:
    """Uniform deviate in (0, 1] of the pair whose lower chain is `lo_chain` in round `round_`."""
    return np.float32((np.float32(_philox_word(seed, lo_chain, round_) >> 1) + np.float32(1.0)) / np.float32(2147483648.0))


def swap_plan(full_ll, beta, round_: int, seed: int):
    """New temperatures after one round: pairs (2k+p, 2k+1+p), p = round & 1, swap iff
    log u < (beta_a - beta_b) (L_b - L_a).  full_ll / beta are the gathered arrays of ALL chains of the job."""
    L = np.asarray(full_ll, np.float64)
    b = np.asarray(beta, np.float64)
    out = np.asarray(beta, np.float32).copy()
    n, par, swapped = len(L), round_ & 1, 0
    for lo in range(par, n - 1, 2):
        hi = lo + 1
        u = swap_uniform(seed, lo, round_)
        if b[lo] != b[hi] and np.log(np.float64(u)) < (b[lo] - b[hi]) * (L[hi] - L[lo]):
            out[lo], out[hi] = np.float32(b[hi]), np.float32(b[lo])
            swapped += 1
    return out, swapped


def full_loglik(ll, noise, n_class):
    """ll - sum_c n_c ln sigma_c per chain (csrc/comm.cu: temper_pack_kernel)."""
    ll = np.asarray(ll, np.float64)
    out = ll.copy()
    for k in range(8):
        out -= float(n_class[k]) * np.log(np.asarray(noise, np.float32)[:, k].astype(np.float64))
    return out
