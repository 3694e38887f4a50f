"""Replay of a recorded reference chain (tests/golden/replay_example2.npz, made by tools/make_golden.py from the
unmodified reference under oracle/replay_log.c): reconstructs, for every evaluated proposal, which arm of the proposal
switch produced it (src/mcmc_eq.c:866-1130) and the proposal-ratio term log_fac of the birth / death / noise arms
(:1038-1039, :1070-1071, :1114-1117) from the recorded models, so that a sampler can be driven with the reference's
own proposal stream and its accept/reject decisions compared one by one."""
from __future__ import annotations

import os

import numpy as np

from tests import util
from tests.util import ptr, f32


def load(name="example2"):
    return dict(np.load(os.path.join(util.GOLDEN, f"replay_{name}.npz")))


def state_of(log, i):
    d = int(log["dim"][i])
    return dict(dim=d, z=log["z"][i, :d].copy(), vp=log["vp"][i, :d].copy(), vpvs=log["vpvs"][i, :d].copy(), eq=log["eq"][i].copy(),
                pres=log["pres"][i].copy(), sres=log["sres"][i].copy(), noise=log["noise"][i].copy(), origin=log["origin"][i].copy())


def classify(cur, prop, calct):
    """-> (kind, q_idx).  `cur` the chain state before the proposal, `prop` the proposed state."""
    if prop["dim"] == cur["dim"] + 1:
        return "B", -1
    if prop["dim"] == cur["dim"] - 1:
        return "D", -1
    if not np.array_equal(prop["noise"], cur["noise"]):
        return "N", -1
    if not np.array_equal(prop["z"], cur["z"]):
        return "M", -1
    if not np.array_equal(prop["vpvs"], cur["vpvs"]):
        return "V", -1
    if not np.array_equal(prop["vp"], cur["vp"]):
        return "P", -1
    ch = np.nonzero((prop["eq"] != cur["eq"]).any(axis=1))[0]
    if len(ch):
        assert len(ch) == 1
        return "Q", int(ch[0])
    if not (np.array_equal(prop["pres"], cur["pres"]) and np.array_equal(prop["sres"], cur["sres"])):
        return "R", -1
    # nothing changed (a perturbation that rounded away): any arm with the recorded calct does
    return {0: "Q", 2: "V", 3: "P"}[int(calct)], 0 if calct == 0 else -1


def log_fac(kind, cur, prop, cfg, n_class):
    """Proposal-ratio term of the arm, through the oracle's restatement (oracle/chain.c)."""
    L = util.oracle()
    if kind == "B":
        d = cur["dim"]
        parent = L.fm_find_in_cell(ptr(f32(cur["z"])), d, float(prop["z"][d]))      # src/mcmc_eq.c:1028
        return L.ch_logfac_birth(cfg.sdevvp, cfg.vpmin, cfg.vpmax, float(prop["vp"][d]), float(cur["vp"][parent]), cfg.sdevvpvs,
                                 cfg.vpvsmin, cfg.vpvsmax, float(prop["vpvs"][d]), float(cur["vpvs"][parent]))
    if kind == "D":
        d = cur["dim"]
        dead = next((i for i in range(d - 1) if cur["z"][i] != prop["z"][i] or cur["vp"][i] != prop["vp"][i]), d - 1)
        nb = L.fm_find_neighbor_cell(ptr(f32(cur["z"])), d, dead)                   # src/mcmc_eq.c:1066
        return L.ch_logfac_death(cfg.sdevvp, cfg.vpmin, cfg.vpmax, float(cur["vp"][dead]), float(cur["vp"][nb]), cfg.sdevvpvs,
                                 cfg.vpvsmin, cfg.vpvsmax, float(cur["vpvs"][dead]), float(cur["vpvs"][nb]))
    if kind == "N":
        nc = np.ascontiguousarray(n_class, np.int32)
        return L.ch_logfac_noise(ptr(nc, util.ip), ptr(f32(cur["noise"])), ptr(f32(prop["noise"])))
    return 0.0


def loglik(mf, noise):
    L = util.oracle()
    return -L.ch_misfit(ptr(f32(mf)), ptr(f32(noise))) / 2.0


def proposals(log, cfg, n_class):
    """Yields (i, kind, q_idx, log_fac, u, accepted, cur_state, proposed_state) for records 1.., tracking the chain state."""
    cur = state_of(log, 0)
    for i in range(1, len(log["u"])):
        prop = state_of(log, i)
        kind, q = classify(cur, prop, log["calct"][i])
        lf = log_fac(kind, cur, prop, cfg, n_class)
        acc = bool(log["accepted"][i])
        yield i, kind, q, lf, float(log["u"][i]), acc, cur, prop
        if acc:
            cur = prop
