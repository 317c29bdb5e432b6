"""Synthetic stand-ins for the reference's fitted inputs.

est/netest.Rds, est/netstats.Rds and est/epistats.Rds are git-LFS pointers in /root/reference, so the
population, the three partnership networks, the mortality table and the act/condom regression coefficients are generated
here with ARTnet-scale placeholder values (SURVEY §8d "Config 1").  Everything is an *input* of the C ABI
(xgc_set_tables / xgc_upload_attr / xgc_set_edgelist), so real fitted values can be dropped in unchanged.
Host-side NumPy, not on the hot path: this plays the role of initialize_msm()/init_msm() (sim1_02_lhs.r:180-184, 201).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

# table offsets mirror include/xgc/abi.h (kept literal here so this module does not need libxgc at import time)
ASMR_AGES, GLM_NCOEF = 101, 26
T_ASMR = 0
T_GLM_AI = T_ASMR + 4 * ASMR_AGES
T_GLM_OI = T_GLM_AI + GLM_NCOEF
T_GLM_CONDMC = T_GLM_OI + GLM_NCOEF
T_GLM_CONDOO = T_GLM_CONDMC + GLM_NCOEF
T_RACE_PROP = T_GLM_CONDOO + GLM_NCOEF
T_ROLE_PROB = T_RACE_PROP + 4
T_NET_DEG = T_ROLE_PROB + 12
T_NET_DUR = T_NET_DEG + 3
T_NET_DEGW = T_NET_DUR + 3
T_NET_RISKW = T_NET_DEGW + 27
NTABLE = T_NET_RISKW + 5
G_B0, G_DUR, G_DUR2, G_PT2, G_DUR_PT2, G_CA, G_CA2, G_HCP, G_PREP, G_PAD, G_RC = range(11)

RACE_PROP = np.array([0.13, 0.19, 0.08, 0.60])                  # B, H, O, W
ROLE_PROB = np.tile(np.array([0.24, 0.27, 0.49]), (4, 1))       # I, R, V
NET_DEG = np.array([0.40, 0.55, 0.07])                          # main, casual, one-off (per week)
NET_DUR = np.array([250.0, 90.0, 1.0])                          # weeks
HIV_PREV0 = np.array([0.22, 0.13, 0.10, 0.08])


# Degree- and risk-conditioned formation of the surrogate network (abi.h XGC_T_NET_DEGW / XGC_T_NET_RISKW): acceptance weight of
# an agent by its current (main, casual) degree class 0 / 1 / 2+ for each layer, and by risk group for one-off contacts.
# Placeholders with the signs of the ARTnet formation models the reference fits (main partnerships close to monogamous and
# rarer among men with casual partners; casual concurrency damped; one-off contacts concentrated in the upper risk groups).
NET_DEGW = np.array([
    [[1.00, 0.60, 0.35], [0.04, 0.03, 0.02], [0.00, 0.00, 0.00]],      # main:   rows = own main degree, columns = casual degree
    [[1.00, 0.55, 0.30], [0.70, 0.40, 0.20], [0.50, 0.30, 0.15]],      # casual
    [[1.00, 1.00, 1.00], [0.60, 0.75, 0.90], [0.60, 0.75, 0.90]],      # one-off
])
NET_RISKW = np.array([0.2, 0.4, 0.6, 0.8, 1.0])         # (a 5 : 1 spread; a steeper one mostly costs formation tries: acceptance ~ E[w]^2)
# XGC_T_NET_DEG is the PROPOSAL intensity (ties proposed per week = deg * n / 2 / duration); the weights above thin it.  These
# factors bring the realised mean degrees of the bench workload back to NET_DEG (oracle, 500 weeks: 0.40 / 0.555 / 0.070).
NET_DEG_PROPOSAL = np.array([1.5, 1.12, 1.15])


def make_tables() -> np.ndarray:
    t = np.zeros(NTABLE)
    ages = np.arange(ASMR_AGES)
    annual = 0.0008 * np.exp(0.055 * np.clip(ages - 18, 0, None))        # ~0.0008/yr at 18 rising to ~0.01/yr at 64
    for r, mult in enumerate([1.5, 0.9, 0.9, 1.0]):
        wk = 1 - (1 - np.clip(annual * mult, 0, 1)) ** (1 / 52)
        wk[ages >= 65] = 1.0                                             # leave the sexually active population at 65
        t[T_ASMR + r * ASMR_AGES: T_ASMR + (r + 1) * ASMR_AGES] = wk
    rc = np.array([[0.00, -0.05, -0.02, -0.08], [-0.05, 0.03, -0.01, -0.04], [-0.02, -0.01, 0.02, -0.03], [-0.08, -0.04, -0.03, 0.05]])

    def glm(off, b0, dur, dur2, pt2, dur_pt2, ca, ca2, hcp, prep):
        t[off + G_B0], t[off + G_DUR], t[off + G_DUR2], t[off + G_PT2], t[off + G_DUR_PT2] = b0, dur, dur2, pt2, dur_pt2
        t[off + G_CA], t[off + G_CA2], t[off + G_HCP], t[off + G_PREP] = ca, ca2, hcp, prep
        t[off + G_RC: off + G_RC + 16] = rc.ravel()

    glm(T_GLM_AI, np.log(95.0), -0.0016, 1.2e-6, -0.45, 0.0004, -0.006, 1.0e-5, 0.10, 0.0)      # anal acts / year
    glm(T_GLM_OI, np.log(75.0), -0.0012, 1.0e-6, -0.30, 0.0003, -0.005, 1.0e-5, 0.05, 0.0)      # oral acts / year
    glm(T_GLM_CONDMC, -0.9, -0.002, 1.5e-6, 0.65, 0.0005, 0.004, -1.0e-5, -0.6, -0.35)          # logit P(condom), main/casual
    glm(T_GLM_CONDOO, 0.1, 0.0, 0.0, 0.0, 0.0, 0.003, -1.0e-5, -0.6, -0.40)                     # logit P(condom), one-off
    t[T_RACE_PROP:T_RACE_PROP + 4] = RACE_PROP
    t[T_ROLE_PROB:T_ROLE_PROB + 12] = ROLE_PROB.ravel()
    t[T_NET_DEG:T_NET_DEG + 3] = NET_DEG * NET_DEG_PROPOSAL
    t[T_NET_DUR:T_NET_DUR + 3] = NET_DUR
    t[T_NET_DEGW:T_NET_DEGW + 27] = NET_DEGW.ravel()
    t[T_NET_RISKW:T_NET_RISKW + 5] = NET_RISKW
    return t


def _cat(u, prob):
    return np.minimum((u[:, None] >= np.cumsum(prob)[None, :]).sum(1), len(prob) - 1)


def init_population(n: int, seed: int, init: Dict[str, float], params: np.ndarray | None = None, hiv: bool = True,
                    at: int = 1) -> Dict[str, np.ndarray]:
    """Attribute vectors of the starting population (the job of initialize_msm + init_status_msm + init_sti_msm)."""
    from . import params as P
    p = P.default_vector() if params is None else params
    rng = np.random.default_rng(seed)
    a: Dict[str, np.ndarray] = {}
    a["age"] = 18.0 + rng.random(n) * 47.0
    race = _cat(rng.random(n), RACE_PROP)
    a["race"] = race + 1.0
    role = _cat(rng.random(n), ROLE_PROB[0])
    a["role.class"] = role.astype(float)
    a["ins.quot"] = np.where(role == 0, 1.0, np.where(role == 1, 0.0, rng.random(n))).astype(np.float32).astype(float)
    a["risk.grp"] = 1.0 + rng.integers(0, 5, n)
    a["circ"] = (rng.random(n) < P.get(p, "circ_prob")[race]).astype(float)
    a["late.tester"] = (rng.random(n) < P.get(p, "hiv_test_late_prob")[race]).astype(float)
    a["tt.traj"] = np.full(n, 2.0)
    a["act.stopper"] = (rng.random(n) < P.get(p, "act_stopper_prob")[0]).astype(float)
    a["arrival.time"] = np.full(n, float(at))
    # gonorrhoea: prev.x * N infected among agents whose role makes the site available (init_sti_msm [BK])
    avail = {"rGC": role != 0, "uGC": role != 1, "pGC": np.ones(n, bool)}
    sp = P.get(p, "gc_sympt_prob")
    for s, (site, key) in enumerate([("rGC", "prev.rgc"), ("uGC", "prev.ugc"), ("pGC", "prev.pgc")]):
        ids = np.flatnonzero(avail[site])
        k = int(round(init.get(key, 0.05) * n))
        inf = rng.choice(ids, size=min(k, ids.size), replace=False)
        st = np.zeros(n); st[inf] = 1
        a[site] = st
        a[site + ".infTime"] = np.where(st == 1, float(at), -1.0)
        a[site + ".sympt"] = ((rng.random(n) < sp[s]) & (st == 1)).astype(float)
        a[site + ".tx"] = np.zeros(n)
        a[site + ".txTime"] = np.full(n, -1.0)
        a[site + ".dur"] = np.where(st == 1, np.maximum(1, rng.poisson(P.get(p, "gc_ntx_int")[s], n)), -1).astype(float)
    if hiv:
        pos = rng.random(n) < HIV_PREV0[race]
        tsi = rng.integers(20, 500, n)
        a["status"] = pos.astype(float)
        a["inf.time"] = np.where(pos, at - tsi - 1.0, -1.0)              # strictly < 0, never the NA code -1
        a["inf.time"] = np.where(pos & (a["inf.time"] == -1), -2.0, a["inf.time"])
        a["stage"] = np.where(pos, 3.0, 0.0)
        a["stage.time"] = np.where(pos, np.minimum(a["inf.time"] + 13, -2.0), -1.0)
        diag = pos & (rng.random(n) < 0.8)
        a["diag.status"] = diag.astype(float)
        a["diag.time"] = np.where(diag, -2.0, -1.0)
        tx = diag & (rng.random(n) < 0.65)
        a["tx.status"] = tx.astype(float)
        ever = tx | (diag & (rng.random(n) < 0.1))
        a["cuml.time.on.tx"] = np.where(ever, rng.integers(1, 200, n), 0).astype(float)
        a["cuml.time.off.tx"] = np.where(pos, rng.integers(0, 100, n), 0).astype(float)
        a["tx.init.time"] = np.where(ever, -2.0, -1.0)
        a["vl"] = np.where(tx, 1.5, np.where(pos, 4.5, 0.0))
    return a


def init_network(attrs: Dict[str, np.ndarray], seed: int, at: int = 1, deg=NET_DEG, dur=NET_DUR) -> Tuple[list, list]:
    """Three cross-sectional edge lists (1-based, tail < head) with role-compatible random pairs, and the partnership
    start weeks of the main/casual ties (dat$temp$plist)."""
    rng = np.random.default_rng(seed + 7919)
    n = attrs["race"].size
    role = attrs["role.class"].astype(int)
    els, starts = [], []
    for l in range(3):
        m = int(round(deg[l] * n / 2))
        pairs = np.empty((0, 2), dtype=np.int64)
        while pairs.shape[0] < m:
            c = rng.integers(0, n, size=(2 * (m - pairs.shape[0]) + 8, 2))
            ok = (c[:, 0] != c[:, 1]) & ~((role[c[:, 0]] == 0) & (role[c[:, 1]] == 0)) & ~((role[c[:, 0]] == 1) & (role[c[:, 1]] == 1))
            c = np.sort(c[ok], axis=1)
            pairs = np.unique(np.vstack([pairs, c]), axis=0) if pairs.size else np.unique(c, axis=0)
            if pairs.shape[0] > m:
                pairs = pairs[rng.permutation(pairs.shape[0])[:m]]
        els.append((pairs + 1).astype(np.int32))
        age_wk = rng.geometric(1.0 / dur[l], size=m) if l < 2 else np.zeros(m, dtype=np.int64)
        starts.append((at - age_wk).astype(np.int32))
    return els, starts


def capacity(n: int, weeks_between_compactions: int = 64, a_rate: float = 0.0006) -> int:
    """Slot capacity: initial N + growth head-room (population drift + tombstones until the next compaction)."""
    # the synthetic demography settles ~16 % above the initial size over a 60-year burn-in (profiles/burnin_3120_r01.json)
    grow = int(n * 0.25) + int(a_rate * n * weeks_between_compactions * 1.5) + 256
    return (n + grow + 127) // 128 * 128


def edge_capacity(n: int, deg=NET_DEG) -> list:
    return [int(deg[l] * n / 2 * 1.35) + 256 for l in range(3)]
