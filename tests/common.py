"""Shared helpers for the parity tests: build the same synthetic case on the CUDA path (through the C ABI) and on the
CPU oracle, and compare every attribute vector, edge list, act count and epi tally bit for bit."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from xgc_msm_b200 import params as P  # noqa: E402
from xgc_msm_b200 import synth  # noqa: E402
from xgc_msm_b200.abi import K  # noqa: E402


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def random_midcourse_state(a: dict, seed: int, at: int) -> dict:
    """Perturb an initial population into a 'mid-epidemic' state so that every branch of every module is exercised:
    treatment clocks, screening clocks, exposure flags, PrEP users, all HIV stages.  All clocks are < at."""
    rng = np.random.default_rng(seed + 31337)
    n = a["race"].size
    a = {k: v.copy() for k, v in a.items()}

    def clock(valid, t):            # a week clock that is never the NA code -1
        t = np.asarray(t, dtype=float)
        return np.where(valid, np.where(t == -1, -2, t), -1.0)

    for site in ("rGC", "uGC", "pGC"):
        inf = a[site] == 1
        a[site + ".infTime"] = clock(inf, at - 1 - rng.integers(0, 12, n))
        tx = inf & (rng.random(n) < 0.35)
        a[site + ".tx"] = tx.astype(float)
        a[site + ".txTime"] = clock(tx, np.maximum(a[site + ".infTime"], at - 1 - rng.integers(0, 4, n)))
        a[site + ".dur"] = np.where(inf, rng.integers(1, 20, n), -1).astype(float)
    a["last.screen"] = clock(rng.random(n) < 0.5, at - 1 - rng.integers(0, 80, n))
    a["sti.expo"] = rng.integers(0, 32, n).astype(float)
    pos = a["status"] == 1
    stage = np.where(pos, rng.integers(1, 5, n), 0)
    tsi = np.where(stage == 1, rng.integers(0, 7, n), np.where(stage == 2, rng.integers(6, 13, n), rng.integers(13, 600, n)))
    a["stage"] = stage.astype(float)
    a["inf.time"] = np.where(pos, at - 1 - tsi, -1).astype(float)
    a["inf.time"] = np.where(pos & (a["inf.time"] == -1), -2, a["inf.time"])
    a["stage.time"] = np.where(pos, np.where(a["inf.time"] + 5 == -1, -2, a["inf.time"] + 5), -1).astype(float)
    diag = a["diag.status"] == 1
    a["last.neg.test"] = clock(~diag & (rng.random(n) < 0.6), at - 1 - rng.integers(0, 60, n))
    a["last.neg.test"] = np.where(rng.random(n) < 0.05, at, a["last.neg.test"]) * (~diag) + (-1.0) * diag
    a["vl"] = np.where(pos, rng.random(n) * 6.5, 0.0).astype(np.float32).astype(float)
    a["tt.traj"] = rng.integers(1, 4, n).astype(float)
    prep = ~diag & (rng.random(n) < 0.15)
    a["prepStat"] = prep.astype(float)
    a["prepClass"] = np.where(prep, rng.integers(1, 4, n), 0).astype(float)
    a["prepStartTime"] = clock(prep, at - 1 - rng.integers(0, 100, n))
    a["prepLastRisk"] = clock(prep, at - 1 - rng.integers(0, 70, n))
    a["prep.ind.last"] = clock(rng.random(n) < 0.4, at - 1 - rng.integers(0, 60, n))
    return a


def make_case(n: int, seed: int, midcourse: bool = False, at0: int = 1, **control):
    env = P.prior_midpoint_env()
    p = P.params_from_env(env, n_init=n)
    # make rare branches common enough to be hit at test sizes
    p = P.param_msm(p, gc_asympt_visit_prob=0.05, hiv_test_rate=0.05, prep_start_prob=0.5, prep_discont_rate=0.03,
                    tt_part_supp=0.3, tt_full_supp=0.4, tt_dur_supp=0.3, tx_halt_full_rr=0.8, tx_halt_dur_rr=0.5,
                    tx_reinit_full_rr=1.1, tx_reinit_dur_rr=1.2, cond_fail=(0.1, 0.05, 0.0, 0.02), sti_cond_fail=(0.05, 0.1, 0.0, 0.03),
                    trans_scale=(2.0, 1.5, 1.2, 1.0), a_rate=0.002, aids_mr=0.02)
    ctl = P.control_msm(**control).flags
    a = synth.init_population(n, seed, P.init_msm(0.1, 0.1, 0.1), p, at=at0)
    if midcourse:
        a = random_midcourse_state(a, seed, at0 + 1)
    els, starts = synth.init_network(a, seed, at=at0)
    cap = synth.capacity(n, a_rate=0.002)
    return dict(n=n, params=p, control=ctl, tables=synth.make_tables(), attrs=a, els=els, starts=starts, cap=cap,
                e_cap=synth.edge_capacity(n), at0=at0)


def load_oracle(case, seed: int, replicate_id: int, t_cap: int = 64):
    from oracle.oracle import Oracle
    o = Oracle(case["cap"], case["e_cap"], t_cap, seed=seed, replicate_id=replicate_id)
    o.set_params(case["params"]); o.set_control(case["control"]); o.set_tables(case["tables"])
    o.set_population(case["n"], case["at0"])
    for k, v in case["attrs"].items():
        o.set_attr(k, v)
    for l in range(3):
        o.set_edgelist(l, case["els"][l], case["starts"][l])
    return o


def load_handle(cases, seed: int, t_cap: int = 64, replicate_offset: int = 0, device: int = 0, compact_every: int = 0):
    """cases: list of cases with identical capacities -> one handle with len(cases) replicates."""
    from xgc_msm_b200.abi import Handle
    c0 = cases[0]
    h = Handle(len(cases), c0["cap"], c0["e_cap"], t_cap, seed=seed, device=device, replicate_offset=replicate_offset,
               compact_every=compact_every)
    h.set_tables(c0["tables"])
    for r, c in enumerate(cases):
        h.set_params(c["params"], r); h.set_control(c["control"], r)
        h.set_population(r, c["n"], c["at0"])
        for k, v in c["attrs"].items():
            h.upload_attr(r, k, v)
        for l in range(3):
            h.set_edgelist(r, l, c["els"][l], c["starts"][l])
    return h


ATTRS = None


def attr_names():
    global ATTRS
    if ATTRS is None:
        from xgc_msm_b200 import abi
        ATTRS = abi.names("attr")
    return ATTRS


def compare_state(h, r: int, o, at: int, what: str = "", acts: bool = True):
    """Assert bit-exact equality of every attribute, every edge list, the act counts and the week's tallies."""
    assert h.num_agents(r) == o.num_agents(), f"{what}: population size {h.num_agents(r)} != {o.num_agents()}"
    for name in attr_names():
        g, c = h.download_attr(r, name), o.get_attr(name)
        if not np.array_equal(g, c):
            bad = np.flatnonzero(g != c)
            raise AssertionError(f"{what}: attribute '{name}' differs at {bad.size} of {g.size} agents; first pos {bad[0]}: gpu {g[bad[0]]!r} oracle {c[bad[0]]!r}")
    for l in range(3):
        (ge, gs), (ce, cs) = h.get_edgelist(r, l), o.get_edgelist(l)
        assert np.array_equal(ge, ce), f"{what}: edge list of layer {l} differs ({ge.shape} vs {ce.shape})"
        assert np.array_equal(gs, cs), f"{what}: partnership start times of layer {l} differ"
        # dat$temp act counts are defined from acts_msm to the end of that week (kissing / rimming counts are lazy)
        assert not acts or np.array_equal(h.get_actcounts(r, l), o.get_actcounts(l)), f"{what}: act counts of layer {l} differ"
    gc, cc = h.counters(r, at, at)[0], o.counters(at)
    if not np.array_equal(gc, cc):
        bad = np.flatnonzero(gc != cc)
        raise AssertionError(f"{what}: epi tallies differ at counters {bad.tolist()[:10]}: gpu {gc[bad][:10]} oracle {cc[bad][:10]}")


def random_inject(rng, n_rep: int, uid_cap: int, e_cap, form_cap: int):
    """Injected-draw tables from an RNG unrelated to Philox (stands for R's runif() stream)."""
    from xgc_msm_b200 import abi
    stride = abi.inject_size(uid_cap, e_cap, form_cap)
    u = rng.random((n_rep, stride))
    u = np.clip(u, 1e-12, 1 - 1e-12)
    return u, stride
