"""The calibration-sweep workload of BASELINE.json configs[1]: R replicates x N agents, every replicate with its own
parameter vector drawn by Latin hypercube over the reference's prior box (burnin/cal/sim1_01_lhs_params.r:22-137),
kissing and rimming routes on, "geom" recovery, "base" screening (sim1_02_lhs.r:219-229), initial site prevalence 0.05
(sim1_02_lhs.r:180-184), seed 1971 (sim1_01_lhs_params.r:118).  The fitted netest/netstats/epistats are LFS pointers,
so populations and networks are synthetic (synth.py).  Shared by bench.py's GPU arm, its CPU arms and the tests so
that all of them run the *same* replicates (replicate `rid` is fully determined by (seed, rid)).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from functools import lru_cache

import numpy as np

from . import params as P
from . import synth

LHS_BLOCK = 1000           # replicates per Latin hypercube (one hypercube per 1000-replicate sweep, as in the reference)
N_DISTINCT = 4             # distinct starting populations (the reference starts every replicate from the same netest fit)


@dataclass
class SweepSpec:
    n_agents: int = 10000
    seed: int = 1971
    network_mode: str = "surrogate"
    weeks_between_compactions: int = 64
    lhs: bool = True               # False: every replicate uses the prior-midpoint parameters (exchangeable replicates)
    n_distinct: int = N_DISTINCT   # distinct starting populations
    n_cap: int = field(init=False)
    e_cap: list = field(init=False)

    def __post_init__(self):
        self.n_cap = synth.capacity(self.n_agents, self.weeks_between_compactions)
        self.e_cap = synth.edge_capacity(self.n_agents)
        # the oracle deletes departed rows immediately, so it needs no tombstone head-room; keep identical caps anyway
        self.n_cap_oracle = self.n_cap
        self.t_cap_oracle = 4096


@lru_cache(maxsize=16)
def _lhs_block(block: int, seed: int):
    return P.lhs_envs(LHS_BLOCK, seed + 7 * block)


@lru_cache(maxsize=16)
def _population(idx: int, n: int, seed: int):
    base = P.params_from_env(P.prior_midpoint_env(), n_init=n)
    a = synth.init_population(n, seed + idx, P.init_msm(0.05, 0.05, 0.05), base)
    els, starts = synth.init_network(a, seed + idx)
    return a, els, starts


def replicate_params(spec: SweepSpec, rid: int) -> np.ndarray:
    if not spec.lhs:
        return P.params_from_env(P.prior_midpoint_env(), n_init=spec.n_agents)
    env = _lhs_block(rid // LHS_BLOCK, spec.seed)[rid % LHS_BLOCK]
    return P.params_from_env(env, n_init=spec.n_agents)


def replicate_case(spec: SweepSpec, rid: int) -> dict:
    a, els, starts = _population(rid % spec.n_distinct, spec.n_agents, spec.seed)
    return dict(n=spec.n_agents, params=replicate_params(spec, rid), control=P.control_msm(network_mode=spec.network_mode).flags,
                tables=synth.make_tables(), attrs=a, els=els, starts=starts, cap=spec.n_cap, e_cap=spec.e_cap, at0=1)


def load_sweep(h, spec: SweepSpec, rid0: int = 0, ids=None):
    """Fill handle `h` with replicates rid0 .. rid0 + R - 1 (or with the global replicate ids `ids`, any subset of the sweep:
    xgc_set_replicate_ids).  The N_DISTINCT starting populations go through the attribute/edge-list ABI once; the other
    replicates are device-to-device forks of them (xgc_fork)."""
    R = h.n_replicates
    rids = [rid0 + r for r in range(R)] if ids is None else [int(x) for x in ids]
    assert len(rids) == R
    if ids is not None:
        h.set_replicate_ids(np.asarray(rids, dtype=np.int32))
    h.set_tables(synth.make_tables())
    first = {}
    for r in range(R):
        rid = rids[r]
        idx = rid % spec.n_distinct
        if idx not in first:
            a, els, starts = _population(idx, spec.n_agents, spec.seed)
            h.set_population(r, spec.n_agents, 1)
            for k, v in a.items():
                h.upload_attr(r, k, v)
            for l in range(3):
                h.set_edgelist(r, l, els[l], starts[l])
            first[idx] = r
        else:
            h.fork_from(h, first[idx], r)
    h.set_params(np.stack([replicate_params(spec, rid) for rid in rids]))
    h.set_control(P.control_msm(network_mode=spec.network_mode).flags)
    h.sync()
