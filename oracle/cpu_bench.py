"""TEST/BENCH INFRASTRUCTURE — times the CPU oracle on the host cores.

Used only by bench.py (`cpu_baseline` leg and `--impl reference`).  Parallelism mirrors the reference's own mechanism:
one single-threaded process per replicate per core (SLURM job array `--array=1-1000`, `-n 1`, NSIMS=1:
/root/reference/burnin/cal/sim1/sim1_batch1.sh:2-9).  Kept free of torch/CUDA imports so workers start fast.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _build_replicates(n_agents: int, ids, seed: int, spec_kw=None, philox_seed=None):
    """Same synthetic workload as bench.py's GPU arm: population `rid % n_distinct`, LHS parameter row `rid`."""
    from oracle.oracle import Oracle
    from xgc_msm_b200 import workload as W
    spec = W.SweepSpec(n_agents=n_agents, seed=seed, **(spec_kw or {}))
    out = []
    for rid in ids:
        case = W.replicate_case(spec, rid)
        o = Oracle(spec.n_cap_oracle, spec.e_cap, spec.t_cap_oracle, seed=seed if philox_seed is None else philox_seed, replicate_id=rid)
        o.set_params(case["params"]); o.set_control(case["control"]); o.set_tables(case["tables"])
        o.set_population(n_agents, 1)
        for k, v in case["attrs"].items():
            o.set_attr(k, v)
        for l in range(3):
            o.set_edgelist(l, case["els"][l], case["starts"][l])
        out.append(o)
    return out


def _batch_worker(args):
    n_agents, ids, seed, weeks, mask = args
    reps = _build_replicates(n_agents, ids, seed)
    t0 = time.perf_counter()
    aw = 0
    for o in reps:
        for at in range(2, 2 + weeks):
            o.step(at, mask)
            aw += o.num_agents()
    return aw, time.perf_counter() - t0


def run_batch(n_agents: int, n_replicates: int, weeks: int, seed: int, mask: int, cores: int | None = None):
    """One-shot sample: n_replicates replicates x weeks, spread over `cores` single-threaded processes.
    Returns (agent_weeks, wall_seconds, cores)."""
    cores = cores or os.cpu_count() or 1
    cores = min(cores, n_replicates)
    chunks = [list(range(c, n_replicates, cores)) for c in range(cores)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_noop, range(cores))               # start the workers before the clock
        t0 = time.perf_counter()
        res = pool.map(_batch_worker, [(n_agents, ids, seed, weeks, mask) for ids in chunks])
        wall = time.perf_counter() - t0
    # building the replicates is inside `wall`; subtract nothing: report the steady stepping rate from worker clocks
    aw = sum(r[0] for r in res)
    step_wall = max(r[1] for r in res)
    return aw, step_wall, cores, wall


def _noop(_):
    return 0


def _persistent_worker(conn, n_agents, ids, seed, mask):
    reps = _build_replicates(n_agents, ids, seed)
    conn.send(("ready", len(reps)))
    while True:
        msg = conn.recv()
        if msg is None:
            break
        at = msg
        n = 0
        for o in reps:
            o.step(at, mask)
            n += o.num_agents()
        conn.send(n)
    conn.close()


class PersistentPool:
    """`cores` workers, each owning a fixed set of replicates; step(at) advances every replicate by one week."""

    def __init__(self, n_agents: int, n_replicates: int, seed: int, mask: int, cores: int | None = None):
        self.cores = min(cores or os.cpu_count() or 1, n_replicates)
        ctx = mp.get_context("spawn")
        self.conns, self.procs = [], []
        for c in range(self.cores):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_persistent_worker, args=(b, n_agents, list(range(c, n_replicates, self.cores)), seed, mask), daemon=True)
            p.start()
            self.conns.append(a); self.procs.append(p)
        for a in self.conns:
            assert a.recv()[0] == "ready"

    def step(self, at: int) -> int:
        for a in self.conns:
            a.send(at)
        return sum(a.recv() for a in self.conns)

    def close(self):
        for a in self.conns:
            try:
                a.send(None)
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)


def _targets_worker(args):
    n_agents, ids, seed, weeks, tailn, mask, spec_kw, philox_seed = args
    reps = _build_replicates(n_agents, ids, seed, spec_kw, philox_seed)
    out = []
    for o in reps:
        for at in range(2, 2 + weeks):
            o.step(at, mask)
        out.append(o.targets(1 + weeks, tailn))
        o.close()
    return ids, np.stack(out)


def run_targets(n_agents: int, n_replicates: int, weeks: int, tailn: int, seed: int, mask: int, spec_kw=None,
                philox_seed=None, cores: int | None = None) -> np.ndarray:
    """[n_replicates, NTARGET] tail-window targets of the oracle (statistical tests)."""
    cores = min(cores or os.cpu_count() or 1, n_replicates)
    chunks = [list(range(c, n_replicates, cores)) for c in range(cores)]
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_targets_worker, [(n_agents, ids, seed, weeks, tailn, mask, spec_kw, philox_seed) for ids in chunks])
    out = None
    for ids, t in res:
        if out is None:
            out = np.empty((n_replicates, t.shape[1]))
        out[ids] = t
    return out


# ---------------------------------------------------------------------------------------------------------------------
# full-state runs for the BASELINE-size parity tests (tests/test_gpu_fullsize_parity.py)
# ---------------------------------------------------------------------------------------------------------------------
def host_edges(seed: int, rid: int, at: int, layer: int, n_alive: int, e_cap):
    """Deterministic stand-in for what host-side simnet_msm returns for (replicate, week, layer): a ragged list of
    1-based positions among the living (tail < head) and partnership start weeks.  Pure function of its arguments, so
    the GPU driver and the oracle workers build identical lists without talking to each other."""
    rng = np.random.default_rng([seed, rid, at, layer])
    ne = int(rng.integers(int(e_cap[layer] * 0.55), int(e_cap[layer] * 0.74)))
    a = rng.integers(1, n_alive + 1, size=ne)
    b = rng.integers(1, n_alive, size=ne)
    b = np.where(b >= a, b + 1, b)
    el = np.stack([np.minimum(a, b), np.maximum(a, b)], axis=1).astype(np.int32)
    st = (at - rng.integers(0, 300 if layer < 2 else 1, size=ne)).astype(np.int32)
    return el, st


def _state_worker(args):
    n_agents, ids, seed, weeks, mask, spec_kw, tailn, attr_names, host_net, overrides = args
    reps = _build_replicates(n_agents, ids, seed, spec_kw)
    from oracle.oracle import NCOUNTER
    from xgc_msm_b200.params import MODULE_BIT as MB
    pre = MB["aging"] | MB["departure"] | MB["arrival"] | MB["hivtest"] | MB["hivtx"] | MB["hivprogress"] | MB["hivvl"]
    out = {}
    for rid, o in zip(ids, reps):
        if overrides and rid in overrides:
            ov = overrides[rid]
            if "params" in ov:
                o.set_params(ov["params"])
            if "control" in ov:
                o.set_control(ov["control"])
        cnt = np.zeros((weeks, NCOUNTER), np.uint32)
        alive = np.zeros(weeks, np.int64)
        for at in range(2, 2 + weeks):
            if host_net:
                o.step(at, pre)
                n = o.num_agents()
                for l in range(3):
                    el, st = host_edges(seed, rid, at, l, n, o.e_cap)
                    o.set_edgelist(l, el, st if l < 2 else None)
                o.step(at, mask & ~pre & ~MB["resim_nets"])
            else:
                o.step(at, mask)
            cnt[at - 2] = o.counters(at)
            alive[at - 2] = o.num_agents()
        out[rid] = dict(counters=cnt, alive=alive, attrs={k: o.get_attr(k) for k in attr_names},
                        els=[o.get_edgelist(l) for l in range(3)], acts=[o.get_actcounts(l) for l in range(3)],
                        targets=o.targets(1 + weeks, tailn), status=o.status())
        o.close()
    return out


def run_states(n_agents: int, ids, weeks: int, tailn: int, seed: int, mask: int, attr_names, spec_kw=None, host_net: bool = False,
               overrides=None, cores: int | None = None) -> dict:
    """{rid: full final state + weekly tallies} of the oracle for the given global replicate ids of the sweep workload."""
    ids = list(ids)
    cores = max(1, min(cores or os.cpu_count() or 1, len(ids)))
    chunks = [ids[c::cores] for c in range(cores)]
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_state_worker, [(n_agents, ch, seed, weeks, mask, spec_kw, tailn, list(attr_names), host_net, overrides) for ch in chunks])
    out = {}
    for r in res:
        out.update(r)
    return out
