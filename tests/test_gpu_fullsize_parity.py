"""Bit-exact GPU-vs-oracle parity AT THE SIZES BASELINE.json NAMES, on the bench workload itself.

bench.py times `workload.SweepSpec(n_agents=10000)` with per-replicate Latin-hypercube parameters over the reference's
prior box (burnin/cal/sim1_01_lhs_params.r:22-137); the reference's own population size is 20 000
(burnin/cal/sim1_02_lhs.r:64-66, inst/cal/sim0.0_setup.r:315-318).  These tests draw random blocks of global replicate ids
from that sweep, run them for two years through the paths the bench times -- `xgc_run` (CUDA-graph replay, two lanes, slot
compaction) and the weekly host-network contract (`xgc_step(pre)` / `xgc_set_edgelists_all` / `xgc_step(post)`) -- and
compare every weekly tally, the final attribute vectors, edge lists, act counts and the targets with the CPU oracle run on
the same replicates by a process pool.  Also: BASELINE configs[0] (one replicate, ~10k agents, 52 weeks, compared every
week) and the corners of the 62-input prior box.
"""
import numpy as np
import pytest

from common import K, P, attr_names, compare_state, load_handle, load_oracle

pytestmark = pytest.mark.gpu
SEED = 1971
PRE = None


def _pre_mask():
    return K.M_AGING | K.M_DEPARTURE | K.M_ARRIVAL | K.M_HIVTEST | K.M_HIVTX | K.M_HIVPROGRESS | K.M_HIVVL


def _compare_final(h, r, ref, what):
    """h replicate r against an oracle state dict of cpu_bench.run_states."""
    assert h.num_agents(r) == ref["alive"][-1], f"{what}: population {h.num_agents(r)} != {ref['alive'][-1]}"
    for name in attr_names():
        g, c = h.download_attr(r, name), ref["attrs"][name]
        if not np.array_equal(g, c):
            bad = np.flatnonzero(g != c)
            raise AssertionError(f"{what}: attribute '{name}' differs at {bad.size} of {g.size} agents; first {bad[0]}: gpu {g[bad[0]]!r} oracle {c[bad[0]]!r}")
    for l in range(3):
        (ge, gs), (ce, cs) = h.get_edgelist(r, l), ref["els"][l]
        assert np.array_equal(ge, ce) and np.array_equal(gs, cs), f"{what}: edge list / start weeks of layer {l} differ"
        assert np.array_equal(h.get_actcounts(r, l), ref["acts"][l]), f"{what}: act counts of layer {l} differ"


def _compare_weekly(h, r, ref, weeks, what):
    g = h.counters(r, 2, 1 + weeks)
    if not np.array_equal(g, ref["counters"]):
        wk, col = np.argwhere(g != ref["counters"])[0]
        raise AssertionError(f"{what}: weekly tallies first differ at week {wk + 2}, counter {col}: gpu {g[wk, col]} oracle {ref['counters'][wk, col]}")


@pytest.mark.parametrize("n_agents,blocks,per_block,compact_every", [(10000, 4, 6, 64), (20000, 3, 4, 40)])
def test_sweep_replicates_run_graph_bit_exact(n_agents, blocks, per_block, compact_every):
    """Random blocks of the 1000-replicate LHS sweep through xgc_run for 104 weeks == the oracle, bit for bit."""
    from oracle import cpu_bench
    from xgc_msm_b200 import workload as W
    from xgc_msm_b200.abi import Handle
    weeks, tail = 104, 52
    spec = W.SweepSpec(n_agents=n_agents, seed=SEED, weeks_between_compactions=compact_every)
    rng = np.random.default_rng(n_agents)
    starts = sorted(int(x) for x in rng.choice(1000 - per_block, size=blocks, replace=False))
    ids = [s + k for s in starts for k in range(per_block)]
    ref = cpu_bench.run_states(n_agents, ids, weeks, tail, SEED, K.M_ALL, attr_names(), spec_kw=dict(weeks_between_compactions=compact_every))
    fired = np.zeros(K.NCOUNTER, np.int64)
    for s in starts:
        h = Handle(per_block, spec.n_cap, spec.e_cap, weeks + 4, seed=SEED, replicate_offset=s, compact_every=compact_every)
        W.load_sweep(h, spec, rid0=s)
        h.run(2, 1 + weeks)
        tg = h.targets(1 + weeks, tail)
        for r in range(per_block):
            o = ref[s + r]
            what = f"{n_agents} agents, replicate {s + r}"
            _compare_weekly(h, r, o, weeks, what)
            _compare_final(h, r, o, what)
            assert np.array_equal(tg[r].view(np.uint64), o["targets"].view(np.uint64)), f"{what}: targets differ"
            assert h.status(r) == o["status"] == 0, (what, h.status(r), o["status"])
            fired += o["counters"].astype(np.int64).sum(0)
        h.close()
    G = K.NCELL
    assert fired[K.K_DEP_GEN] > 0 and fired[K.K_NNEW] > 0 and fired[K.K_HIVTESTS] > 0 and fired[K.K_VISITS] > 0
    assert fired[K.K_INCID_HIV * G:(K.K_INCID_HIV + 1) * G].sum() > 0 and fired[K.K_PATH:K.K_PATH + K.NPATH].min() > 0
    assert fired[K.K_TXSTART:K.K_TXSTART + 3].min() > 0 and fired[K.K_PREPCURR * G:(K.K_PREPCURR + 1) * G].sum() >= 0


def test_sweep_replicates_host_network_path_bit_exact():
    """The weekly contract bench.py's e2e arm times, on sampled replicates of the 10k sweep for 104 weeks: pre-network
    modules -> living counts -> three ragged edge lists per replicate (one xgc_set_edgelists_all per layer) -> remaining
    modules; slot compaction happens inside xgc_step on the schedule of xgc_run.  Bit-exact against the oracle driven with
    the same lists."""
    from oracle import cpu_bench
    from xgc_msm_b200 import workload as W
    from xgc_msm_b200.abi import Handle
    n_agents, weeks, tail, R, s0 = 10000, 104, 52, 8, 613
    spec = W.SweepSpec(n_agents=n_agents, seed=SEED, network_mode="host", weeks_between_compactions=32)
    ids = list(range(s0, s0 + R))
    ref = cpu_bench.run_states(n_agents, ids, weeks, tail, SEED, K.M_ALL, attr_names(), host_net=True,
                               spec_kw=dict(network_mode="host", weeks_between_compactions=32))
    h = Handle(R, spec.n_cap, spec.e_cap, weeks + 4, seed=SEED, replicate_offset=s0, compact_every=32)
    W.load_sweep(h, spec, rid0=s0)
    pre = _pre_mask()
    post = K.M_ALL & ~pre & ~K.M_RESIM_NETS
    stride = max(spec.e_cap)
    for at in range(2, 2 + weeks):
        h.step(at, pre)
        alive = h.num_agents_all()
        assert alive.tolist() == [int(ref[i]["alive"][at - 2]) for i in ids], f"living agents differ in week {at}"
        for l in range(3):
            el = np.zeros((R, 2, stride), np.int32); st = np.zeros((R, stride), np.int32); ne = np.zeros(R, np.int32)
            for r, rid in enumerate(ids):
                e, s = cpu_bench.host_edges(SEED, rid, at, l, int(alive[r]), spec.e_cap)
                ne[r] = e.shape[0]; el[r, 0, :ne[r]] = e[:, 0]; el[r, 1, :ne[r]] = e[:, 1]; st[r, :ne[r]] = s
            h.set_edgelists_all(l, el.ctypes.data, ne.ctypes.data, st.ctypes.data if l < 2 else 0, stride=stride)
            h.sync()
        h.step(at, post)
    tg = h.targets(1 + weeks, tail)
    for r, rid in enumerate(ids):
        what = f"host-network path, replicate {rid}"
        _compare_weekly(h, r, ref[rid], weeks, what)
        _compare_final(h, r, ref[rid], what)
        assert np.array_equal(tg[r].view(np.uint64), ref[rid]["targets"].view(np.uint64))
        assert h.status(r) == ref[rid]["status"] == 0
    h.close()


def test_host_network_longer_than_slot_headroom():
    """the step-by-step boundary compacts slots on xgc_run's schedule.  A replicate with ~6 arrivals a
    week and 384 slots of head-room would overflow within ~70 weeks if tombstones were never reclaimed; 300 weeks through
    xgc_step alone stay within capacity, flag nothing and match the oracle."""
    from common import make_case
    n, weeks = 1500, 300
    case = make_case(n, 77, midcourse=True, network_mode="surrogate")
    case["params"] = P.param_msm(case["params"], a_rate=0.001)
    case["cap"] = 1792
    h = load_handle([case], SEED, t_cap=weeks + 4, compact_every=16)
    o = load_oracle(case, SEED, 0, t_cap=weeks + 4)
    for at in range(2, 2 + weeks):
        h.step(at)
        o.step(at, K.M_ALL)
    assert h.status(0) == o.status() == 0
    compare_state(h, 0, o, 1 + weeks, "300 weeks of xgc_step with scheduled slot compaction")
    tot = h.counters(0, 2, 1 + weeks).astype(np.int64).sum(0)
    assert tot[K.K_NNEW] > 1792 - n                   # more arrivals than there ever was head-room
    h.close()


def test_config0_single_replicate_52_weeks_every_week():
    """BASELINE configs[0]: one replicate, ~10k agents, 52 weekly steps; full state compared after every week."""
    from xgc_msm_b200 import workload as W
    spec = W.SweepSpec(n_agents=10000, seed=SEED)
    case = W.replicate_case(spec, 17)
    h = load_handle([case], SEED, t_cap=60, replicate_offset=17)
    o = load_oracle(case, SEED, 17, t_cap=60)
    for at in range(2, 54):
        h.step(at)
        o.step(at, K.M_ALL)
        compare_state(h, 0, o, at, f"configs[0] week {at}")
    assert np.array_equal(h.targets(53, 52)[0].view(np.uint64), o.targets(53, 52).view(np.uint64))
    assert h.status(0) == o.status() == 0
    h.close()


def _corner_env(kind: str, rng=None):
    env = {}
    for name, lo, hi in P.PRIORS:
        env[name] = lo if kind == "lo" else hi if kind == "hi" else (lo if rng.random() < 0.5 else hi)
    return env


def test_prior_box_corners():
    """Every one of the 62 calibration inputs at the bottom of its prior, at the top (kissing / rimming 10 acts a week,
    transmission probabilities at their maxima, treated and untreated durations at their extremes) and at six random
    vertices of the box; 10k agents, 40 weeks each, through xgc_run.  Bit-exact, and no capacity / truncation flag: the
    XGC_MAX_ACTS = 32 truncation of a Poisson(10) draw has mass ~1e-9 (flagged only above XGC_TRUNC_TOL)."""
    from oracle import cpu_bench
    from xgc_msm_b200 import workload as W
    from xgc_msm_b200.abi import Handle
    n_agents, weeks, tail = 10000, 40, 20
    rng = np.random.default_rng(62)
    envs = [_corner_env("lo"), _corner_env("hi")] + [_corner_env("vertex", rng) for _ in range(6)]
    params = [P.params_from_env(e, n_init=n_agents) for e in envs]
    R, s0 = len(params), 200
    spec = W.SweepSpec(n_agents=n_agents, seed=SEED)
    ov = {s0 + r: {"params": params[r]} for r in range(R)}
    ref = cpu_bench.run_states(n_agents, list(ov), weeks, tail, SEED, K.M_ALL, attr_names(), overrides=ov)
    h = Handle(R, spec.n_cap, spec.e_cap, weeks + 4, seed=SEED, replicate_offset=s0, compact_every=16)
    W.load_sweep(h, spec, rid0=s0)
    h.set_params(np.stack(params))
    h.run(2, 1 + weeks)
    tg = h.targets(1 + weeks, tail)
    for r in range(R):
        what = f"prior-box corner {r}"
        _compare_weekly(h, r, ref[s0 + r], weeks, what)
        _compare_final(h, r, ref[s0 + r], what)
        assert np.array_equal(tg[r].view(np.uint64), ref[s0 + r]["targets"].view(np.uint64))
        assert h.status(r) == ref[s0 + r]["status"] == 0
    lo = ref[s0]["counters"].astype(np.int64).sum(0)
    hi = ref[s0 + 1]["counters"].astype(np.int64).sum(0)
    assert lo[K.K_PATH:K.K_PATH + K.NPATH].sum() == 0          # every transmission probability 0: no new GC infection
    assert hi[K.K_PATH:K.K_PATH + K.NPATH].min() > 0           # every pathway fires at the top corner
    # beyond the box: a weekly rate whose Poisson mass above XGC_MAX_ACTS is not negligible is flagged on both sides
    p = P.param_msm(params[1], kiss_rate_main=22.0)
    h.set_params(p, 1)
    from oracle.oracle import Oracle
    o = Oracle(spec.n_cap, spec.e_cap, 8)
    o.set_params(p)
    assert h.status(1) == o.status() == K.ST_RATE_TRUNCATED
    assert h.status(0) == 0
    h.close()


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_host_network_weekly_deltas(mode):
    """xgc_update_edgelists_all: the persistent layers (main, casual) are kept on the device and only the week's toggles
    cross the boundary -- removed ties as indices into the current list, formed ties as position pairs -- while the
    one-off layer is replaced.  Strict mode: bit-exact against the oracle driven with the full lists.  Fast mode: the
    device lists equal the host's lists every week (the update path does not depend on the arithmetic mode)."""
    from xgc_msm_b200 import workload as W
    from xgc_msm_b200.abi import Handle
    n_agents, weeks, R, s0 = 10000, 40, 4, 321
    spec = W.SweepSpec(n_agents=n_agents, seed=SEED, network_mode="host", weeks_between_compactions=16)
    cases = [W.replicate_case(spec, s0 + r) for r in range(R)]
    h = Handle(R, spec.n_cap, spec.e_cap, weeks + 4, seed=SEED, replicate_offset=s0, compact_every=16)
    W.load_sweep(h, spec, rid0=s0)
    h.set_mode(mode)
    orcs = [load_oracle(c, SEED, s0 + r, t_cap=weeks + 4) for r, c in enumerate(cases)]
    pre = _pre_mask()
    post = K.M_ALL & ~pre & ~K.M_RESIM_NETS
    rng = np.random.default_rng(7)
    dur = [250.0, 90.0]
    for at in range(2, 2 + weeks):
        h.step(at, pre)
        for o in orcs:
            o.step(at, pre)
        alive = h.num_agents_all()
        if mode == "strict":
            assert alive.tolist() == [o.num_agents() for o in orcs]
        lists = []
        for l in range(2):
            rows = []
            for r in range(R):
                # the host's view of the layer after this week's departures: in strict mode the oracle's own list; in fast mode
                # the device's (the two trajectories differ), which the delta then edits
                el, st = orcs[r].get_edgelist(l) if mode == "strict" else h.get_edgelist(r, l)
                rem = np.flatnonzero(rng.random(el.shape[0]) < 1.0 / dur[l]).astype(np.int32)
                if at == 5 and r == 1:
                    rem = np.arange(el.shape[0], dtype=np.int32)[:: 2]          # a big delta: every second tie dissolves
                nadd = int(rem.size) + int(rng.integers(0, 4))
                a = rng.integers(1, alive[r] + 1, size=nadd); b = rng.integers(1, alive[r], size=nadd); b = np.where(b >= a, b + 1, b)
                add = np.stack([np.minimum(a, b), np.maximum(a, b)], axis=1).astype(np.int32)
                keep = np.ones(el.shape[0], bool); keep[rem] = False
                full = np.vstack([el[keep], add]); fst = np.concatenate([st[keep], np.full(nadd, at, np.int32)])
                orcs[r].set_edgelist(l, full, fst)
                rows.append(np.concatenate([[rem.size, nadd], rem, add[:, 0], add[:, 1], np.full(nadd, at)]).astype(np.int32))
                lists.append((r, l, full, fst))
            stride = max(x.size for x in rows) + 5
            delta = np.zeros((R, stride), np.int32)
            for r, x in enumerate(rows):
                delta[r, :x.size] = x
            h.update_edgelists_all(l, delta.ctypes.data, stride, has_start=True)
            h.sync()
        ne2 = rng.integers(300, 500, size=R).astype(np.int32); stride2 = int(ne2.max())
        el2 = np.zeros((R, 2, stride2), np.int32)
        for r in range(R):
            a = rng.integers(1, alive[r] + 1, size=ne2[r]); b = rng.integers(1, alive[r], size=ne2[r]); b = np.where(b >= a, b + 1, b)
            el2[r, 0, :ne2[r]] = np.minimum(a, b); el2[r, 1, :ne2[r]] = np.maximum(a, b)
            orcs[r].set_edgelist(2, el2[r, :, :ne2[r]].T, None)
        h.set_edgelists_all(2, el2.ctypes.data, ne2.ctypes.data, 0, stride=stride2)
        h.sync()
        if mode == "fast" or at % 7 == 0:
            for r, l, full, fst in lists:
                ge, gs = h.get_edgelist(r, l)
                assert np.array_equal(ge, full) and np.array_equal(gs, fst), f"week {at} replicate {r} layer {l}: device list != host list"
        h.step(at, post)
        for r, o in enumerate(orcs):
            o.step(at, post)
            if mode == "strict":
                assert np.array_equal(h.counters(r, at, at)[0], o.counters(at)), f"week {at} replicate {r}: tallies differ"
    for r, o in enumerate(orcs):
        assert h.status(r) == 0
        if mode == "strict":
            compare_state(h, r, o, 1 + weeks, f"delta path, replicate {s0 + r}")
    # a removed index beyond the list is reported, not dereferenced
    bad = np.zeros((R, 8), np.int32); bad[2, 0] = 1; bad[2, 2] = 10 ** 6
    h.update_edgelists_all(0, bad.ctypes.data, 8, has_start=False); h.sync()
    assert h.status(2) & K.ST_BAD_EDGE and not (h.status(0) & K.ST_BAD_EDGE)
    h.close()
