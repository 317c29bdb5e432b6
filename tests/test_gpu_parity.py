"""GPU parity tests: the CUDA path, driven through the C ABI (libxgc.so), against the CPU oracle on the same seeded
inputs.  Bar: BIT-EXACT for every attribute vector, edge list, act count and epi tally (integer/byte/index work), and
for the double-precision targets (same IEEE operation order on both sides; tolerance 0).

The oracle is 'parity unpinned' with respect to the R modules (they are not in /root/reference; see
oracle/xgc_oracle.h) — these tests pin the CUDA path to the oracle, and tests/test_oracle_cpu.py pins the oracle to
what the reference does hold (calibration-target literals, prior box, published Philox vectors).
"""
import numpy as np
import pytest

from common import K, P, compare_state, load_handle, load_oracle, make_case, random_inject

pytestmark = pytest.mark.gpu

SEED = 1971   # the reference's own LHS seed (burnin/cal/sim1_01_lhs_params.r:118)


def test_device_arithmetic_contract():
    """xgc_exp and Philox4x32-10 on the device agree bit for bit with the oracle's restatement."""
    import ctypes as C
    from oracle import oracle as O
    from xgc_msm_b200 import abi
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-50, 50, 20000), rng.uniform(-1, 1, 20000), [0.0, -0.0, 700.0, -700.0, 1e-300, -745.0, 710.0]])
    y = np.empty_like(x)
    assert abi.lib.xgc_debug_exp(x.ctypes.data_as(C.POINTER(C.c_double)), y.ctypes.data_as(C.POINTER(C.c_double)), x.size, 0) == 0
    ref = np.array([O.lib.orc_exp(float(v)) for v in x])
    assert np.array_equal(y.view(np.uint64), ref.view(np.uint64))
    ctr = rng.integers(0, 2**32, size=(4096, 4), dtype=np.uint64).astype(np.uint32)
    out = np.empty_like(ctr)
    assert abi.lib.xgc_debug_philox(1971, 7, ctr.ctypes.data_as(C.POINTER(C.c_uint32)), out.ctypes.data_as(C.POINTER(C.c_uint32)), 4096, 0) == 0
    key = (C.c_uint32 * 2)(1971, 7)
    for i in range(0, 4096, 64):
        o4 = (C.c_uint32 * 4)()
        O.lib.orc_philox4x32_10((C.c_uint32 * 4)(*[int(v) for v in ctr[i]]), key, o4)
        assert list(o4) == [int(v) for v in out[i]]


@pytest.mark.parametrize("recov,protocol,kiss", [("geom", "base", False), ("pois", "cdc", True), ("geom", "universal", False),
                                                 ("pois", "symptomatic", False), ("geom", "cdc", False)])
def test_full_step_native_draws(recov, protocol, kiss):
    """All 17 modules, surrogate network, native Philox draws, 3 replicates x 10 weeks, compared every week."""
    n, weeks = 3000, 10
    cases = [make_case(n, 100 + r, midcourse=True, gcUntreatedRecovDist=recov, stiScreeningProtocol=protocol,
                       cdcExposureSite_Kissing=kiss, network_mode="surrogate") for r in range(3)]
    h = load_handle(cases, SEED, replicate_offset=40)
    orcs = [load_oracle(c, SEED, 40 + r) for r, c in enumerate(cases)]
    for r, o in enumerate(orcs):
        compare_state(h, r, o, 1, f"after upload r={r}", acts=False)
    for at in range(2, 2 + weeks):
        h.step(at)
        for r, o in enumerate(orcs):
            o.step(at, K.M_ALL)
            compare_state(h, r, o, at, f"{recov}/{protocol} week {at} r={r}")
    tg = h.targets(1 + weeks, 5)
    for r, o in enumerate(orcs):
        assert np.array_equal(tg[r].view(np.uint64), o.targets(1 + weeks, 5).view(np.uint64)), "targets differ"
        assert h.status(r) == o.status() == 0
    # the comparison above is only meaningful if the branches actually fired
    tot = h.counters(0, 2, 1 + weeks).astype(np.int64).sum(0)
    G = K.NCELL
    fired = {"departures": tot[K.K_DEP_GEN], "aids deaths": tot[K.K_DEP_AIDS], "arrivals": tot[K.K_NNEW], "hiv tests": tot[K.K_HIVTESTS],
             "hiv infections": tot[K.K_INCID_HIV * G:(K.K_INCID_HIV + 1) * G].sum(), "prep users": tot[K.K_PREPCURR * G:(K.K_PREPCURR + 1) * G].sum(),
             "rectal infections": tot[K.K_INCID_RGC * G:(K.K_INCID_RGC + 1) * G].sum(), "urethral infections": tot[K.K_INCID_UGC * G:(K.K_INCID_UGC + 1) * G].sum(),
             "pharyngeal infections": tot[K.K_INCID_PGC * G:(K.K_INCID_PGC + 1) * G].sum(), "visits": tot[K.K_VISITS],
             "positives": tot[K.K_POS:K.K_POS + 3].sum(), "treatments": tot[K.K_TXSTART:K.K_TXSTART + 3].sum(),
             "anal acts": tot[K.K_NACTS], "oral acts": tot[K.K_NACTS + 1], "edges": tot[K.K_NEDGES:K.K_NEDGES + 3].min()}
    for p in range(K.NPATH):
        fired[f"pathway {p}"] = tot[K.K_PATH + p]
    assert all(v > 0 for v in fired.values()), fired


def test_module_by_module_contract():
    """The drop-in contract: each module called on its own, `dat <- FUN(dat, at)`, in the reference's order
    (burnin/cal/sim1_02_lhs.r:201-218), compared after every call."""
    case = make_case(2500, 7, midcourse=True, network_mode="surrogate", stiScreeningProtocol="cdc")
    h = load_handle([case], SEED)
    o = load_oracle(case, SEED, 0)
    for at in (2, 3, 4):
        for name in P.MODULE_ORDER:
            bit = P.MODULE_BIT[name]
            h.step(at, bit)
            o.step(at, bit)
            compare_state(h, 0, o, at, f"week {at} after {name}", acts=P.MODULE_ORDER.index(name) >= P.MODULE_ORDER.index("acts"))


def test_injected_draws_bit_exact():
    """north_star test 1: with uniform draws injected from outside (an RNG unrelated to Philox, standing for R's
    runif stream), agent and edge state transitions are bit-exact."""
    from oracle.oracle import make_inject as omake
    from xgc_msm_b200 import abi
    n = 1200
    cases = [make_case(n, 300 + r, midcourse=True, network_mode="surrogate", gcUntreatedRecovDist="pois" if r else "geom",
                       stiScreeningProtocol="base" if r else "cdc") for r in range(2)]
    h = load_handle(cases, SEED)
    orcs = [load_oracle(c, SEED, r) for r, c in enumerate(cases)]
    rng = np.random.default_rng(99)
    uid_cap, e_cap, form_cap = cases[0]["cap"] + 256, cases[0]["e_cap"], 256
    for at in range(2, 8):
        u, stride = random_inject(rng, 2, uid_cap, e_cap, form_cap)
        h.step(at, K.M_ALL, abi.make_inject(u, uid_cap, e_cap, form_cap))
        for r, o in enumerate(orcs):
            o.step(at, K.M_ALL, omake(u[r], stride, uid_cap, e_cap, form_cap))
            compare_state(h, r, o, at, f"injected week {at} r={r}")
    # with injected draws every oral / rimming / kissing act is its own trial (abi.h xgc_stream_width): those pathways must have fired
    paths = sum(h.counters(r, 2, 7).astype(np.int64)[:, K.K_PATH:K.K_PATH + K.NPATH].sum(0) for r in range(2))
    assert paths[[1, 3, 4, 5, 6]].min() > 0, paths                  # p2r, p2u, u2p, r2p, p2p (u2r, r2u are the anal pathways)


def test_actlist_condoms_position():
    """condoms_msm / position_msm: the materialised act list (dat$temp$al) equals the oracle's row for row."""
    case = make_case(2000, 11, midcourse=True, network_mode="surrogate")
    h = load_handle([case], SEED)
    o = load_oracle(case, SEED, 0)
    pre = K.M_AGING | K.M_DEPARTURE | K.M_ARRIVAL | K.M_RESIM_NETS | K.M_ACTS
    for at in (2, 3):
        h.step(at, pre); o.step(at, pre)
        ga, ca = h.get_actlist(0, at), o.get_actlist(at)
        assert ga.shape == ca.shape and ga.shape[0] > 1000
        assert np.array_equal(ga, ca)
        rest = K.M_ALL & ~pre
        h.step(at, rest); o.step(at, rest)
        compare_state(h, 0, o, at, f"week {at}")


def test_run_graph_compaction_and_sharding_invariance():
    """xgc_run (CUDA-graph replay, periodic slot compaction) == week-by-week xgc_step == oracle; and a replicate's
    trajectory does not depend on which shard / local index it runs at (Philox is keyed by the global replicate id)."""
    n, T = 2000, 40
    cases = [make_case(n, 500 + r, network_mode="surrogate") for r in range(4)]
    ha = load_handle(cases, SEED, t_cap=T + 2, compact_every=16)
    ha.run(2, T)
    hb = load_handle(cases[2:], SEED, t_cap=T + 2, replicate_offset=2)       # second "GPU": replicates 2,3
    for at in range(2, T + 1):
        hb.step(at)
        if at % 7 == 0:
            hb.compact()
    o = load_oracle(cases[3], SEED, 3, t_cap=T + 2)
    for at in range(2, T + 1):
        o.step(at, K.M_ALL)
    compare_state(ha, 3, o, T, "graph run vs oracle")
    compare_state(hb, 1, o, T, "sharded stepwise run vs oracle")
    ta, tb = ha.targets(T, 20), hb.targets(T, 20)
    assert np.array_equal(ta[2:].view(np.uint64), tb.view(np.uint64))
    assert np.array_equal(ta[3].view(np.uint64), o.targets(T, 20).view(np.uint64))
    for v in ("num", "prev.gc", "incid.gc", "i.prev", "incid.B.rgc", "incid.age.2.pgc", "prop.rect.tested", "prob.uGC.tested",
              "cc.vsupp.W", "ir100", "prepCurr.H", "incid.u2rgc", "i.num.dx.B.age1"):
        g, c = ha.epi(3, v, 2, T), o.epi(v, 2, T)
        assert np.array_equal(np.nan_to_num(g, nan=-1.0), np.nan_to_num(c, nan=-1.0)), v


def test_snapshot_restore_fork():
    """Checkpoint at the end of burn-in, fork into screening arms (inst/analysis02_screening/02.00_epi.r:40-76,292-296)."""
    n = 1500
    case = make_case(n, 21, network_mode="surrogate")
    h = load_handle([case, case], SEED, t_cap=40)
    h.run(2, 12)
    snap = h.snapshot()
    h.run(13, 20)
    ref = h.counters(0, 13, 20).copy()
    h.restore(snap)
    h.run(13, 20)
    assert np.array_equal(h.counters(0, 13, 20), ref), "restore + rerun must reproduce the same trajectory"
    # fork replicate 0 of the burn-in into replicate 1 and switch that arm to universal screening
    h.restore(snap)
    h.fork_from(h, 0, 1)
    ctl = case["control"].copy(); ctl[K.C_stiScreeningProtocol] = 3
    h.set_control(ctl, 1)
    h.run(13, 20)
    assert np.array_equal(h.counters(0, 13, 20), ref)
    tested_base = h.counters(0, 13, 20)[:, K.K_TESTED:K.K_TESTED + 3].sum()
    tested_univ = h.counters(1, 13, 20)[:, K.K_TESTED:K.K_TESTED + 3].sum()
    assert tested_univ > tested_base


def test_edge_cases_empty_and_ragged():
    """Empty layers, a replicate with no edges at all, a tiny population, and a replicate of a different size."""
    big = make_case(1000, 1, midcourse=True, network_mode="surrogate")
    small = make_case(40, 2, midcourse=True, network_mode="host")
    small["cap"], small["e_cap"] = big["cap"], big["e_cap"]
    small["els"] = [np.zeros((0, 2), np.int32)] * 3
    small["starts"] = [np.zeros(0, np.int32)] * 3
    h = load_handle([big, small], SEED)
    ob, os_ = load_oracle(big, SEED, 0), load_oracle(small, SEED, 1)
    for at in range(2, 8):
        h.step(at)
        ob.step(at, K.M_ALL); os_.step(at, K.M_ALL)
        compare_state(h, 0, ob, at, f"big week {at}")
        compare_state(h, 1, os_, at, f"small/no-edge week {at}")


def test_large_population_100k_plus():
    """BASELINE configs[3]/[4] shape: a 170k-agent replicate (k_rank / pos2slot path of the surrogate network, long
    per-layer edge loops, parallel slot compaction) next to a second replicate, 4 weeks, bit-exact vs the oracle."""
    case = make_case(170000, 77, network_mode="surrogate")
    h = load_handle([case, case], SEED, t_cap=8, replicate_offset=5)
    orcs = [load_oracle(case, SEED, 5 + r, t_cap=8) for r in range(2)]
    for at in (2, 3, 4, 5):
        h.step(at)
        for r, o in enumerate(orcs):
            o.step(at, K.M_ALL)
            compare_state(h, r, o, at, f"170k agents week {at} r={r}")
        if at == 3:
            # slot compaction of a large replicate (parallel gathers through pos2slot) is invisible in R positions
            h.compact()
            for r, o in enumerate(orcs):
                compare_state(h, r, o, at, f"170k agents after compaction r={r}", acts=False)


def test_config5_one_million_agents():
    """BASELINE configs[4]: ONE replicate of 1,000,000 agents (~0.5M ties), the shape bench.py --replicates 1 --agents 1000000
    times: every per-replicate step is multi-CTA (look-back edge compaction and formation, rank plane, parallel slot
    compaction).  Three weeks + a slot compaction, bit-exact vs the oracle: all attributes, edge lists, act counts, tallies."""
    case = make_case(1000000, 5, network_mode="surrogate")
    h = load_handle([case], SEED, t_cap=8, replicate_offset=2)
    o = load_oracle(case, SEED, 2, t_cap=8)
    for at in (2, 3, 4):
        h.step(at)
        o.step(at, K.M_ALL)
        compare_state(h, 0, o, at, f"1M agents week {at}")
        if at == 3:
            h.compact()
            compare_state(h, 0, o, at, "1M agents after compaction", acts=False)
    assert h.status(0) == 0 and h.num_agents(0) > 990000
    h.close()


def test_host_network_weekly_contract_all_replicates():
    """The end-to-end contract bench.py's e2e arm times (INTEGRATION.md "weekly host <-> device traffic"): pre-network
    modules -> living counts to the host -> three ragged edge lists for EVERY replicate in one call each (1-based
    positions among the living, as tergmLite returns them) -> remaining modules -> the week's tallies for every replicate.
    Bit-exact against the oracle driven through its per-replicate edge-list setter."""
    n, weeks, R = 2500, 8, 4
    cases = [make_case(n, 300 + r, midcourse=True, network_mode="host", stiScreeningProtocol="cdc" if r == 1 else "base") for r in range(R)]
    h = load_handle(cases, SEED, replicate_offset=7)
    orcs = [load_oracle(c, SEED, 7 + r) for r, c in enumerate(cases)]
    pre = K.M_AGING | K.M_DEPARTURE | K.M_ARRIVAL | K.M_HIVTEST | K.M_HIVTX | K.M_HIVPROGRESS | K.M_HIVVL
    post = K.M_ALL & ~pre & ~K.M_RESIM_NETS
    rng = np.random.default_rng(5)
    e_cap = cases[0]["e_cap"]
    for at in range(2, 2 + weeks):
        h.step(at, pre)
        for o in orcs:
            o.step(at, pre)
        alive = h.num_agents_all()
        assert alive.tolist() == [o.num_agents() for o in orcs]
        for l in range(3):
            ne = rng.integers(e_cap[l] // 3, e_cap[l] // 2, size=R).astype(np.int32)
            if at == 4:
                ne[2] = 0                                   # an empty layer in one replicate
            if at == 5 and l == 2:
                ne[0] = e_cap[l]                            # a layer at capacity
            stride = int(ne.max()) + 3                      # ragged: rows longer than any list
            el = np.zeros((R, 2, stride), dtype=np.int32)
            st = np.zeros((R, stride), dtype=np.int32)
            use_start = not (at == 6 and l == 1)            # NULL start weeks = "formed this week"
            for r in range(R):
                a = rng.integers(1, alive[r] + 1, size=ne[r]); b = rng.integers(1, alive[r], size=ne[r])
                b = np.where(b >= a, b + 1, b)
                el[r, 0, :ne[r]] = np.minimum(a, b); el[r, 1, :ne[r]] = np.maximum(a, b)
                st[r, :ne[r]] = at - rng.integers(0, 150, size=ne[r])
                orcs[r].set_edgelist(l, el[r, :, :ne[r]].T, st[r, :ne[r]] if use_start else None)
            h.set_edgelists_all(l, el.ctypes.data, ne.ctypes.data, st.ctypes.data if use_start else 0, stride=stride)
            h.sync()                                        # el / st are reused next iteration
        h.step(at, post)
        call = h.counters_all(at)
        for r, o in enumerate(orcs):
            o.step(at, post)
            compare_state(h, r, o, at, f"host network week {at} r={r}")
            assert np.array_equal(call[r], o.counters(at))
    for r, o in enumerate(orcs):
        assert h.status(r) == o.status() == 0
    tot = h.counters(1, 2, 1 + weeks).astype(np.int64).sum(0)
    assert tot[K.K_NACTS] > 0 and tot[K.K_PATH:K.K_PATH + K.NPATH].sum() > 0 and tot[K.K_VISITS] > 0
    # a position outside 1..n_alive is reported, not dereferenced
    bad = np.zeros((R, 2, 4), dtype=np.int32); bad[:, 0, :] = 1; bad[:, 1, :] = 2; bad[1, 1, 2] = int(alive[1]) + 50
    h.set_edgelists_all(0, bad.ctypes.data, np.full(R, 4, np.int32).ctypes.data, 0, stride=4); h.sync()
    assert h.status(1) & K.ST_BAD_EDGE and not (h.status(0) & K.ST_BAD_EDGE)


def test_capacity_overflow_is_flagged_and_contained():
    """A replicate whose arrivals outgrow the slot capacity (and whose formation outgrows a layer's edge capacity) sets
    its status flags, stays within its capacities and keeps stepping; the replicate next to it in the same handle is
    untouched -- still bit-exact against the oracle."""
    n = 600
    ok = make_case(n, 11, midcourse=True, network_mode="surrogate")
    hot = make_case(n, 12, midcourse=True, network_mode="surrogate")
    hot["params"] = P.param_msm(hot["params"], a_rate=0.08)                 # ~48 arrivals a week into ~10 % head-room
    e_cap = [int(len(ok["els"][0]) * 1.02) + 4, ok["e_cap"][1], ok["e_cap"][2]]   # main layer: almost no head-room
    for c in (ok, hot):
        c["e_cap"] = e_cap
    h = load_handle([ok, hot], SEED, t_cap=40)
    o = load_oracle(ok, SEED, 0, t_cap=40)
    cap = ok["cap"]
    for at in range(2, 30):
        h.step(at)
        o.step(at, K.M_ALL)
        if o.status() == 0:
            compare_state(h, 0, o, at, f"neighbour of the overflowing replicate, week {at}")
        assert h.num_agents(1) <= cap
        for l in range(3):
            assert h.get_edgelist(1, l)[0].shape[0] <= e_cap[l]
    assert h.status(1) & K.ST_AGENT_OVERFLOW
    assert h.num_agents(1) > n                                             # it did grow until the cap stopped it
    assert not (h.status(0) & K.ST_AGENT_OVERFLOW)
    c = h.counters(1, 2, 29).astype(np.int64)
    G = K.NCELL
    assert (c[:, K.K_NUM * G:(K.K_NUM + 1) * G].sum(1) <= cap).all() and c[-1, K.K_NACTS] > 0      # still a living epidemic


def test_abi_error_behaviour():
    """The reference-facing error contract: every misuse returns a non-zero status with a message through
    xgc_last_error, nothing throws across the ABI or kills the process, and the handle keeps working afterwards."""
    from xgc_msm_b200.abi import XgcError
    case = make_case(300, 21, network_mode="surrogate")
    h = load_handle([case], SEED, t_cap=16)
    o = load_oracle(case, SEED, 0, t_cap=16)

    def fails(f, *words):
        with pytest.raises(XgcError) as ei:
            f()
        msg = str(ei.value)
        assert all(w in msg for w in words), msg

    fails(lambda: h.step(0), "outside")                                    # at = 0 is not a week
    fails(lambda: h.step(16), "t_cap")                                     # beyond the counter rows
    fails(lambda: h.run(5, 99), "week range")
    fails(lambda: h.download_attr(3, "age"), "replicate")                  # only replicate 0 exists
    fails(lambda: h.download_attr(0, "no.such.attr"), "no.such.attr")
    fails(lambda: h.upload_attr(0, "age", np.zeros(7)), "length")          # wrong vector length
    fails(lambda: h.set_edgelist(0, 5, np.zeros((0, 2), np.int32), np.zeros(0, np.int32)), "layer")
    fails(lambda: h.set_edgelist(0, 0, np.array([[1, 9999]], np.int32), np.zeros(1, np.int32)))   # endpoint beyond the population
    fails(lambda: h.epi(0, "not.a.series", 2, 3), "not.a.series")
    fails(lambda: h.targets(3, 50), "tail")                                # window longer than the history
    fails(lambda: h.set_edgelists_all(7, 0, 0, 0, stride=4), "layer")
    # the handle is still usable and still exact
    for at in (2, 3):
        h.step(at); o.step(at, K.M_ALL)
        compare_state(h, 0, o, at, f"after the misuse, week {at}")
    h.close()
    h.close()                                                               # closing twice is harmless
