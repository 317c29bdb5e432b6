"""CPU tests (`-m "not gpu"`): pin the oracle to everything the reference DOES hold for this path, and check the
host logic and the C-ABI surface.  The reference has no tests or golden vectors of its own, so the pins
are: the calibration targets recomputed from inst/cal/calval_targets.r literals, the prior box and
its size (burnin/cal/sim1_01_lhs_params.r:22-119), published Philox4x32-10 known answers, and libm's exp."""
import ctypes as C
import json
import math
import os
import re
import sys

import numpy as np
import pytest

from common import K, P, ROOT, load_oracle, make_case
from oracle import oracle as O
from xgc_msm_b200 import synth, targets as T


# ---- arithmetic contract ------------------------------------------------------------------------------------------
def _philox(ctr, key):
    out = (C.c_uint32 * 4)()
    O.lib.orc_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return list(out)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert _philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_exp_within_one_ulp_of_libm():
    for x in np.concatenate([np.linspace(-40, 40, 40001), np.linspace(-1e-3, 1e-3, 2001)]):
        a, b = O.lib.orc_exp(float(x)), math.exp(float(x))
        assert abs(a - b) <= 2.3e-16 * b


def test_bernoulli_follows_r_inversion():
    # rbinom(1,1,p): p <= .5 -> (u >= 1-p); p > .5 -> (u < p)   
    assert O.lib.orc_bern(0.95, 0.1) == 1 and O.lib.orc_bern(0.85, 0.1) == 0
    assert O.lib.orc_bern(0.65, 0.7) == 1 and O.lib.orc_bern(0.75, 0.7) == 0
    assert O.lib.orc_bern(0.5, 0.0) == 0 and O.lib.orc_bern(0.999999, 1.0) == 1
    u = np.random.default_rng(1).random(20000)
    for p in (0.03, 0.5, 0.83):
        assert abs(np.mean([O.lib.orc_bern(float(v), p) for v in u]) - p) < 0.01


def test_poisson_cdf_walk_matches_scipy():
    from scipy.stats import poisson
    for mu in (0.2, 1.7, 9.5):
        for u in np.linspace(0.001, 0.999, 97):
            assert O.lib.orc_pois(float(u), mu, 255) == int(poisson.ppf(u, mu))
    assert O.lib.orc_pois(0.999999999, 10.0, 32) == 32 and O.lib.orc_pois(0.5, 0.0, 32) == 0


# ---- golden values recomputed from the reference's literals ----------------------------------------------------------
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "caltargets.json")))


def test_any_of_n_is_the_distribution_of_n_per_act_draws():
    """one Bernoulli(1 - (1-t)^n) stands for n per-act Bernoulli(t) draws of which only "any success" is
    observed.  Closed form within a few ulp, exact at the edges, and the frequency of "any of n uniforms < t" matches."""
    rng = np.random.default_rng(3)
    for t in (0.0, 1e-6, 0.01, 0.25, 0.5, 0.9, 1.0):
        for n in (0, 1, 2, 3, 7, 16, 31, 32, 63):
            got = O.lib.orc_any_of_n(t, n)
            want = 1.0 - (1.0 - t) ** n
            assert abs(got - want) <= 1e-14, (t, n, got, want)
    assert O.lib.orc_any_of_n(0.3, 0) == 0.0 and O.lib.orc_any_of_n(0.3, 1) == 1.0 - (1.0 - 0.3)
    t, n, N = 0.07, 9, 400000
    per_act = (rng.random((N, n)) < t).any(axis=1).mean()
    one_draw = np.mean([O.lib.orc_bern(float(u), O.lib.orc_any_of_n(t, n)) for u in rng.random(20000)])
    p = 1.0 - (1.0 - t) ** n
    assert abs(per_act - p) < 4 * np.sqrt(p * (1 - p) / N) and abs(one_draw - p) < 4 * np.sqrt(p * (1 - p) / 20000)


def test_caltargets_match_reference_literals():
    got = {(t, s): (v, lo, hi) for t, v, s, lo, hi in T.caltargets()}
    assert len(got) == len(GOLD["rows"]) == 30
    for row in GOLD["rows"]:
        v, lo, hi = got[(row["target"], row["subgroup"])]
        assert v == pytest.approx(row["value"], abs=5e-5), row
        if row["ll95"] is not None:
            assert lo == pytest.approx(row["ll95"], abs=1.5e-3) and hi == pytest.approx(row["ul95"], abs=1.5e-3), row


def test_prior_box_matches_reference():
    assert len(P.PRIORS) == 62                                            # length(priors)
    assert sum(1 for _, lo, hi in P.PRIORS if lo == hi) == 7              # 7 fixed -> 55 free inputs
    d = {k: (lo, hi) for k, lo, hi in P.PRIORS}
    assert d["U2RGC_PROB"] == (0, 0.35) and d["PHAR_GC_DURAT_NOTX"] == (3, 20) and d["HIV_TRANS_RR_RGC"] == (1.2, 6.4)
    assert d["HIV_RX_HALT_PROB_BLACK"][0] == pytest.approx(1 - (1 - 0.464) ** (1 / 52))
    # the reference's own source text, when present in this container, must agree line by line
    src = "/root/reference/burnin/cal/sim1_01_lhs_params.r"
    if os.path.exists(src):
        txt = open(src).read()
        found = re.findall(r'pvec\(([^,]+),\s*([^,]+),\s*"(\w+)"\)', txt)
        assert len(found) == 62
        for lo, hi, name in found:
            lo_v, hi_v = (eval(s.strip().replace("^", "**")) for s in (lo, hi))
            assert d[name] == pytest.approx((lo_v, hi_v)), name
    env = P.prior_midpoint_env()
    p = P.params_from_env(env)
    assert P.get(p, "rgc_tx_recov_pr").tolist() == pytest.approx([1 - 0.06, 0.5, 1.0])     # c(1 - INFPR_WK1, 0.5, 1)
    assert P.get(p, "a_rate")[0] == pytest.approx(P.A_RATE_BASE + 1.285 / 20000)      # synthetic marginal exit rate + the reference tweak (sim1_02_lhs.r:66, sim1_batch1.sh:9)
    lhs = P.lhs_envs(50, 1971)
    u = np.array([[(e[k] - lo) / (hi - lo) if hi > lo else 0.5 for k, lo, hi in P.PRIORS] for e in lhs])
    free = [j for j, (_, lo, hi) in enumerate(P.PRIORS) if hi > lo]
    for j in free:                                                       # Latin property: one point per stratum
        assert sorted(np.floor(u[:, j] * 50).astype(int).tolist()) == list(range(50))


def test_params_def_consistent_across_python_c_and_cuda():
    assert P.NPARAM == O.NPARAM == K.NPARAM
    off, cnt = C.c_int(), C.c_int()
    from xgc_msm_b200 import abi
    for name, (o, c, _) in P.PARAM_INDEX.items():
        assert abi.lib.xgc_param_index(name.encode(), C.byref(off), C.byref(cnt)) == 0
        assert (off.value, cnt.value) == (o, c), name
    for nm in ("NTABLE", "NCONTROL", "NCOUNTER", "NTARGET", "NSTREAM", "M_ALL", "S_EDGE0", "K_PATH", "T_NET_DUR"):
        assert getattr(K, nm) == O.const(nm)
    assert synth.NTABLE == K.NTABLE and synth.T_GLM_CONDOO == K.T_GLM_CONDOO and synth.T_NET_DEG == K.T_NET_DEG
    assert [P.MODULE_BIT[m] for m in P.MODULE_ORDER] == [getattr(K, "M_" + m.upper()) for m in P.MODULE_ORDER]
    assert len(T.device_target_names()) == K.NTARGET


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    from xgc_msm_b200 import abi
    hdr = open(os.path.join(ROOT, "include", "xgc", "abi.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(xgc_\w+)\s*\(", hdr, flags=re.M))
    declared -= {"xgc_stream_width", "xgc_stream_entities"}              # static inline helpers
    assert len(declared) >= 40
    for sym in sorted(declared):
        assert hasattr(abi.lib, sym), f"libxgc.so does not export {sym}"
    assert abi.lib.xgc_abi_version() == K.ABI_VERSION
    # no GPU here: creation must fail loudly, never fall back to a CPU path
    if not __import__("torch").cuda.is_available():
        with pytest.raises(abi.XgcError, match="no CUDA device|CPU fallback"):
            abi.Handle(1, 128, [8, 8, 8], 8)


def test_inject_layout_agrees_between_python_and_c():
    from xgc_msm_b200 import abi
    j = abi.make_inject(None, 300, [40, 50, 10], 16)
    assert j.stride == sum(abi.stream_entities(s, 300, [40, 50, 10], 16) * abi.stream_width(s) for s in range(K.NSTREAM))
    # the oracle reads draw (stream, entity, k) at off(stream) + entity*width + k: plant one value and watch it being used
    case = make_case(200, 3, network_mode="surrogate")
    o = load_oracle(case, 1, 0)
    stride = abi.inject_size(case["cap"], case["e_cap"], 64)
    u = np.full(stride, 0.5)
    off = sum(abi.stream_entities(s, case["cap"], case["e_cap"], 64) * abi.stream_width(s) for s in range(K.S_AGENT1))
    u[off + 17 * abi.stream_width(K.S_AGENT1) + 0] = 1.0 - 1e-9              # uid 17, natural-mortality draw -> dies
    o.step(2, K.M_AGING | K.M_DEPARTURE, O.make_inject(u, stride, case["cap"], case["e_cap"], 64))
    gone = sorted(set(range(200)) - set(o.get_attr("uid").astype(int)))
    assert 17 in gone and all(g == 17 or case["attrs"]["age"][g] + 7 / 365 >= 65 for g in gone)   # others only by ageing out


# ---- oracle behaviour ------------------------------------------------------------------------------------------------
def test_oracle_module_invariants():
    case = make_case(1500, 5, midcourse=True, network_mode="surrogate")
    o = load_oracle(case, 1971, 0)
    n0 = o.num_agents()
    for at in range(2, 14):
        before = o.num_agents()
        o.step(at, K.M_AGING | K.M_DEPARTURE)
        c = o.counters(at)
        dead = before - o.num_agents()
        assert 0 <= dead <= c[K.K_DEP_GEN] + c[K.K_DEP_AIDS]
        for l in range(3):                                               # delete_vertices: ids stay inside 1..n, tail < head
            el, _ = o.get_edgelist(l)
            assert el.size == 0 or (el.min() >= 1 and el.max() <= o.num_agents() and np.all(el[:, 0] < el[:, 1]))
        o.step(at, K.M_ALL & ~(K.M_AGING | K.M_DEPARTURE))
        c = o.counters(at)
        assert c[K.K_NUM * K.NCELL:(K.K_NUM + 1) * K.NCELL].sum() == o.num_agents()
        for s, site in enumerate(("rGC", "uGC", "pGC")):
            st, sy, tx, it = (o.get_attr(site + x) for x in ("", ".sympt", ".tx", ".infTime"))
            assert np.all(sy <= st) and np.all(tx <= st) and np.all((it != -1) == (st == 1))
            assert c[(K.K_INUM_RGC + s) * K.NCELL:(K.K_INUM_RGC + s + 1) * K.NCELL].sum() == st.sum()
        assert np.all(o.get_attr("tx.status") <= o.get_attr("diag.status")) and np.all(o.get_attr("diag.status") <= o.get_attr("status"))
        assert c[K.K_PATH:K.K_PATH + K.NPATH].sum() == sum(c[(K.K_INCID_RGC + s) * K.NCELL:(K.K_INCID_RGC + s + 1) * K.NCELL].sum() for s in range(3))
    age = o.get_attr("age")
    assert age.min() >= 18 and age.max() < 65.2 and abs(o.num_agents() - n0) < 0.1 * n0
    assert np.all(np.diff(o.get_attr("uid")) > 0)                       # arrivals are appended, order is kept


def test_oracle_control_flags_change_behaviour():
    tot = {}
    for proto in ("base", "symptomatic", "cdc", "universal"):
        case = make_case(1500, 9, midcourse=True, network_mode="surrogate", stiScreeningProtocol=proto)
        o = load_oracle(case, 1971, 0)
        t = np.zeros(K.NCOUNTER, dtype=np.int64)
        for at in range(2, 28):
            o.step(at, K.M_ALL); t += o.counters(at)
        tot[proto] = t
    v = {k: t[K.K_VISITS] - t[K.K_VISITS_SYMPT] for k, t in tot.items()}
    assert v["symptomatic"] == 0 and v["base"] > 0 and v["cdc"] > 0
    tested = {k: t[K.K_TESTED:K.K_TESTED + 3].sum() / max(t[K.K_VISITS], 1) for k, t in tot.items()}
    assert tested["universal"] == pytest.approx(3.0) and tested["base"] < 3.0
    # kissing / rimming routes off -> no p2p / p2r / r2p infections (sim1_02_lhs.r:220-221)
    case = make_case(1500, 9, midcourse=True, network_mode="surrogate", transRoute_Kissing=False, transRoute_Rimming=False)
    o = load_oracle(case, 1971, 0); t = np.zeros(K.NCOUNTER, dtype=np.int64)
    for at in range(2, 20):
        o.step(at, K.M_ALL); t += o.counters(at)
    assert t[K.K_PATH + K.NPATH - 1] == 0 and t[K.K_PATH + 1] == 0 and t[K.K_PATH + 5] == 0 and t[K.K_PATH + 0] > 0


def test_epi_series_and_target_pipeline():
    case = make_case(2000, 13, network_mode="surrogate")
    o = load_oracle(case, 1971, 0, t_cap=80)
    for at in range(2, 70):
        o.step(at, K.M_ALL)
    from xgc_msm_b200 import abi
    epi = {v: o.epi(v, 2, 69) for v in abi.names("epi")}                  # every name the library advertises parses
    assert np.allclose(epi["num"], epi["num.B"] + epi["num.H"] + epi["num.O"] + epi["num.W"])
    assert np.allclose(epi["incid.gc"], epi["incid.rgc"] + epi["incid.ugc"] + epi["incid.pgc"])
    assert np.allclose(epi["incid.rgc"], sum(epi[f"incid.{r}.rgc"] for r in "BHOW"))
    assert np.allclose(epi["incid.gc"], sum(epi[k] for k in epi if re.search(r"(r|u|p)2", k)))      # 98_summarize_sims.r:75
    assert np.allclose(epi["i.num.dx.B"], sum(epi[f"i.num.dx.B.age{a}"] for a in range(1, 6)))
    # device target columns == sim0.0_setup.r pipeline (derived measures -> NA->0 -> tail mean) on the same series
    tm = T.tail_means(T.derived_measures(epi), 52)
    tg = dict(zip(T.device_target_names(), o.targets(69, 52)))
    for k in ("num", "i.prev", "ir100.pop", "i.prev.dx.inf.B", "i.prev.dx.inf.age3", "cc.vsupp.W", "cc.vsupp.age2", "prepCov.H",
              "prop.rect.tested", "prob.uGC.tested", "prev.gc", "prev.pgc", "ir100.rgc"):
        assert tg[k] == pytest.approx(tm[k], rel=1e-12, abs=1e-15), k
    # calculate_targets.r family on the same series
    assert T.target_prob_prep_byrace(epi, 52)["prep.cov.W"] == pytest.approx(tm["prepCov.W"], rel=1e-9)
    assert set(T.target_prop_anatsites_tested(epi, 52)) == {"prop.rect.tested", "prop.ureth.tested", "prop.phar.tested"}
    assert set(T.target_prob_gcpos_tested_anatsites(epi, 52)) == {"prob.rGC.tested", "prob.uGC.tested", "prob.pGC.tested"}
    assert T.target_prob_vls(epi, 52, "race")["prob.vsupp.B"] == pytest.approx(tm["cc.vsupp.B"])
    d = T.get_absdiff("ct_hiv_prev", np.array([0.10, 0.124, 0.30]))
    assert d["output_within_5pts"].tolist() == [1, 1, 0] and d["output_in_targcl"].tolist() == [0, 1, 0]
    assert T.get_absdiff("ct_hiv_incid_per100pop", np.array([0.9]))["output_within_5pts"].tolist() == [1]
    assert T.select_by_popsize(np.array([19100, 20900, 21100])).tolist() == [True, True, False]


def test_r_shim_compiles_against_the_r_api_and_binds_only_declared_symbols():
    """rshim/xgc_r.c (the .Call binding a maintainer of the reference would build with `R CMD SHLIB`) cannot be built
    here (no R), so it is syntax-checked against a stub of the R API, and every xgc_* function it calls must be
    declared in include/xgc/abi.h."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(root, "tests", "stubs"), "-I" + os.path.join(root, "include"),
                        os.path.join(root, "rshim", "xgc_r.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(root, "rshim", "xgc_r.c")).read()
    hdr = open(os.path.join(root, "include", "xgc", "abi.h")).read()
    used = set(re.findall(r"\b(xgc_[a-z_]+)\s*\(", src))
    declared = set(re.findall(r"\b(xgc_[a-z_]+)\s*\(", hdr))
    assert used and used <= declared, used - declared
    mods = open(os.path.join(root, "rshim", "modules.R")).read()
    calls = set(re.findall(r'\.Call\("(xgcR_[a-z_]+)"', mods))
    defined = set(re.findall(r"^SEXP (xgcR_[a-z_]+)\(", src, flags=re.M))
    assert calls and calls <= defined, calls - defined


def test_host_mirror_matches_the_reference_driver_scripts():
    """tests/golden/driver_contract.json is parsed from the reference's own driver scripts (make_driver_contract.py):
    param_msm argument names and the env var each reads, init_msm literals, the control_msm module slots in the order
    written, the option flags, and the screening arms.  The host-side mirror must agree with all of it."""
    from xgc_msm_b200 import scenarios
    root = os.path.dirname(os.path.abspath(__file__))
    ref = json.load(open(os.path.join(root, "golden", "driver_contract.json")))
    # modules: same slots, same order (initialize.FUN is the constructor, not a weekly module)
    slots = [m["slot"] for m in ref["modules"]]
    assert slots[0] == "initialize" and slots[1:] == P.MODULE_ORDER
    assert [m["fun"] for m in ref["modules"]][-2:] == ["stitrans_msm_rand", "prevalence_msm"]
    assert [P.MODULE_BIT[m] for m in P.MODULE_ORDER] == [1 << k for k in range(17)] and K.M_ALL == (1 << 17) - 1
    # option flags: the defaults of the mirror are the reference's settings
    f = ref["control_flags"]
    ctl = P.control_msm()
    idx = P.CONTROL_INDEX
    assert ctl.flags[idx["transRoute_Kissing"]] == int(f["transRoute_Kissing"]) == 1
    assert ctl.flags[idx["transRoute_Rimming"]] == int(f["transRoute_Rimming"]) == 1
    assert ctl.flags[idx["cdcExposureSite_Kissing"]] == int(f["cdcExposureSite_Kissing"]) == 0
    assert f["gcUntreatedRecovDist"] == "geom" and ctl.flags[idx["gcUntreatedRecovDist"]] == K.RECOV_GEOM
    assert f["stiScreeningProtocol"] == "base" and ctl.flags[idx["stiScreeningProtocol"]] == K.SCREEN_BASE
    # every epidemic argument of the param_msm() call is a parameter of the ABI (site-specific triplets are one vector)
    known = set(P.PARAM_INDEX) | set(P._SITE_ALIASES)
    for a in ref["param_msm"]:
        if a["name"] in ("netstats", "epistats"):
            continue                                   # fitted inputs: xgc_set_tables
        n = a["name"].replace(".", "_")
        assert n in known, f"param_msm argument {a['name']} (sim1_02_lhs.r:{a['line']}) has no ABI parameter"
    # the env vars the call reads are exactly the 62 calibration inputs plus the arrival-rate tweak
    envs = {e for a in ref["param_msm"] for e in a["env"]}
    assert envs == {p[0] for p in P.PRIORS} | {"ARRIVE_RATE_ADD_PER20K"}
    assert ref["init_msm"] == {"prev.ugc": 0.05, "prev.rgc": 0.05, "prev.pgc": 0.05}
    i0 = P.init_msm()
    assert i0 == ref["init_msm"]
    # and the R-style call itself goes through the mirror with the reference's own argument names
    kw = {a["name"].replace(".", "_"): 0.5 for a in ref["param_msm"] if a["name"] not in ("netstats", "epistats")}
    vec = P.param_msm(**kw)
    assert vec.shape == (K.NPARAM,) and P.get(vec, "gc_ntx_int").tolist() == [0.5, 0.5, 0.5]
    # screening arms
    arms = {a["name"].replace("sti_", ""): (a["scenario"], a["kiss_exposure"]) for a in ref["screening_arms"] if a["runtype"] == "scenario"}
    assert arms == {k: tuple(v) for k, v in scenarios.ARMS.items()}


def test_rngstreams_mrg32k3a_known_answers_and_jumps():
    """xgc_msm_b200/rngstreams.py (rlecuyer-compatible uniforms for injected-draw tables): published known answers of
    MRG32k3a with the default seed, jump-ahead == stepping, stream / substream spacing, and use as an inject table."""
    from xgc_msm_b200 import abi
    from xgc_msm_b200 import rngstreams as RS
    g = RS.RngStream()
    u = [g.rand_u01() for _ in range(3)]
    assert 0 <= u[1] - 0.3185275653 < 1e-10 and 0 <= u[2] - 0.3091860155 < 1e-10 and abs(u[0] - 0.12701112) < 1e-8   # published to 10 digits (truncated)
    assert all(0.0 < v < 1.0 for v in u)
    a, b = RS.RngStream(), RS.RngStream()
    ref = [a.rand_u01() for _ in range(1500)]
    assert np.array_equal(b.uniform(1500), np.array(ref))                      # vector path == scalar path
    c = RS.RngStream(); c.advance(1000)
    assert c.rand_u01() == ref[1000]                                           # O(log n) jump == 1000 steps
    # streams are 2^127 apart, substreams 2^76: jumping in two halves gives the same state
    s = RS.RngStream.create_streams(3)
    d = RS.RngStream(); d.advance(1 << 126); d.advance(1 << 126)
    assert d.cg == s[1].ig and s[0].ig == list(RS.DEFAULT_SEED)
    e = RS.RngStream(s[1].ig); e.advance(1 << 127)
    assert e.cg == s[2].ig
    f = RS.RngStream(); f.reset_next_substream()
    h = RS.RngStream(); h.advance(1 << 76)
    assert f.cg == h.cg
    f.rand_u01(); f.reset_start_stream()
    assert f.cg == list(RS.DEFAULT_SEED)
    with pytest.raises(ValueError):
        RS.RngStream([0, 0, 0, 1, 2, 3])
    # as an injected-draw table: same stream -> same week, different stream -> different week
    case = make_case(150, 9, network_mode="surrogate")
    stride = abi.inject_size(case["cap"], case["e_cap"], 64)
    res = []
    for k in (0, 0, 1):
        o = load_oracle(case, 1, 0)
        tab = RS.inject_table(RS.RngStream.create_streams(2)[k], 1, stride)
        assert tab.shape == (1, stride)
        o.step(2, K.M_ALL, O.make_inject(tab[0], stride, case["cap"], case["e_cap"], 64))
        res.append(o.counters(2).copy())
    assert np.array_equal(res[0], res[1]) and not np.array_equal(res[0], res[2])


def test_product_path_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device xgc_create returns a non-zero status with a message, the Python Handle
    raises, and bench.py's CUDA arm exits with an error instead of computing on the host."""
    import subprocess
    import torch
    from xgc_msm_b200 import abi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the failure path cannot be exercised")
    with pytest.raises(abi.XgcError) as ei:
        abi.Handle(1, 256, [64, 64, 64], 8)
    assert "xgc_create failed" in str(ei.value) and len(str(ei.value)) > 25
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout) and '"value"' not in r.stdout


def test_oracle_degree_conditioned_formation():
    """Surrogate simnet_msm of the oracle (abi.h XGC_T_NET_DEGW / XGC_T_NET_RISKW): hard weights are respected exactly -- nobody
    with a main partner forms another main tie, casual degree stops at 2, risk group 1 has no one-off contacts."""
    from common import load_oracle, make_case
    from xgc_msm_b200 import synth
    case = make_case(3000, 21, network_mode="surrogate")
    case["els"] = [np.zeros((0, 2), np.int32)] * 3; case["starts"] = [np.zeros(0, np.int32)] * 3
    t = synth.make_tables()
    W = t[synth.T_NET_DEGW:synth.T_NET_DEGW + 27].reshape(3, 3, 3)
    W[0, 1:, :] = 0.0; W[1, :, 2] = 0.0; t[synth.T_NET_RISKW] = 0.0
    case["tables"] = t
    o = load_oracle(case, 1971, 0, t_cap=48)
    for at in range(2, 42):
        o.step(at, K.M_ALL)
    n = o.get_attr("race").size
    deg = [np.bincount(o.get_edgelist(l)[0].ravel() - 1, minlength=n) for l in range(3)]
    assert deg[0].max() == 1 and deg[0].sum() > 100
    assert deg[1].max() == 2 and deg[1].sum() > 200
    assert deg[2].sum() > 10 and deg[2][o.get_attr("risk.grp") == 1].sum() == 0
