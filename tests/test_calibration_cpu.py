"""Host logic of the calibration loop and of the epi export (no GPU): LHS, out_vs_targ, mase selection, prior narrowing.
Reference behaviour: sim1_01_lhs_params.r:118-137, sim0.0_setup.r:299-379,513-528, sim3_01_analyze.r:120-192,
sim_clean.r:55-101."""
import numpy as np
import pytest

from xgc_msm_b200 import calibration as CAL
from xgc_msm_b200 import params as P
from xgc_msm_b200 import targets as T


def test_lhs_one_point_per_stratum_and_inside_the_box():
    n = 50
    x = CAL.draw_lhs(P.PRIORS, n, seed=3)
    assert x.shape == (n, 62)
    for j, (name, lo, hi) in enumerate(P.PRIORS):
        assert x[:, j].min() >= lo and x[:, j].max() <= hi
        if hi > lo:
            strata = np.floor((x[:, j] - lo) / (hi - lo) * n).astype(int)
            assert sorted(strata.tolist()) == list(range(n)), name
        else:
            assert np.all(x[:, j] == lo)
    envs = CAL.inputs_to_envs(P.PRIORS, x)
    assert len(envs) == n and envs[7]["U2RGC_PROB"] == x[7, [p[0] for p in P.PRIORS].index("U2RGC_PROB")]
    P.params_from_env(envs[0])           # every LHS row is a complete param_msm() call


def _synthetic_targets(n, rng, noise):
    names = T.device_target_names()
    tm = np.zeros((n, len(names)))
    for j, nm in enumerate(names):
        if nm in T.TARGET_OF_COLUMN:
            tg, sub = T.TARGET_OF_COLUMN[nm]
            tm[:, j] = T.target_value(tg, sub) + noise * rng.standard_normal(n)
        elif nm == "num":
            tm[:, j] = 10000 + 100 * rng.standard_normal(n)
    return names, tm


def test_out_vs_targ_mase_and_selection():
    rng = np.random.default_rng(0)
    names, tm = _synthetic_targets(200, rng, 0.08)
    for j, nm in enumerate(names):                 # simulation 17 hits every target exactly
        if nm in T.TARGET_OF_COLUMN:
            tm[17, j] = T.target_value(*T.TARGET_OF_COLUMN[nm])
    tm[5, names.index("num")] = 12000.0            # outside the +-5 % population window
    keep = T.select_by_popsize(tm[:, names.index("num")], 10000.0, 0.05)
    assert not keep[5] and keep.sum() == 199
    ovt = CAL.out_vs_targ(tm, names, keep)
    assert set(ovt) == set(T.TARGET_OF_COLUMN)
    m = CAL.mase(ovt)
    assert m.shape == (199,)
    kept = np.flatnonzero(keep)
    assert kept[CAL.best_fits(m, 10)[0]] == 17 and m.min() == pytest.approx(0.0, abs=1e-12)
    # the default varpats take HIV prevalence/incidence, vsupp and dx by race, dx by age, GC positivity -- not PrEP, not prop.*.tested
    one = {k: v for k, v in ovt.items() if k in ("prepCov.B", "prop.rect.tested")}
    with pytest.raises(ValueError):
        CAL.mase(one)
    q = CAL.select_mase_quantile(m, 0.05)
    assert 9 <= q.sum() <= 11 and q[np.argmin(m)]
    sel = CAL.select_within(ovt, [f"prob.{g}.tested" for g in T.GCLABS])
    ref = np.ones(199, bool)
    for g in T.GCLABS:
        ref &= np.abs(tm[keep, names.index(f"prob.{g}.tested")] - T.target_value("ct_prop_anatsite_pos", g)) <= 0.05
    assert np.array_equal(sel, ref)


def test_input_quantiles_by_race_group():
    rng = np.random.default_rng(1)
    x = CAL.draw_lhs(P.PRIORS, 400, seed=5)
    selB = rng.random(400) < 0.2
    selAll = rng.random(400) < 0.3
    new = CAL.input_quantiles(P.PRIORS, x, {"B": selB, "*": selAll})
    idx = {p[0]: j for j, p in enumerate(P.PRIORS)}
    jb = idx["HIV_RX_INIT_PROB_BLACK"]
    assert new[jb][1:] == (pytest.approx(np.quantile(x[selB, jb], 0.25)), pytest.approx(np.quantile(x[selB, jb], 0.75)))
    ju = idx["U2RGC_PROB"]
    assert new[ju][1:] == (pytest.approx(np.quantile(x[selAll, ju], 0.25)), pytest.approx(np.quantile(x[selAll, ju], 0.75)))
    jh = idx["HIV_RX_INIT_PROB_HISP"]               # no "H" selection -> falls back to "*"
    assert new[jh][1] == pytest.approx(np.quantile(x[selAll, jh], 0.25))
    jw = idx["SCALAR_HIV_TRANS_PROB_WHITE"]          # degenerate prior (1, 1) stays
    assert new[jw] == P.PRIORS[jw]
    for (n0, lo0, hi0), (n1, lo1, hi1) in zip(P.PRIORS, new):
        assert n0 == n1 and lo0 <= lo1 <= hi1 <= hi0
    few = np.zeros(400, bool); few[:2] = True       # too few selected rows: keep the prior
    assert CAL.input_quantiles(P.PRIORS, x, few) == list(P.PRIORS)


def test_keepvars_are_series_the_library_serves():
    from xgc_msm_b200 import abi
    from xgc_msm_b200.export import KEEPVARS
    known = set(abi.names("epi"))
    assert len(KEEPVARS) == 81 and not [k for k in KEEPVARS if k not in known]


def test_selection_constants_match_the_reference_scripts():
    """mase variable patterns, LHS seed / size and the population filter as the reference's scripts state them
    (tests/golden/driver_contract.json, parsed from sim3_01_analyze.r:169-176, sim1_01_lhs_params.r:118-120,
    sim0.0_setup.r:317-318)."""
    import json, os, re
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "driver_contract.json")))
    names = [n for n in T.device_target_names() if n in T.TARGET_OF_COLUMN]
    theirs = re.compile("|".join(ref["mase_varpats"]))
    ours = re.compile(CAL.MASE_VARPATS)
    assert [n for n in names if theirs.search(n)] == [n for n in names if ours.search(n)] and len([n for n in names if ours.search(n)]) == 18
    assert ref["lhs_seed"] == 1971 and ref["lhs_n"] == 5000 and ref["pop_N"] == 20000
    import inspect
    assert inspect.signature(P.lhs_envs).parameters["seed"].default == ref["lhs_seed"]
    assert inspect.signature(T.select_by_popsize).parameters["pop_n"].default == float(ref["pop_N"])
    assert inspect.signature(T.select_by_popsize).parameters["tol"].default == 0.05
