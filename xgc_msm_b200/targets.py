"""Calibration targets and the target-comparison pipeline of the reference (host-side, tiny compute).

* `caltargets()` / `valtargets()` restate /root/reference/inst/cal/calval_targets.r:23-350 (the literature counts
  are literals there; est/caltargets.Rds itself is a git-LFS pointer).
* `target_*` mirror the exported family of R/calculate_targets.r:14-190 on a dict of epi series.
* `derived_measures`, `tail_means`, `get_absdiff` mirror inst/cal/sim0.0_setup.r:205-308, 333-379 — the pipeline the
  reference actually uses to select runs.  The device computes the same tail means in k_targets (xgc_targets()).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional

import numpy as np

RACES = ("B", "H", "O", "W")               # R/helpers.r:13
ANAT = ("rect", "ureth", "phar")
GCLABS = ("rGC", "uGC", "pGC")


def _binom_exact_ci(x: int, n: int, conf: float = 0.95):
    """Clopper-Pearson interval (binom.test(...)$conf.int, calval_targets.r:175-180)."""
    from scipy.stats import beta
    a = (1 - conf) / 2
    lo = 0.0 if x == 0 else float(beta.ppf(a, x, n - x + 1))
    hi = 1.0 if x == n else float(beta.ppf(1 - a, x + 1, n - x))
    return lo, hi


def caltargets() -> list:
    """Rows (target, value, subgroups, ll95, ul95) of est/caltargets.Rds, recomputed from calval_targets.r literals."""
    rows = []
    # HIV diagnosis (calval_targets.r:23-90): counts by race (B, H, W, Total) x age group 1..5; O = Total - others
    prev_n = {"B": [20700, 74300, 42200, 41700, 39700], "H": [10600, 45800, 42200, 42900, 31600],
              "W": [5000, 30200, 36100, 67800, 102700], "T": [38500, 162100, 130700, 164100, 184500]}
    dx_n = {"B": [11522, 53772, 36247, 39073, 38071], "H": [5169, 29838, 34094, 38988, 29934],
            "W": [2886, 21307, 30142, 62802, 98819], "T": [20873, 113483, 109109, 151648, 176821]}
    prev_n["O"] = [prev_n["T"][a] - sum(prev_n[r][a] for r in "BHW") for a in range(5)]
    dx_n["O"] = [dx_n["T"][a] - sum(dx_n[r][a] for r in "BHW") for a in range(5)]
    for r in RACES:
        rows.append(("ct_hivdx_pr_byrace", sum(dx_n[r]) / sum(prev_n[r]), r, None, None))
    for a in range(5):
        rows.append(("ct_hivdx_pr_byage", sum(dx_n[r][a] for r in RACES) / sum(prev_n[r][a] for r in RACES), f"age{a + 1}", None, None))
    # viral suppression (calval_targets.r:93-156)
    tot = {"B": [773 + 9381, 27792, 24205, 31274, 16437], "H": [316 + 3242, 16715, 22581, 24927, 11364],
           "O": [88 + 1046, 4264, 5832, 7134, 3896], "W": [115 + 2107, 13797, 26550, 58134, 46177]}
    vls = {"B": [360 + 4246, 13379, 12712, 17378, 9315], "H": [191 + 1891, 9637, 13419, 15703, 7186],
           "O": [52 + 605, 2642, 3793, 4981, 2740], "W": [61 + 1279, 8850, 17404, 39492, 31731]}
    z = 1.959963984540054
    for r in RACES:
        n, p = sum(tot[r]), sum(vls[r]) / sum(tot[r])
        rows.append(("ct_vls_pr_byrace", p, r, p - z * math.sqrt(p * (1 - p) / n), p + z * math.sqrt(p * (1 - p) / n)))
    for a in range(5):
        n = sum(tot[r][a] for r in RACES); p = sum(vls[r][a] for r in RACES) / n
        rows.append(("ct_vls_pr_byage", p, f"age{a + 1}", p - z * math.sqrt(p * (1 - p) / n), p + z * math.sqrt(p * (1 - p) / n)))
    # PrEP coverage (calval_targets.r:159-186)
    prep_n, tot_n = [222, 357, 137, 697], [846, 1191, 344, 1645]
    for r, x, n in zip(RACES, prep_n, tot_n):
        lo, hi = _binom_exact_ci(x, n)
        rows.append(("ct_prep", x / n, r, lo, hi))
    # HIV incidence and prevalence (calval_targets.r:189-218)
    rows.append(("ct_hiv_incid_per100pop", 514 / 1000, None, 444 / 1000, 584 / 1000))
    rows.append(("ct_hiv_prev", 0.124, None, 0.109, 0.138))
    # gonorrhoea in STD clinics (calval_targets.r:221-322)
    visits, tests = 139718, [101466, 126972, 123326]
    for s, t in zip(ANAT, tests):
        rows.append(("ct_prop_anatsite_tested", t / visits * 0.732, s, t / visits * 0.648, t / visits * 0.817))
    for s, v, lo, hi in zip(GCLABS, (0.118, 0.075, 0.091), (0.104, 0.057, 0.079), (0.132, 0.093, 0.103)):
        rows.append(("ct_prop_anatsite_pos", v, s, lo, hi))
    return rows


def valtargets() -> dict:
    """calval_targets.r:329-350"""
    return {"vt_age_among_hivdx": [0.095, 0.340, 0.193, 0.191, 0.182, 0.061, 0.265, 0.244, 0.248, 0.183,
                                   0.047, 0.254, 0.220, 0.252, 0.226, 0.021, 0.125, 0.149, 0.280, 0.425],
            "vt_hiv_incid_100k": {"B": 2459, "H": 1140, "W": 234}}


def target_value(target: str, subgroup: Optional[str] = None) -> float:
    for t, v, s, _, _ in caltargets():
        if t == target and s == subgroup:
            return v
    raise KeyError((target, subgroup))


# ---------------------------------------------------------------------------------------------------------------------
# R/calculate_targets.r mirrors.  `epi` is a mapping name -> 1-D array over weeks (dat$epi).
# ---------------------------------------------------------------------------------------------------------------------
def _tail(x, n):
    return np.asarray(x, dtype=float)[-n:]


def _nan0(v):
    v = np.asarray(v, dtype=float)
    return np.where(np.isnan(v) | np.isinf(v), 0.0, v)


def target_prob_hivdx(epi, tailn: int, strat: str) -> Dict[str, float]:
    """R/calculate_targets.r:14-65"""
    assert strat in ("age", "race", "both")
    out = {}
    if strat == "age":
        for a in range(1, 6):
            out[f"prob.hivdx.age{a}"] = np.mean(_tail(epi[f"i.num.dx.{a}"], tailn) / _tail(epi[f"i.num.{a}"], tailn))
    elif strat == "race":
        for r in RACES:
            out[f"prob.hivdx.{r}"] = np.mean(_tail(epi[f"i.num.dx.{r}"], tailn) / _tail(epi[f"i.num.{r}"], tailn))
    else:
        for r in RACES:
            for a in range(1, 6):
                out[f"prob.hivdx.{r}.age{a}"] = np.mean(_tail(epi[f"i.num.dx.{r}.age{a}"], tailn) / _tail(epi[f"i.num.{r}.age{a}"], tailn))
    return {k: float(_nan0(v)) for k, v in out.items()}


AGEDX_CELLS = {"B": (1, 2), "H": (1, 5), "O": (1,), "W": (1, 4, 5)}        # R/calculate_targets.r:72-77


def target_prob_agedx_byrace(epi, tailn: int) -> Dict[str, float]:
    """R/calculate_targets.r:67-103"""
    dl = AGEDX_CELLS
    return {f"prob.agedx.hiv.{r}.age{a}": float(np.mean(_tail(epi[f"i.num.dx.{r}.age{a}"], tailn) / _tail(epi[f"i.num.dx.{r}"], tailn)))
            for r, ages in dl.items() for a in ages}


def target_prob_vls(epi, tailn: int, strat: str) -> Dict[str, float]:
    """R/calculate_targets.r:105-153"""
    assert strat in ("age", "race", "both")
    if strat == "age":
        out = {f"prob.vsupp.age{a}": np.mean(_tail(epi[f"cc.vsupp.age{a}"], tailn)) for a in range(1, 6)}
    elif strat == "race":
        out = {f"prob.vsupp.{r}": np.mean(_tail(epi[f"cc.vsupp.{r}"], tailn)) for r in RACES}
    else:
        out = {f"cc.vsupp.{r}.age{a}": np.mean(_tail(epi[f"cc.vsupp.{r}.age{a}"], tailn)) for r in RACES for a in range(1, 6)}
    return {k: float(_nan0(v)) for k, v in out.items()}


def target_prob_prep_byrace(epi, tailn: int) -> Dict[str, float]:
    """R/calculate_targets.r:155-171"""
    return {f"prep.cov.{r}": float(_nan0(np.mean(_tail(epi[f"prepCurr.{r}"], tailn) / _tail(epi[f"prepElig.{r}"], tailn)))) for r in RACES}


def target_prop_anatsites_tested(epi, tailn: int) -> Dict[str, float]:
    """R/calculate_targets.r:173-181: every series whose name starts with 'prop'."""
    return {k: float(np.mean(_tail(v, tailn))) for k, v in epi.items() if k.startswith("prop")}


def target_prob_gcpos_tested_anatsites(epi, tailn: int) -> Dict[str, float]:
    """R/calculate_targets.r:183-190: every series whose name starts with 'prob'."""
    return {k: float(np.mean(_tail(v, tailn))) for k, v in epi.items() if k.startswith("prob")}


# ---------------------------------------------------------------------------------------------------------------------
# inst/cal/sim0.0_setup.r pipeline
# ---------------------------------------------------------------------------------------------------------------------
def derived_measures(epi: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """sim0.0_setup.r:205-289"""
    e = dict(epi)
    with np.errstate(divide="ignore", invalid="ignore"):
        for lab in ("",) + tuple("." + r for r in RACES):
            e["i.prev.dx.inf" + lab] = e["i.prev.dx" + lab] / e["i.prev" + lab]
            e["prepCov" + lab] = e["prepCurr" + lab] / e["prepElig" + lab]
        for a in range(1, 6):
            e[f"i.prev.dx.inf.age{a}"] = e[f"i.prev.dx.age.{a}"] / e[f"i.prev.age.{a}"]
        for r in RACES[:3]:
            e[f"i.prev.dx.inf.{r}.rel.ref.W"] = e[f"i.prev.dx.inf.{r}"] / e["i.prev.dx.inf.W"]
            e[f"cc.vsupp.{r}.rel.ref.W"] = e[f"cc.vsupp.{r}"] / e["cc.vsupp.W"]
            e[f"prepCov.{r}.rel.ref.W"] = e[f"prepCov.{r}"] / e["prepCov.W"]
        for a in range(1, 5):
            e[f"i.prev.dx.inf.age{a}.rel.ref.age5"] = e[f"i.prev.dx.inf.age{a}"] / e["i.prev.dx.inf.age5"]
            e[f"cc.vsupp.age{a}.rel.ref.age5"] = e[f"cc.vsupp.age{a}"] / e["cc.vsupp.age5"]
        e["ir100.pop"] = e["incid"] / e["num"] * 5200
        for r in RACES:
            e[f"ir100.pop.{r}"] = e[f"incid.{r}"] / e[f"num.{r}"] * 5200
            e[f"prop.{r}"] = e[f"num.{r}"] / e["num"]
        for a in range(1, 6):
            e[f"prop.age.{a}"] = e[f"num.age.{a}"] / e["num"]
    return e


def tail_means(epi: Dict[str, np.ndarray], tailn: int = 52) -> Dict[str, float]:
    """sim0.0_setup.r:299-308: NA -> 0, then the mean over the last `tailn` weeks."""
    return {k: float(np.mean(_nan0(_tail(v, tailn)))) for k, v in epi.items()}


def select_by_popsize(num_mean: np.ndarray, pop_n: float = 20000.0, tol: float = 0.05) -> np.ndarray:
    """sim0.0_setup.r:315-318"""
    return np.abs(pop_n - np.asarray(num_mean)) / pop_n <= tol


def get_absdiff(target: str, output: np.ndarray, subgroup: Optional[str] = None) -> Dict[str, np.ndarray]:
    """sim0.0_setup.r:333-379 for a vector of per-simulation outputs."""
    row = [r for r in caltargets() if r[0] == target and r[2] == subgroup]
    if not row:
        raise KeyError((target, subgroup))
    _, val, _, lo, hi = row[0]
    out = np.asarray(output, dtype=float)
    diff = out - val
    ad = np.abs(diff)
    sd = np.std(out, ddof=1) if out.size > 1 else np.nan
    pct = np.round(ad / val * 100, 2)
    t5, t10 = (0.5, 1.0) if target == "ct_hiv_incid_per100pop" else (0.05, 0.10)
    incl = np.full(out.shape, np.nan) if lo is None else ((out >= lo) & (out <= hi)).astype(float)
    return {"diff": diff, "abs_diff": ad, "abs_std_diff": ad / sd, "pct_diff": pct, "output_in_targcl": incl,
            "output_within_10pct": (pct <= 10).astype(float), "output_within_5pts": (ad <= t5).astype(float),
            "output_within_10pts": (ad <= t10).astype(float), "target_val": val, "ll95": lo, "ul95": hi}


def device_target_names() -> list:
    """Names of the columns of xgc_targets() (enum xgc_target_index), in order."""
    n = ["num", "i.prev", "ir100.pop"]
    n += [f"i.prev.dx.inf.{r}" for r in RACES] + [f"i.prev.dx.inf.age{a}" for a in range(1, 6)]
    n += [f"cc.vsupp.{r}" for r in RACES] + [f"cc.vsupp.age{a}" for a in range(1, 6)]
    n += [f"prepCov.{r}" for r in RACES]
    n += [f"prop.{s}.tested" for s in ANAT] + [f"prob.{g}.tested" for g in GCLABS]
    n += ["prev.gc", "prev.rgc", "prev.ugc", "prev.pgc", "ir100.gc", "ir100.rgc", "ir100.ugc", "ir100.pgc"]
    n += [f"prob.agedx.hiv.{r}.age{a}" for r, ages in AGEDX_CELLS.items() for a in ages]       # XGC_TG_AGEDX
    return n


# which calibration target each device column is compared with (inst/cal/sim0.1_fmt_targetdat.r:92-168)
TARGET_OF_COLUMN = {"i.prev": ("ct_hiv_prev", None), "ir100.pop": ("ct_hiv_incid_per100pop", None),
                    **{f"i.prev.dx.inf.{r}": ("ct_hivdx_pr_byrace", r) for r in RACES},
                    **{f"i.prev.dx.inf.age{a}": ("ct_hivdx_pr_byage", f"age{a}") for a in range(1, 6)},
                    **{f"cc.vsupp.{r}": ("ct_vls_pr_byrace", r) for r in RACES},
                    **{f"cc.vsupp.age{a}": ("ct_vls_pr_byage", f"age{a}") for a in range(1, 6)},
                    **{f"prepCov.{r}": ("ct_prep", r) for r in RACES},
                    **{f"prop.{s}.tested": ("ct_prop_anatsite_tested", s) for s in ANAT},
                    **{f"prob.{g}.tested": ("ct_prop_anatsite_pos", g) for g in GCLABS}}
