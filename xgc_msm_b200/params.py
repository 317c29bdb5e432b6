"""param_msm / control_msm / init_msm mirrors and the calibration-input surface of the reference.

Reference: /root/reference/burnin/cal/sim1_02_lhs.r:58-230 (how the 62 environment-variable inputs become
param_msm()/control_msm() arguments) and burnin/cal/sim1_01_lhs_params.r:22-137 (uniform priors, Latin hypercube).
The parameter vector layout is include/xgc/params.def, parsed here so Python, C and CUDA share one list.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field
from typing import Dict, Iterable, Optional

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARAMS_DEF = os.path.join(_ROOT, "include", "xgc", "params.def")


def _parse_def():
    out, off = {}, 0
    pat = re.compile(r"^XGC_PARAM\(\s*(\w+)\s*,\s*(\d+)\s*,\s*([^)]*?(?:\([^)]*\))?[^)]*)\)\s*(?:/\*.*)?$")
    for line in open(PARAMS_DEF):
        line = line.strip()
        if not line.startswith("XGC_PARAM("):
            continue
        m = pat.match(line)
        if not m:
            raise ValueError(f"cannot parse params.def line: {line}")
        name, cnt, dflt = m.group(1), int(m.group(2)), float(eval(m.group(3), {"__builtins__": {}}))
        out[name] = (off, cnt, dflt)
        off += cnt
    return out, off


PARAM_INDEX, NPARAM = _parse_def()

# vector defaults that are not a single repeated value ([BK] EpiModelHIV defaults)
_VECTOR_DEFAULTS = {
    "prep_adhr_dist": (0.089, 0.127, 0.784),
    "prep_adhr_hr": (0.69, 0.19, 0.01),
    "hiv_test_rate": (0.01325, 0.0125, 0.0130, 0.0138),
    "rgc_tx_recov_pr": (1 - 0.06, 0.5, 1.0),
    "ugc_tx_recov_pr": (1 - 0.105, 0.5, 1.0),
    "pgc_tx_recov_pr": (1 - 0.13, 0.5, 1.0),
}

# R-style aliases of the reference call that map onto slices of a vector parameter (sim1_02_lhs.r:77-86)
_SITE_ALIASES = {"rgc_ntx_int": ("gc_ntx_int", 0), "ugc_ntx_int": ("gc_ntx_int", 1), "pgc_ntx_int": ("gc_ntx_int", 2),
                 "rgc_sympt_prob": ("gc_sympt_prob", 0), "ugc_sympt_prob": ("gc_sympt_prob", 1), "pgc_sympt_prob": ("gc_sympt_prob", 2)}


def default_vector() -> np.ndarray:
    p = np.zeros(NPARAM)
    for name, (off, cnt, d) in PARAM_INDEX.items():
        p[off:off + cnt] = _VECTOR_DEFAULTS.get(name, d)
    return p


def param_msm(base: Optional[np.ndarray] = None, **kw) -> np.ndarray:
    """Build one replicate's parameter vector.  Keyword names are the reference's param_msm() argument names with
    '.' written as '_' (e.g. u2rgc_tprob=0.2, gc_asympt_prob_test=(.6,.58,.63), rgc_ntx_int=12)."""
    p = default_vector() if base is None else np.array(base, dtype=np.float64)
    for k, v in kw.items():
        k = k.replace(".", "_")
        if k in _SITE_ALIASES:
            name, j = _SITE_ALIASES[k]
            p[PARAM_INDEX[name][0] + j] = float(v)
            continue
        if k not in PARAM_INDEX:
            raise KeyError(f"param_msm: unknown parameter '{k}'")
        off, cnt, _ = PARAM_INDEX[k]
        v = np.atleast_1d(np.asarray(v, dtype=np.float64))
        if v.size == 1:
            p[off:off + cnt] = v[0]
        elif v.size == cnt:
            p[off:off + cnt] = v
        else:
            raise ValueError(f"param_msm: '{k}' expects {cnt} values, got {v.size}")
    return p


def get(p: np.ndarray, name: str) -> np.ndarray:
    off, cnt, _ = PARAM_INDEX[name.replace(".", "_")]
    return p[..., off:off + cnt]


# ---------------------------------------------------------------------------------------------------------------------
# control_msm (sim1_02_lhs.r:195-230)
# ---------------------------------------------------------------------------------------------------------------------
CONTROL_INDEX = {"transRoute_Kissing": 0, "transRoute_Rimming": 1, "gcUntreatedRecovDist": 2, "stiScreeningProtocol": 3,
                 "cdcExposureSite_Kissing": 4, "network_mode": 5}
NCONTROL = 8
_RECOV = {"geom": 0, "pois": 1}
_SCREEN = {"base": 0, "symptomatic": 1, "cdc": 2, "universal": 3}
MODULE_ORDER = ["aging", "departure", "arrival", "hivtest", "hivtx", "hivprogress", "hivvl", "resim_nets", "acts", "condoms",
                "position", "prep", "hivtrans", "stirecov", "stitx", "stitrans", "prev"]
MODULE_BIT = {m: 1 << i for i, m in enumerate(MODULE_ORDER)}


@dataclass
class Control:
    nsims: int = 1
    nsteps: int = 52
    ncores: int = 1
    start: int = 1
    flags: np.ndarray = field(default_factory=lambda: np.array([1, 1, 0, 0, 0, 0, 0, 0], dtype=np.int32))
    modules: Dict[str, object] = field(default_factory=dict)     # name -> FUN(dat, at) -> dat
    extra: Dict[str, object] = field(default_factory=dict)       # skip.check, tergmLite, debug_stitx, save.network, verbose


def control_msm(nsims=1, nsteps=52, ncores=1, start=1, transRoute_Kissing=True, transRoute_Rimming=True,
                gcUntreatedRecovDist="geom", stiScreeningProtocol="base", cdcExposureSite_Kissing=False,
                network_mode="host", **kw) -> Control:
    """Mirror of control_msm(...).  `<name>_FUN` keywords register module functions (defaults: xgc_msm_b200.modules)."""
    if gcUntreatedRecovDist not in _RECOV:
        raise ValueError("gcUntreatedRecovDist must be 'geom' or 'pois'")
    if stiScreeningProtocol not in _SCREEN:
        raise ValueError("stiScreeningProtocol must be one of base, symptomatic, cdc, universal")
    c = Control(nsims=int(nsims), nsteps=int(nsteps), ncores=int(ncores), start=int(start))
    c.flags = np.array([int(bool(transRoute_Kissing)), int(bool(transRoute_Rimming)), _RECOV[gcUntreatedRecovDist],
                        _SCREEN[stiScreeningProtocol], int(bool(cdcExposureSite_Kissing)),
                        {"host": 0, "surrogate": 1}[network_mode], 0, 0], dtype=np.int32)
    for k, v in kw.items():
        if k.endswith("_FUN") or k.endswith(".FUN"):
            c.modules[k[:-4].replace(".", "_")] = v
        else:
            c.extra[k] = v
    return c


def init_msm(prev_ugc=0.05, prev_rgc=0.05, prev_pgc=0.05) -> dict:
    """sim1_02_lhs.r:180-184"""
    return {"prev.rgc": float(prev_rgc), "prev.ugc": float(prev_ugc), "prev.pgc": float(prev_pgc)}


# ---------------------------------------------------------------------------------------------------------------------
# the 62 calibration inputs: priors (sim1_01_lhs_params.r:22-102) and their mapping to param_msm (sim1_02_lhs.r:58-178)
# ---------------------------------------------------------------------------------------------------------------------
def _halt(p):
    return 1 - (1 - p) ** (1 / 52)


PRIORS = [
    ("U2RGC_PROB", 0, 0.35), ("U2PGC_PROB", 0, 0.45), ("R2UGC_PROB", 0, 0.55), ("P2UGC_PROB", 0, 0.25),
    ("R2PGC_PROB", 0, 0.45), ("P2RGC_PROB", 0, 0.20), ("P2PGC_PROB", 0, 0.10),
    ("RECT_GC_DURAT_NOTX", 2, 22), ("URETH_GC_DURAT_NOTX", 1, 24), ("PHAR_GC_DURAT_NOTX", 3, 20),
    ("RECT_GC_RX_INFPR_WK1", 0.01, 0.11), ("URETH_GC_RX_INFPR_WK1", 0.01, 0.20), ("PHAR_GC_RX_INFPR_WK1", 0.06, 0.20),
    ("RECT_GC_SYMPT_PROB", 0.06, 0.46), ("URETH_GC_SYMPT_PROB", 0.46, 0.99), ("PHAR_GC_SYMPT_PROB", 0.00, 0.15),
    ("STITEST_PROB_UGC_SYMPT", 0.5, 1), ("STITEST_RGC_RR_SYMPT", 0.01, 0.99), ("STITEST_PGC_RR_SYMPT", 0.01, 0.99),
    ("RECT_ASYMP_STITEST_PROB", 0.400, 0.800), ("URETH_ASYMP_STITEST_PROB", 0.497, 0.657), ("PHAR_ASYMP_STITEST_PROB", 0.522, 0.749),
    ("RECT_SYMP_STITEST_PROB", 0.5, 1), ("URETH_SYMP_STITEST_PROB", 0.5, 1), ("PHAR_SYMP_STITEST_PROB", 0.5, 1),
    ("HIV_LATE_TESTER_PROB_BLACK", 0.160, 0.419), ("HIV_LATE_TESTER_PROB_HISP", 0.202, 0.391),
    ("HIV_LATE_TESTER_PROB_OTHER", 0.207, 0.369), ("HIV_LATE_TESTER_PROB_WHITE", 0.222, 0.377),
    ("HIV_RX_INIT_PROB_BLACK", 0.001, 0.99), ("HIV_RX_INIT_PROB_HISP", 0.001, 0.99), ("HIV_RX_INIT_PROB_OTHER", 0.001, 0.99),
    ("HIV_RX_INIT_PROB_WHITE", 0.001, 0.99),
    ("HIV_RX_HALT_PROB_BLACK", _halt(0.464), _halt(0.464)), ("HIV_RX_HALT_PROB_HISP", _halt(0.416), _halt(0.416)),
    ("HIV_RX_HALT_PROB_OTHER", _halt(0.366), _halt(0.366)), ("HIV_RX_HALT_PROB_WHITE", _halt(0.406), _halt(0.406)),
    ("HIV_RX_REINIT_PROB_BLACK", 0.001, 0.99), ("HIV_RX_REINIT_PROB_HISP", 0.001, 0.99), ("HIV_RX_REINIT_PROB_OTHER", 0.001, 0.99),
    ("HIV_RX_REINIT_PROB_WHITE", 0.001, 0.99),
    ("SCALAR_AI_ACT_RATE", 1, 1), ("SCALAR_OI_ACT_RATE", 1, 1),
    ("KISS_RATE_MAIN", 0, 10), ("KISS_RATE_CASUAL", 0, 10), ("KISS_PROB_ONETIME", 0.5, 1),
    ("RIM_RATE_MAIN", 0, 10), ("RIM_RATE_CASUAL", 0, 10), ("RIM_PROB_ONETIME", 0, 1),
    ("SCALAR_HIV_TRANS_PROB_BLACK", 0.75, 3), ("SCALAR_HIV_TRANS_PROB_HISP", 0.75, 3), ("SCALAR_HIV_TRANS_PROB_OTHER", 0.75, 3),
    ("SCALAR_HIV_TRANS_PROB_WHITE", 1, 1),
    ("CONDOM_EFF_HIV", 0.6, 1), ("CONDOM_EFF_GC", 0.7, 1),
    ("PREP_DISCONT_RATE_BLACK", 0, 0.05), ("PREP_DISCONT_RATE_HISP", 0, 0.05), ("PREP_DISCONT_RATE_OTHER", 0, 0.05),
    ("PREP_DISCONT_RATE_WHITE", 0, 0.05),
    ("ACT_STOPPER_PROB", 0.2, 1), ("HIV_TRANS_RR_RGC", 1.2, 6.4), ("HIV_TRANS_RR_UGC", 1, 3),
]
assert len(PRIORS) == 62        # length(priors), sim1_01_lhs_params.r:119

_R4 = ("BLACK", "HISP", "OTHER", "WHITE")
_S3 = ("RECT", "URETH", "PHAR")


# Weekly arrivals per initial-population member that hold the SYNTHETIC demography (synth.make_tables: ages 18-65, the
# age-specific mortality table, AIDS deaths) at its initial size: 1 / (52 x ~43.5 years mean residence) minus the reference's own
# tweak.  The reference accepts a burn-in whose population stays within +-5 % of its start (inst/cal/sim0.0_setup.r:315-318);
# tools/run_burnin.py records the 3120-week drift (profiles/burnin_3120_r02.json).
A_RATE_BASE = 0.000378


def params_from_env(env: Dict[str, float], base: Optional[np.ndarray] = None, a_rate_base: float = A_RATE_BASE,
                    arrive_rate_add_per20k: float = 1.285, n_init: float = 10000.0) -> np.ndarray:
    """The param_msm() call of sim1_02_lhs.r:58-178 with `env` standing for the environment variables."""
    e = lambda k: float(env[k])
    return param_msm(
        base,
        arrival_age=18, a_rate=a_rate_base + arrive_rate_add_per20k / 20000.0, n_init=n_init,
        u2rgc_tprob=e("U2RGC_PROB"), u2pgc_tprob=e("U2PGC_PROB"), r2ugc_tprob=e("R2UGC_PROB"), p2ugc_tprob=e("P2UGC_PROB"),
        r2pgc_tprob=e("R2PGC_PROB"), p2rgc_tprob=e("P2RGC_PROB"), p2pgc_tprob=e("P2PGC_PROB"),
        gc_ntx_int=[e(f"{s}_GC_DURAT_NOTX") for s in _S3],
        rgc_tx_recov_pr=[1 - e("RECT_GC_RX_INFPR_WK1"), 0.5, 1], ugc_tx_recov_pr=[1 - e("URETH_GC_RX_INFPR_WK1"), 0.5, 1],
        pgc_tx_recov_pr=[1 - e("PHAR_GC_RX_INFPR_WK1"), 0.5, 1],
        gc_sympt_prob=[e(f"{s}_GC_SYMPT_PROB") for s in _S3],
        ugc_sympt_seek_test_prob=e("STITEST_PROB_UGC_SYMPT"), rgc_sympt_seek_test_rr=e("STITEST_RGC_RR_SYMPT"),
        pgc_sympt_seek_test_rr=e("STITEST_PGC_RR_SYMPT"),
        gc_asympt_prob_test=[e(f"{s}_ASYMP_STITEST_PROB") for s in _S3], gc_sympt_prob_test=[e(f"{s}_SYMP_STITEST_PROB") for s in _S3],
        hiv_test_late_prob=[e(f"HIV_LATE_TESTER_PROB_{r}") for r in _R4],
        tt_part_supp=0, tt_full_supp=1, tt_dur_supp=0,
        tx_init_prob=[e(f"HIV_RX_INIT_PROB_{r}") for r in _R4], tx_halt_part_prob=[e(f"HIV_RX_HALT_PROB_{r}") for r in _R4],
        tx_halt_full_rr=1, tx_halt_dur_rr=1, tx_reinit_part_prob=[e(f"HIV_RX_REINIT_PROB_{r}") for r in _R4],
        tx_reinit_full_rr=1, tx_reinit_dur_rr=1,
        ai_acts_scale_mc=e("SCALAR_AI_ACT_RATE"), oi_acts_scale_mc=e("SCALAR_OI_ACT_RATE"),
        kiss_rate_main=e("KISS_RATE_MAIN"), kiss_rate_casl=e("KISS_RATE_CASUAL"), kiss_prob_oo=e("KISS_PROB_ONETIME"),
        rim_rate_main=e("RIM_RATE_MAIN"), rim_rate_casl=e("RIM_RATE_CASUAL"), rim_prob_oo=e("RIM_PROB_ONETIME"),
        trans_scale=[e(f"SCALAR_HIV_TRANS_PROB_{r}") for r in _R4], cdc_sti_int=12, cdc_sti_hr_int=6,
        cond_eff=e("CONDOM_EFF_HIV"), cond_fail=0, sti_cond_fail=0, sti_cond_eff=e("CONDOM_EFF_GC"),
        prep_discont_rate=[e(f"PREP_DISCONT_RATE_{r}") for r in _R4],
        act_stopper_prob=e("ACT_STOPPER_PROB"), hiv_rgc_rr=e("HIV_TRANS_RR_RGC"), hiv_ugc_rr=e("HIV_TRANS_RR_UGC"))


def prior_midpoint_env() -> Dict[str, float]:
    return {k: 0.5 * (lo + hi) for k, lo, hi in PRIORS}


def random_lhs(n: int, k: int, rng: np.random.Generator) -> np.ndarray:
    """Latin hypercube in [0,1)^k (lhs::randomLHS semantics: one point per stratum per column, random pairing)."""
    u = np.empty((n, k))
    for j in range(k):
        u[:, j] = (rng.permutation(n) + rng.random(n)) / n
    return u


def lhs_envs(n: int, seed: int = 1971) -> list:
    """sim1_01_lhs_params.r:118-137: n parameter sets, draw = (hi - lo) * u + lo."""
    u = random_lhs(n, len(PRIORS), np.random.default_rng(seed))
    return [{name: (hi - lo) * u[i, j] + lo for j, (name, lo, hi) in enumerate(PRIORS)} for i in range(n)]
