"""Screening-scenario grid (BASELINE configs[2]): checkpoint at the end of burn-in, fork every replicate into the
reference's screening arms, continue, and summarise infections averted.

Reference: inst/analysis02_screening/01_setup.r:116-159 (arms: base, symptomatic, cdc with/without kissing as an
exposure, universal; continuation weeks 3121..3380), 02.00_epi.r:40-76, 292-296 (burn-in output reused as the start
state, `reinit_msm`, independent RNG streams per arm from seeds_scenario.rds), 99_analyze.r:146-208 (tests administered,
NIA = base - arm infections, PIA = NIA / base, NTIA = tests / NIA, matched by simid).
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from . import params as P
from .abi import Handle, K

ARMS = {                      # 01_setup.r:116-159
    "base": ("base", False),
    "symp": ("symptomatic", False),
    "cdc1": ("cdc", False),
    "cdc2": ("cdc", True),
    "univ": ("universal", False),
}
SITES = ("gc", "rgc", "ugc", "pgc")


def fork_arms(burnin: Handle, seed: int = 1971, device: int = 0) -> Dict[str, Handle]:
    """One handle per arm, every replicate a device-to-device copy of the burn-in replicate (xgc_fork); arm k draws from
    its own Philox key (seed + 1 + k), like the reference's separate scenario streams.  Every arm continues with the
    burn-in's OWN tables (netstats / epistats stand-ins) and per-replicate parameters, read back from the burn-in handle."""
    tables, params = burnin.get_tables(), burnin.get_params_all()
    arms = {}
    for k, (name, (proto, kiss)) in enumerate(ARMS.items()):
        h = Handle(burnin.n_replicates, burnin.n_cap, burnin.e_cap, burnin.t_cap, seed=seed + 1 + k, device=device,
                   replicate_offset=burnin.cfg.replicate_offset)
        h.set_tables(tables)
        for r in range(burnin.n_replicates):
            h.fork_from(burnin, r, r)
        h.set_params(params)
        h.set_mode(burnin.get_mode())                    # the arms continue on the kernels the burn-in ran on
        arms[name] = h
    return arms


def run_screening_grid(burnin: Handle, at_burnin_end: int, weeks: int, control_base: P.Control, seed: int = 1971,
                       device: int = 0) -> Dict[str, Dict[str, np.ndarray]]:
    """Continue every arm for `weeks` weeks and return per-arm, per-replicate summaries plus NIA / PIA / NTIA vs base."""
    burnin.sync()
    arms = fork_arms(burnin, seed, device)
    t0, t1 = at_burnin_end + 1, at_burnin_end + weeks
    out: Dict[str, Dict[str, np.ndarray]] = {}
    G = K.NCELL
    for name, h in arms.items():
        proto, kiss = ARMS[name]
        flags = control_base.flags.copy()
        flags[P.CONTROL_INDEX["stiScreeningProtocol"]] = {"base": 0, "symptomatic": 1, "cdc": 2, "universal": 3}[proto]
        flags[P.CONTROL_INDEX["cdcExposureSite_Kissing"]] = int(kiss)
        h.set_control(flags)
        h.run(t0, t1)
        h.sync()
        s = {}
        cnt = np.stack([h.counters(r, t0, t1).astype(np.int64).sum(0) for r in range(h.n_replicates)])      # [R, NCOUNTER]
        inc = [cnt[:, (K.K_INCID_RGC + k) * G:(K.K_INCID_RGC + k + 1) * G].sum(1) for k in range(3)]
        s["sum_incid.rgc"], s["sum_incid.ugc"], s["sum_incid.pgc"] = inc
        s["sum_incid.gc"] = inc[0] + inc[1] + inc[2]
        for k, site in enumerate(SITES[1:]):
            s[f"sum_num.{site}.test"] = cnt[:, K.K_TESTED + k]
        s["sum_num.gc.test"] = cnt[:, K.K_TESTED:K.K_TESTED + 3].sum(1)                 # 99_analyze.r:146
        s["asympt_visits"] = cnt[:, K.K_VISITS] - cnt[:, K.K_VISITS_SYMPT]
        out[name] = s
        h.close()
    base = out["base"]
    with np.errstate(divide="ignore", invalid="ignore"):
        for name, s in out.items():
            for site in SITES:
                nia = base[f"sum_incid.{site}"] - s[f"sum_incid.{site}"]                # 99_analyze.r:165-175
                s[f"nia_{site}"] = nia
                s[f"pia_{site}"] = nia / base[f"sum_incid.{site}"]
                s[f"ntia_{site}"] = s[f"sum_num.{site}.test"] / nia
    return out
