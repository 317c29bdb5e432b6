"""Calibration loop around the device path.

One calibration round of the reference is ~5000 SLURM array tasks plus three R scripts; here it is one device sweep:

  priors --LHS--> [n x 62] inputs --params_from_env--> R replicates on the device --xgc_run--> xgc_targets [n x NTARGET]
         --population-size filter--> out_vs_targ (get_absdiff per target) --> selections --> narrowed priors

Reference (behaviour only; none of its code is reused):
  burnin/cal/sim1_01_lhs_params.r:20-137   priors, `randomLHS(5000, 62)`, draw = (hi - lo) * u + lo
  inst/cal/sim0.0_setup.r:299-318          NA -> 0, tail-52 means, |N - 20000| / 20000 <= 0.05 filter
  inst/cal/sim0.0_setup.r:333-379          get_absdiff: abs / standardised / percent difference, within-5/10-points flags
  inst/cal/sim3_01_analyze.r:120-160       per-target "within 5 points" id sets and their intersections
  inst/cal/sim3_01_analyze.r:169-192       `mase` = sum of abs_std_diff over the variables matching `varpats`; best
                                           5 % / 1 % / 0.25 % quantiles and the 10 smallest
  inst/cal/sim0.0_setup.r:513-528          input_quantiles: 25 % / 75 % quantiles of the selected inputs = next priors;
                                           race-specific inputs use the selection made on that race's targets

Everything here is host-side numpy over the [n x K] targets matrix the device returns; the simulation itself never
leaves the CUDA path (no oracle, no CPU fallback).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import params as P
from . import targets as T

Prior = Tuple[str, float, float]

# sim3_01_analyze.r:169-176
MASE_VARPATS = "|".join([r"i\.prev$", r"ir100\.pop", r"cc\.vsupp\.[A-Z]{1}$", r"dx\.inf\.[A-Z]{1}$",
                         r"i\.prev\.dx\.inf\.age[1-5]{1}$", r"prob\.[A-Za-z]{3}\.tested"])
_RACE_SLUG = {"BLACK": "B", "HISP": "H", "OTHER": "O", "WHITE": "W"}


def draw_lhs(priors: Sequence[Prior], n: int, seed: int = 1971) -> np.ndarray:
    """[n, len(priors)] inputs: one Latin hypercube over the prior box (sim1_01_lhs_params.r:118-137)."""
    u = P.random_lhs(n, len(priors), np.random.default_rng(seed))
    lo = np.array([p[1] for p in priors]); hi = np.array([p[2] for p in priors])
    return (hi - lo) * u + lo


def inputs_to_envs(priors: Sequence[Prior], inputs: np.ndarray) -> List[Dict[str, float]]:
    return [{p[0]: float(v) for p, v in zip(priors, row)} for row in inputs]


def out_vs_targ(tmat: np.ndarray, names: Optional[Sequence[str]] = None, keep: Optional[np.ndarray] = None) -> Dict[str, Dict]:
    """The reference's `out_vs_targ` table for the device's targets matrix: for every column that has a calibration
    target, get_absdiff() over the kept simulations (standard deviation taken over those, like the data.table `by`)."""
    names = list(names or T.device_target_names())
    tmat = np.asarray(tmat, dtype=float)
    keep = np.ones(tmat.shape[0], bool) if keep is None else np.asarray(keep, bool)
    out = {}
    for j, nm in enumerate(names):
        if nm not in T.TARGET_OF_COLUMN:
            continue
        target, sub = T.TARGET_OF_COLUMN[nm]
        d = T.get_absdiff(target, tmat[keep, j], sub)
        d["output"] = tmat[keep, j]
        out[nm] = d
    return out


def mase(ovt: Dict[str, Dict], varpats: str = MASE_VARPATS) -> np.ndarray:
    """Per-simulation sum of abs_std_diff over the variables matching `varpats` (sim3_01_analyze.r:178-182)."""
    rx = re.compile(varpats)
    cols = [v["abs_std_diff"] for k, v in ovt.items() if rx.search(k)]
    if not cols:
        raise ValueError("no variable matches varpats")
    return np.nansum(np.stack(cols), axis=0)


def select_within(ovt: Dict[str, Dict], variables: Sequence[str], flag: str = "output_within_5pts") -> np.ndarray:
    """Intersection over `variables` of the simulations whose output carries `flag` (sim3_01_analyze.r:132-160)."""
    sel = None
    for v in variables:
        m = ovt[v][flag] == 1
        sel = m if sel is None else (sel & m)
    return sel


def select_mase_quantile(m: np.ndarray, q: float) -> np.ndarray:
    """mase <= quantile(mase, q) (sim3_01_analyze.r:188-192; R's default type-7 quantile == numpy's default)."""
    return m <= np.quantile(m, q)


def best_fits(m: np.ndarray, n: int = 10) -> np.ndarray:
    """Indices of the n smallest mase values, ascending (sim3_01_analyze.r:194)."""
    return np.argsort(m, kind="stable")[:n]


def input_quantiles(priors: Sequence[Prior], inputs: np.ndarray, selection, ql: float = 0.25, qu: float = 0.75,
                    min_selected: int = 4) -> List[Prior]:
    """Next round's priors (sim0.0_setup.r:513-528).  `selection` is a boolean mask over the rows of `inputs`, or a dict
    {"B"|"H"|"O"|"W"|"*": mask}: an input whose name ends in a race slug takes the mask of that race, others take "*".
    Inputs with a degenerate prior (lo == hi) or fewer than `min_selected` selected rows keep their prior."""
    out = []
    for j, (name, lo, hi) in enumerate(priors):
        if isinstance(selection, dict):
            slug = _RACE_SLUG.get(name.rsplit("_", 1)[-1])
            mask = selection.get(slug, selection.get("*"))
        else:
            mask = selection
        if lo == hi or mask is None or int(np.sum(mask)) < min_selected:
            out.append((name, lo, hi))
            continue
        v = inputs[np.asarray(mask, bool), j]
        out.append((name, float(np.quantile(v, ql)), float(np.quantile(v, qu))))
    return out


@dataclass
class RoundResult:
    priors: List[Prior]
    inputs: np.ndarray                    # [n, k] LHS draws
    targets: np.ndarray                   # [n, NTARGET] tail-window means from xgc_targets
    names: List[str]
    keep: np.ndarray                      # population-size filter
    ovt: Dict[str, Dict] = field(repr=False, default_factory=dict)
    mase: Optional[np.ndarray] = None     # over the kept simulations
    status: int = 0                       # OR of the replicates' status words (capacity overflow flags)

    def kept_index(self) -> np.ndarray:
        return np.flatnonzero(self.keep)


def run_round(priors: Sequence[Prior], n_sims: int, weeks: int, tailn: int = 52, n_agents: int = 10000, seed: int = 1971,
              device: int = 0, pop_tol: float = 0.05, replicate_offset: int = 0, n_distinct: int = 4, mode: str = "fast") -> RoundResult:
    """One calibration round on one GPU: n_sims replicates, each with its own LHS parameter vector, `weeks` weeks on the
    device, tail-`tailn` targets, then the reference's filters.  The start populations are the synthetic ones of
    workload.py (the fitted netest is an LFS pointer).  mode "fast": the replicate-resident native-draw kernels (strict
    kernels when a replicate does not fit one SM's shared memory); "strict": the bit-exact kernels."""
    from . import workload as W
    from .abi import Handle
    spec = W.SweepSpec(n_agents=n_agents, seed=seed, n_distinct=n_distinct)
    inputs = draw_lhs(priors, n_sims, seed)
    envs = inputs_to_envs(priors, inputs)
    h = Handle(n_sims, spec.n_cap, spec.e_cap, weeks + 4, seed=seed, device=device, replicate_offset=replicate_offset,
               compact_every=spec.weeks_between_compactions)
    try:
        W.load_sweep(h, spec, rid0=replicate_offset)
        h.set_params(np.stack([P.params_from_env(e, n_init=n_agents) for e in envs]))
        if mode == "fast":
            try:
                h.set_mode("fast")
            except Exception:
                pass
        h.run(2, 1 + weeks)
        tm = h.targets(1 + weeks, min(tailn, weeks))
        status = 0
        for r in range(0, n_sims, max(n_sims // 64, 1)):
            status |= h.status(r)
    finally:
        h.close()
    names = T.device_target_names()
    keep = T.select_by_popsize(tm[:, names.index("num")], float(n_agents), pop_tol)
    res = RoundResult(list(priors), inputs, tm, names, keep, status=status)
    if keep.sum() >= 2:
        res.ovt = out_vs_targ(tm, names, keep)
        res.mase = mase(res.ovt)
    return res


def narrow(res: RoundResult, q: float = 0.05, ql: float = 0.25, qu: float = 0.75) -> List[Prior]:
    """Priors for the next round from the best-`q` mase quantile of the kept simulations."""
    if res.mase is None:
        return list(res.priors)
    sel_kept = select_mase_quantile(res.mase, q)
    mask = np.zeros(res.inputs.shape[0], bool)
    mask[res.kept_index()[sel_kept]] = True
    return input_quantiles(res.priors, res.inputs, mask, ql, qu)


def calibrate(priors: Sequence[Prior], rounds: int, n_sims: int, weeks: int, **kw) -> List[RoundResult]:
    """`rounds` rounds of run_round() with the priors narrowed in between (the reference's sim1 -> sim2 -> sim3 sequence)."""
    out = []
    pr = list(priors)
    seed = kw.pop("seed", 1971)
    for k in range(rounds):
        res = run_round(pr, n_sims, weeks, seed=seed + 1000 * k, **kw)
        out.append(res)
        pr = narrow(res)
    return out
