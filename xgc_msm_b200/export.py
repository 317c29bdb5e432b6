"""`sim$epi`-compatible output tables.

The reference post-processes every simulation through `as.data.table(sim$epi)` + a `simid` and an `at` column, keeps the
weeks `at >= FIRST_TIME`, stacks the simulations and keeps `keepvars` (inst/cal/sim_clean.r:55-101); the analysis
scripts (inst/cal/sim0.0_setup.r, inst/analysis02_screening/98_summarize_sims.r) then work on that long table.
epi_frame() builds the same table from the device tallies (same column names — Appendix C of SURVEY.md — one row per
(simid, at)); write_epi() stores it as Parquet or Arrow IPC/Feather, both of which R reads with `arrow::read_parquet` /
`arrow::read_feather` into a data.table-compatible frame (RDS itself needs R, which this image does not have).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from .abi import Handle, names as _abi_names

RACELABS = [".B", ".H", ".O", ".W"]
AGELABS = [f".age{a}" for a in range(1, 6)]
AGEDOTS = [f".age.{a}" for a in range(1, 6)]
_STD = [""] + RACELABS + AGELABS
_DOT = [""] + RACELABS + AGEDOTS

# inst/cal/sim_clean.r:76-96 ("subset variables to reduce filesize"), minus simid / at which epi_frame always adds
KEEPVARS: List[str] = (
    [f"num{s}" for s in _DOT] + [f"i.num{s}" for s in _STD] + [f"i.prev{s}" for s in _DOT] + [f"i.prev.dx{s}" for s in _DOT] +
    [f"incid{s}" for s in _STD] + [f"cc.vsupp{s}" for s in _STD] + [f"prepElig{s}" for s in _STD[:5]] +
    [f"prepCurr{s}" for s in _STD[:5]] + [f"i.num.{s}" for s in ("gc", "rgc", "ugc", "pgc")] + ["ir100"] +
    [f"prop.{s}.tested" for s in ("rect", "ureth", "phar")] + [f"prob.{s}.tested" for s in ("rGC", "uGC", "pGC")])


def epi_frame(h: Handle, at0: int, at1: int, replicates: Optional[Iterable[int]] = None,
              names: Optional[Sequence[str]] = None, first_time: Optional[int] = None) -> Dict[str, np.ndarray]:
    """Long table {column: vector}: `simid` (global replicate id, 1-based like the reference's array task ids), `at`, and
    one float64 column per epi name.  `names` defaults to every series the library knows (xgc_epi_names);
    `first_time` drops the weeks before it (sim_clean.r:63)."""
    names = list(names) if names is not None else _abi_names("epi")
    reps = list(range(h.n_replicates)) if replicates is None else list(replicates)
    if first_time is not None:
        at0 = max(at0, int(first_time))
    T = at1 - at0 + 1
    if T < 1:
        raise ValueError("empty week range")
    cols = {n: np.empty(len(reps) * T) for n in names}
    simid = np.empty(len(reps) * T, dtype=np.int32)
    at = np.empty(len(reps) * T, dtype=np.int32)
    for k, r in enumerate(reps):
        tab = h.epi_table(r, names, at0, at1)
        sl = slice(k * T, (k + 1) * T)
        for j, n in enumerate(names):
            cols[n][sl] = tab[j]
        simid[sl] = h.cfg.replicate_offset + r + 1
        at[sl] = np.arange(at0, at1 + 1)
    return {"simid": simid, "at": at, **cols}


def write_epi(path: str, frame: Dict[str, np.ndarray]) -> str:
    """Parquet (`*.parquet`) or Arrow IPC / Feather v2 (anything else).  Returns the path."""
    import pyarrow as pa
    table = pa.table(frame)
    if path.endswith(".parquet"):
        import pyarrow.parquet as pq
        pq.write_table(table, path)
    else:
        import pyarrow.feather as pf
        pf.write_feather(table, path)
    return path


def read_epi(path: str) -> Dict[str, np.ndarray]:
    import pyarrow.parquet as pq
    import pyarrow.feather as pf
    t = pq.read_table(path) if path.endswith(".parquet") else pf.read_table(path)
    return {n: t.column(n).to_numpy() for n in t.column_names}
