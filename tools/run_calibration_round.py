"""BASELINE configs[1] end to end on one GPU: one calibration round = LHS over the reference prior box -> 1000 replicates
x 10k agents x 520 weeks on the device -> tail-52 targets -> population filter -> out_vs_targ / mase -> narrowed priors.
Writes a summary JSON (gpurun_out/calibration_round_<tag>.json).  Usage: python tools/run_calibration_round.py [n_sims] [weeks] [mode] [tag]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xgc_msm_b200 import calibration as CAL, params as P, targets as T

n_sims = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
weeks = int(sys.argv[2]) if len(sys.argv) > 2 else 520
t0 = time.perf_counter()
mode = sys.argv[3] if len(sys.argv) > 3 else "fast"
tag = sys.argv[4] if len(sys.argv) > 4 else "r02"
res = CAL.run_round(P.PRIORS, n_sims, weeks, tailn=52, n_agents=10000, seed=1971, mode=mode)
wall = time.perf_counter() - t0
names = res.names
out = {"mode": mode, "n_sims": n_sims, "weeks": weeks, "wall_seconds_incl_setup": wall, "status_or": res.status,
       "kept_by_population_filter": int(res.keep.sum()),
       "targets_median": {n: float(np.nanmedian(res.targets[:, j])) for j, n in enumerate(names)},
       "targets_p05_p95": {n: [float(np.nanpercentile(res.targets[:, j], 5)), float(np.nanpercentile(res.targets[:, j], 95))] for j, n in enumerate(names)}}
if res.mase is not None:
    kept = res.kept_index()
    best = CAL.best_fits(res.mase, 10)
    out["mase"] = {"min": float(res.mase.min()), "median": float(np.median(res.mase)), "best10_simid": [int(kept[b]) + 1 for b in best],
                   "best10_mase": [float(res.mase[b]) for b in best]}
    out["within_5pts_share"] = {k: float(np.mean(v["output_within_5pts"])) for k, v in res.ovt.items()}
    out["target_values"] = {k: v["target_val"] for k, v in res.ovt.items()}
    new = CAL.narrow(res, q=0.05)
    out["narrowed_priors"] = {n: [lo, hi] for (n, lo, hi), (_, lo0, hi0) in zip(new, P.PRIORS) if (lo, hi) != (lo0, hi0)}
path = os.path.join(ROOT, "gpurun_out", f"calibration_round_{tag}.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps({k: out[k] for k in ("n_sims", "weeks", "wall_seconds_incl_setup", "status_or", "kept_by_population_filter")}), out.get("mase", {}).get("min"))
