"""BASELINE configs[2] end to end on one GPU: burn in R replicates, checkpoint, fork every replicate into the reference's
five screening arms (inst/analysis02_screening/01_setup.r:116-159), continue every arm, and summarise infections averted
(99_analyze.r:146-208: NIA, PIA, NTIA matched by simid).  Writes gpurun_out/screening_grid_<tag>.json.
Usage: python tools/run_screening_grid.py [replicates] [burnin_weeks] [scenario_weeks] [mode] [tag]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xgc_msm_b200 import params as P, scenarios as S, workload as W
from xgc_msm_b200.abi import Handle

R = int(sys.argv[1]) if len(sys.argv) > 1 else 500
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 520
cont = int(sys.argv[3]) if len(sys.argv) > 3 else 260
spec = W.SweepSpec(n_agents=10000, seed=1971, lhs=False)          # calibrated-model stand-in: prior midpoints for every replicate
t0 = time.perf_counter()
h = Handle(R, spec.n_cap, spec.e_cap, burn + cont + 8, seed=1971, device=0, compact_every=spec.weeks_between_compactions)
W.load_sweep(h, spec)
mode = sys.argv[4] if len(sys.argv) > 4 else "fast"
tag = sys.argv[5] if len(sys.argv) > 5 else "r02"
h.set_mode(mode)
h.run(2, 1 + burn); h.sync()
t1 = time.perf_counter()
res = S.run_screening_grid(h, 1 + burn, cont, P.control_msm(network_mode="surrogate"), seed=1971)
t2 = time.perf_counter()
h.close()
med = lambda x: float(np.nanmedian(np.where(np.isfinite(x), x, np.nan)))
iqr = lambda x: [float(np.nanpercentile(np.where(np.isfinite(x), x, np.nan), q)) for q in (25, 75)]
out = {"mode": mode, "replicates": R, "burnin_weeks": burn, "scenario_weeks": cont, "burnin_seconds": t1 - t0, "five_arms_seconds": t2 - t1, "arms": {}}
for arm, s in res.items():
    out["arms"][arm] = {"sum_incid.gc_median": med(s["sum_incid.gc"]), "sum_num.gc.test_median": med(s["sum_num.gc.test"]),
                        "pia_gc_median": med(s["pia_gc"]), "pia_gc_iqr": iqr(s["pia_gc"]), "nia_gc_median": med(s["nia_gc"]),
                        "ntia_gc_median": med(s["ntia_gc"]),
                        **{f"pia_{site}_median": med(s[f"pia_{site}"]) for site in ("rgc", "ugc", "pgc")}}
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"screening_grid_{tag}.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "arms"}))
for arm, v in out["arms"].items():
    print(arm, {k: round(x, 3) if isinstance(x, float) else x for k, x in v.items() if k in ("sum_incid.gc_median", "sum_num.gc.test_median", "pia_gc_median", "ntia_gc_median")})
