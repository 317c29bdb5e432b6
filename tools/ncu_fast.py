"""One short fast-mode run for ncu: R replicates x N agents, a few weeks in ONE launch of k_week_fast."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgc_msm_b200 import workload
from xgc_msm_b200.abi import Handle
ap = argparse.ArgumentParser()
ap.add_argument("--replicates", type=int, default=148)
ap.add_argument("--agents", type=int, default=10000)
ap.add_argument("--weeks", type=int, default=12)
ap.add_argument("--mode", default="fast")
a = ap.parse_args()
spec = workload.SweepSpec(n_agents=a.agents)
h = Handle(a.replicates, spec.n_cap, spec.e_cap, a.weeks * 2 + 16, seed=1971, compact_every=64)
workload.load_sweep(h, spec)
h.set_mode(a.mode)
h.run(2, 1 + a.weeks); h.sync()
h.run(2 + a.weeks, 1 + 2 * a.weeks); h.sync()
print("ok", h.status(0))
h.close()
