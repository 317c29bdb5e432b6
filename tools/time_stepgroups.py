"""GPU cost of the weekly host-network call pattern in fast mode: step(pre) + step(post) per week (state staged and written
back twice a week) against the resident multi-week run, and the phase clocks of the split pattern."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
from xgc_msm_b200 import abi, workload
from xgc_msm_b200.abi import Handle, K

ap = argparse.ArgumentParser()
ap.add_argument("--replicates", type=int, default=1000)
ap.add_argument("--agents", type=int, default=10000)
ap.add_argument("--weeks", type=int, default=40)
a = ap.parse_args()
spec = workload.SweepSpec(n_agents=a.agents)
h = Handle(a.replicates, spec.n_cap, spec.e_cap, a.weeks + 32, seed=1971, compact_every=64)
workload.load_sweep(h, spec)
h.set_mode("fast")
st = torch.cuda.ExternalStream(h.stream())
pre = K.M_AGING | K.M_DEPARTURE | K.M_ARRIVAL | K.M_HIVTEST | K.M_HIVTX | K.M_HIVPROGRESS | K.M_HIVVL
post = K.M_ALL & ~pre
h.run(2, 9); h.sync()
names = ["stage", "arrive", "pre", "net", "acts", "b", "trans", "apply_tally", "row", "writeback", "squeeze_wb", "stage_tables"] + ["-"] * 4
buf = (C.c_ulonglong * 16)()
abi.lib.xgc_debug_fast_profile.argtypes = [C.POINTER(C.c_ulonglong), C.c_int, C.c_int]
abi.lib.xgc_debug_fast_profile(buf, 16, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for at in range(10, 10 + a.weeks):
    h.step(at, pre); h.step(at, post)
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.weeks
n = abi.lib.xgc_debug_fast_profile(buf, 16, 1)
tot = sum(buf[:n]) or 1
print(json.dumps({"pattern": "step(pre) + step(post) per week", "replicates": a.replicates, "ms_per_week": ms,
                  "phase_share": {names[k]: round(buf[k] / tot, 4) for k in range(n) if buf[k]},
                  "cycles_per_replicate_week": tot / (a.replicates * a.weeks)}))
h.close()
