"""Device-resident weeks/s of the calibration sweep in both modes (CUDA events on the handle's stream)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from xgc_msm_b200 import workload
from xgc_msm_b200.abi import Handle, K

ap = argparse.ArgumentParser()
ap.add_argument("--replicates", type=int, default=1000)
ap.add_argument("--agents", type=int, default=10000)
ap.add_argument("--weeks", type=int, default=128)
ap.add_argument("--modes", default="fast,strict")
a = ap.parse_args()
spec = workload.SweepSpec(n_agents=a.agents)
for mode in a.modes.split(","):
    h = Handle(a.replicates, spec.n_cap, spec.e_cap, a.weeks + 16, seed=1971, compact_every=64)
    workload.load_sweep(h, spec)
    h.set_mode(mode)
    st = torch.cuda.ExternalStream(h.stream())
    h.run(2, 9); h.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); h.run(10, 9 + a.weeks); e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    G = K.NCELL
    aw = sum(float(h.counters_all(t)[:, :G].sum()) for t in range(10, 10 + a.weeks))
    print(json.dumps({"mode": mode, "ms_per_week": ms / a.weeks, "agent_weeks_per_s": aw / (ms * 1e-3), "replicates": a.replicates, "agents": a.agents,
                      "status": int(np.bitwise_or.reduce([h.status(r) for r in range(0, a.replicates, max(1, a.replicates // 16))]))}), flush=True)
    if mode == "fast":
        import ctypes as C
        from xgc_msm_b200 import abi
        names = ["stage", "arrive", "pre", "net", "acts", "b", "trans", "apply_tally", "row", "writeback", "squeeze_wb", "stage_tables", "hostnet"] + ["-"] * 3
        buf = (C.c_ulonglong * 16)()
        abi.lib.xgc_debug_fast_profile.argtypes = [C.POINTER(C.c_ulonglong), C.c_int, C.c_int]
        n = abi.lib.xgc_debug_fast_profile(buf, 16, 1)
        tot = sum(buf[:n]) or 1
        print(json.dumps({"phase_share": {names[k]: round(buf[k] / tot, 4) for k in range(n) if buf[k]},
                          "cycles_per_replicate_week": tot / (a.replicates * (a.weeks + 8))}), flush=True)
    h.close()
