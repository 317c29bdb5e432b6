"""Determinism check of the bit-exact run path: repeated fresh 1000-replicate handles, `xgc_run` (CUDA-graph replay; XGC_LANES=2: two
concurrent lanes) against the first run, and a 500-replicate shard stepped week by week against the same reference.
Usage: [XGC_LANES=2] [XGC_OVERLAP=0] [XGC_PDL=0] python tools/lane_determinism.py [trials]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xgc_msm_b200 import workload as W
from xgc_msm_b200.abi import Handle, K
R, weeks, trials = 1000, 3, int(sys.argv[1]) if len(sys.argv) > 1 else 12
spec = W.SweepSpec(n_agents=10000, seed=1971)
def make(r0, n):
    h = Handle(n, spec.n_cap, spec.e_cap, weeks + 4, seed=1971, device=0, replicate_offset=r0, compact_every=50)
    W.load_sweep(h, spec, rid0=r0)
    if os.environ.get("DBG_UNIFORM"):
        from xgc_msm_b200 import synth
        t = synth.make_tables(); t[synth.T_NET_DEGW:synth.T_NET_RISKW + 5] = 1.0; h.set_tables(t)
    if os.environ.get("DBG_HOSTNET"):
        from xgc_msm_b200.params import control_msm
        h.set_control(control_msm(network_mode="host").flags)
    return h
ref = None; bad_run = bad_step = 0; info = []
for trial in range(trials):
    full = make(0, R)
    if os.environ.get("DBG_SETTLE"):
        import torch, time
        full.sync(); torch.cuda.synchronize(); time.sleep(0.05)
    if os.environ.get("DBG_STEPFIRST"):
        full.step(2); full.run(3, 1 + weeks)
    else:
        full.run(2, 1 + weeks)
    C = np.stack([full.counters_all(t) for t in range(2, 2 + weeks)]).astype(np.int64)
    full.close()
    if ref is None: ref = C
    elif not np.array_equal(ref, C):
        bad_run += 1; w, rr, cc = np.nonzero(ref != C); info.append(("run", trial, "weeks", np.unique(w + 2).tolist(), "reps", int(rr.min()), int(rr.max()), np.unique(rr).size, "counter groups", sorted(set((np.unique(cc) // 20).tolist())), "scalars", [int(c) for c in np.unique(cc) if c >= 280]))
    half = make(R // 2, R // 2)
    for at in range(2, 2 + weeks): half.step(at)
    D = np.stack([half.counters_all(t) for t in range(2, 2 + weeks)]).astype(np.int64)
    half.close()
    if not np.array_equal(ref[:, R // 2:], D):
        bad_step += 1; w, rr, cc = np.nonzero(ref[:, R // 2:] != D); info.append(("step", trial, int(w[0]) + 2, np.unique(rr)[:5].tolist(), np.unique(cc)[:8].tolist()))
print({k: os.environ.get(k) for k in ("XGC_PDL", "XGC_OVERLAP", "XGC_LANES")}, "trials", trials, "run mismatches", bad_run, "step mismatches", bad_step, info[:4])
