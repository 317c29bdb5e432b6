import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xgc_msm_b200 import workload as W
from xgc_msm_b200.abi import Handle, K
# poison the allocator's memory
xs = [torch.full((1 << 28,), 0x7F7F7F7F, dtype=torch.int32, device="cuda") for _ in range(24)]
torch.cuda.synchronize(); del xs; torch.cuda.empty_cache()
R, weeks = 1000, 4
spec = W.SweepSpec(n_agents=10000, seed=1971)
def make(r0, n):
    h = Handle(n, spec.n_cap, spec.e_cap, weeks + 4, seed=1971, device=0, replicate_offset=r0, compact_every=5)
    W.load_sweep(h, spec, rid0=r0)
    return h
full = make(0, R); full.run(2, 1 + weeks)
C = np.stack([full.counters_all(t) for t in range(1, 2 + weeks)]).astype(np.int64)
halves = [make(0, R // 2), make(R // 2, R // 2)]
for at in range(2, 2 + weeks):
    for h in halves: h.step(at)
for k, h in enumerate(halves):
    for t in range(2, 2 + weeks):
        a = h.counters_all(t).astype(np.int64); b = C[t - 1, k * (R // 2):(k + 1) * (R // 2)]
        if not np.array_equal(a, b):
            rr, cc = np.nonzero(a != b)
            print("half", k, "week", t, "replicates differing", np.unique(rr)[:10], "n", np.unique(rr).size, "counters", np.unique(cc)[:20], "first", a[rr[0], cc[0]], b[rr[0], cc[0]])
            break
    else:
        print("half", k, "identical")
