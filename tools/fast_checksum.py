"""Checksum of the fast path's weekly tallies and final attributes on a fixed workload (A/B of kernel changes that must not move a draw)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xgc_msm_b200 import workload as W
from xgc_msm_b200.abi import Handle, K
spec = W.SweepSpec(n_agents=10000, seed=1971)
R, T = 64, 80
h = Handle(R, spec.n_cap, spec.e_cap, T + 4, seed=1971, compact_every=32)
W.load_sweep(h, spec)
h.set_mode("fast")
h.run(2, T)
m = hashlib.sha256()
for t in range(2, T + 1):
    m.update(h.counters_all(t).tobytes())
for r in (0, 17, 63):
    for a in ("rGC", "uGC", "pGC", "status", "prepStat", "last.screen", "sti.expo", "age.grp"):
        m.update(np.ascontiguousarray(h.download_attr(r, a)).tobytes())
print("fast checksum", m.hexdigest()[:16], "status", int(np.bitwise_or.reduce([h.status(r) for r in range(R)])))
h.close()
