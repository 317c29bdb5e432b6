import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgc_msm_b200 import workload
from xgc_msm_b200.abi import Handle
R, N = int(sys.argv[1]), int(sys.argv[2])
spec = workload.SweepSpec(n_agents=N, seed=1)
h = Handle(R, spec.n_cap, spec.e_cap, 200, seed=1, device=0, compact_every=1000)
workload.load_sweep(h, spec, rid0=0)
at = 2
for rep in range(3):
    h.run(at, at + 39); at += 40; h.sync()
    t0 = time.perf_counter(); h.compact(); h.sync(); t1 = time.perf_counter()
    print(f"compact #{rep}: {(t1 - t0) * 1e3:.3f} ms")
h.profile(True); h.compact(); h.sync()
for k, v in sorted(h.profile_read().items(), key=lambda kv: -kv[1][0]): print(f"  {k:40s} {v[0]:.4f} ms x{v[1]}")
t0 = time.perf_counter(); h.run(at, at + 39); h.sync(); print("40 weeks", (time.perf_counter() - t0) * 1e3 / 40, "ms/week")
