"""Does splitting one GPU's replicates over several handles (streams, week graphs) running concurrently beat one handle?
python tools/concurrent_handles.py [groups ...]"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xgc_msm_b200 import workload
from xgc_msm_b200.abi import Handle
R, weeks = 996, 60
spec = workload.SweepSpec(n_agents=10000, seed=1971)
for G in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    hs = []
    for gi in range(G):
        h = Handle(R // G, spec.n_cap, spec.e_cap, weeks + 16, seed=1971, device=0, replicate_offset=gi * (R // G), compact_every=1000)
        workload.load_sweep(h, spec, rid0=gi * (R // G)); h.run(2, 6); hs.append(h)
    for h in hs: h.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=lambda h=h: (h.run(7, 6 + weeks), h.sync())) for h in hs]
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print(f"groups {G}: {dt / weeks * 1e3:.4f} ms/week for {R} replicates")
    for h in hs: h.close()
