"""Where the end-to-end week goes: times every phase of the host-network contract on ONE handle with a stream sync
after each (so nothing overlaps) and the raw pinned host->device copy rate.  Diagnostic for bench.py's e2e arm."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from xgc_msm_b200 import workload
from xgc_msm_b200.abi import Handle, K
from xgc_msm_b200.params import control_msm

R, N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 10000
spec = workload.SweepSpec(n_agents=N, seed=1)
h = Handle(R, spec.n_cap, spec.e_cap, 200, seed=1, device=0, compact_every=spec.weeks_between_compactions)
workload.load_sweep(h, spec, rid0=0)
h.run(2, 60)
h.set_control(control_msm(network_mode="host").flags)
rng = np.random.default_rng(1)
pool = bench.host_network_pool(R, spec.e_cap, int(h.num_agents_all().min() * 0.97), rng, torch,
                               n_edges=np.rint(h.counters_all(60)[:, K.K_NEDGES:K.K_NEDGES + 3].astype(np.float64).mean(0)))
pre = K.M_AGING | K.M_DEPARTURE | K.M_ARRIVAL | K.M_HIVTEST | K.M_HIVTX | K.M_HIVPROGRESS | K.M_HIVVL
post = K.M_ALL & ~pre & ~K.M_RESIM_NETS
n_host = torch.empty(R, dtype=torch.int32).pin_memory().numpy()
c_host = torch.empty((R, K.NCOUNTER), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
acc = {}
def T(name, f):
    t0 = time.perf_counter(); f(); h.sync(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
at = 61
for it in range(13):
    if it == 3: acc.clear()
    T("step_pre", lambda: h.step(at, pre))
    T("num_agents_all", lambda: h.num_agents_all(n_host))
    for l, (t_el, t_ne, t_st, _) in enumerate(pool):
        T(f"set_edgelists_{l}", lambda: h.set_edgelists_all(l, t_el.data_ptr(), t_ne.data_ptr(), t_st.data_ptr(), stride=t_el.shape[2]))
    T("step_post", lambda: h.step(at, post))
    T("counters_all", lambda: h.counters_all(at, c_host))
    at += 1
for k, v in acc.items():
    print(f"{k:18s} {v / 10 * 1e3:8.3f} ms")
print("total", sum(acc.values()) / 10 * 1e3)
# raw pinned H2D rate
for t_el, _, t_st, _ in pool:
    d = torch.empty_like(t_el, device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): d.copy_(t_el, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"raw H2D {t_el.numel() * 4 / 1e6:.1f} MB in {dt * 1e3:.3f} ms = {t_el.numel() * 4 / dt / 1e9:.1f} GB/s")
# per-kernel device times of the host-network week
h.profile(True)
for it in range(4):
    h.step(at, pre); h.num_agents_all(n_host)
    for l, (t_el, t_ne, t_st, _) in enumerate(pool):
        h.set_edgelists_all(l, t_el.data_ptr(), t_ne.data_ptr(), t_st.data_ptr(), stride=t_el.shape[2])
    h.step(at, post); h.counters_all(at, c_host); at += 1
prof = h.profile_read()
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:26s} {v[0] / 4:8.4f} ms/week  ({v[1] // 4} launches/week)")
print("edges/week", int(c_host[:, K.K_NEDGES:K.K_NEDGES + 3].sum()), "queues?", )
