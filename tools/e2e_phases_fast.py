"""Where the end-to-end week goes in fast mode with the weekly-toggle contract: ONE handle of R replicates, every phase of the
host-network week timed with a stream sync after it (nothing overlaps), then the per-kernel device times of the same week."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from xgc_msm_b200 import workload
from xgc_msm_b200.abi import Handle, K
from xgc_msm_b200.params import control_msm

R, N = int(sys.argv[1]) if len(sys.argv) > 1 else 143, 10000
mode = sys.argv[2] if len(sys.argv) > 2 else "fast"
spec = workload.SweepSpec(n_agents=N, seed=1)
h = Handle(R, spec.n_cap, spec.e_cap, 200, seed=1, device=0, compact_every=spec.weeks_between_compactions)
workload.load_sweep(h, spec, rid0=0)
h.set_mode(mode)
h.run(2, 60)
ne_dev = h.counters_all(60)[:, K.K_NEDGES:K.K_NEDGES + 3].astype(np.float64).mean(0)
h.set_control(control_msm(network_mode="host").flags)
rng = np.random.default_rng(1)
nmin = int(h.num_agents_all().min())
oneoff = [bench.host_network_pool(R, spec.e_cap, max(int(nmin * 0.97), 16), rng, torch, n_edges=np.rint(ne_dev), layers=(2,))[0] for _ in range(4)]
deltas = [bench.host_delta_pool(R, max(int(nmin * 0.97), 16), float(ne_dev[l]), (250.0, 90.0)[l], rng, torch, weeks=4) for l in range(2)]
pre = K.M_AGING | K.M_DEPARTURE | K.M_ARRIVAL | K.M_HIVTEST | K.M_HIVTX | K.M_HIVPROGRESS | K.M_HIVVL
post = K.M_ALL & ~pre & ~K.M_RESIM_NETS
n_host = torch.empty(R, dtype=torch.int32).pin_memory().numpy()
c_host = torch.empty((R, K.NCOUNTER), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
acc = {}
SPAN = bool(int(os.environ.get("SPAN", "0")))
def T(name, f):
    t0 = time.perf_counter(); f(); h.sync(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
at = 61
def week(timed):
    global at
    run = T if timed else (lambda name, f: f())
    if not SPAN or at == 61: run("step_pre", lambda: h.step(at, pre))
    run("num_agents_all", lambda: h.num_agents_all(n_host))
    for l in range(2):
        d = deltas[l][at % 4]
        run(f"update_edgelists_{l}", lambda: h.update_edgelists_all(l, d.data_ptr(), d.shape[1], has_start=False))
    t_el, t_ne, _, _ = oneoff[at % 4]
    run("set_edgelists_2", lambda: h.set_edgelists_all(2, t_el.data_ptr(), t_ne.data_ptr(), 0, stride=t_el.shape[2]))
    run("step_post", (lambda: h.step_span(at, post, pre)) if SPAN else (lambda: h.step(at, post)))
    run("counters_all", lambda: h.counters_all(at, c_host))
    at += 1
for it in range(3): week(False)
for it in range(10): week(True)
for k, v in acc.items():
    print(f"{k:20s} {v / 10 * 1e3:8.4f} ms")
print("total (serialised)", sum(acc.values()) / 10 * 1e3, "ms per week of", R, "replicates")
t0 = time.perf_counter()
for it in range(10): week(False)
h.sync(); print("unsynchronised loop  ", (time.perf_counter() - t0) / 10 * 1e3, "ms per week")
h.profile(True)
for it in range(4): week(False)
prof = h.profile_read()
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:26s} {v[0] / 4:8.4f} ms/week  ({v[1] // 4} launches/week)")
print("status", h.status(0))
h.profile(False)
if mode == "fast":
    import ctypes as C
    from xgc_msm_b200 import abi
    names = ["stage", "arrive", "pre", "net", "acts", "b", "trans", "apply_tally", "row", "writeback", "squeeze_wb", "stage_tables", "hostnet"] + ["-"] * 3
    buf = (C.c_ulonglong * 16)()
    abi.lib.xgc_debug_fast_profile.argtypes = [C.POINTER(C.c_ulonglong), C.c_int, C.c_int]
    abi.lib.xgc_debug_fast_profile(buf, 16, 1)
    for it in range(8): week(False)
    h.sync()
    n = abi.lib.xgc_debug_fast_profile(buf, 16, 1)
    print("step-mode phase cycles per replicate-week:", {names[k]: int(buf[k] / (R * 8)) for k in range(n) if buf[k]}, "total", int(sum(buf[:n]) / (R * 8)))
    for it in range(3):                                           # per-replicate CTA time of ONE call of the weekly contract
        h.replicate_cost(reset=True)
        week(False); h.sync()
        c = h.replicate_cost().astype(np.float64) / 1.965e3      # us at 1965 MHz
        print("CTA time per replicate in one week's calls (us): min %.1f  p10 %.1f  median %.1f  mean %.1f  p90 %.1f  p99 %.1f  max %.1f" %
              (c.min(), np.percentile(c, 10), np.median(c), c.mean(), np.percentile(c, 90), np.percentile(c, 99), c.max()))
    if SPAN: sys.exit(0)
    h.set_control(control_msm(network_mode="surrogate").flags)
    h.run(at, at + 15); h.sync(); at += 16
    n = abi.lib.xgc_debug_fast_profile(buf, 16, 1)
    print("resident 16-week launch, same handle:      ", {names[k]: int(buf[k] / (R * 16)) for k in range(n) if buf[k]}, "total", int(sum(buf[:n]) / (R * 16)))
    st = torch.cuda.ExternalStream(h.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); h.run(at, at + 15); e1.record(st); torch.cuda.synchronize(); at += 16
    print("resident ms per week of", R, "replicates:", e0.elapsed_time(e1) / 16)
    for it in range(16):
        e0.record(st); h.run(at, at); e1.record(st); torch.cuda.synchronize(); at += 1
        if it >= 12: print("one-week resident launch ms:", e0.elapsed_time(e1))
