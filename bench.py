#!/usr/bin/env python
"""bench.py — agent-weeks/s of the weekly epidemic step (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus N --steps K --warmup W            # CUDA path (libxgc.so through the C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU restatement on the host cores

Workload (N = 1): BASELINE.json configs[1], "calibration sweep: 1000 replicates x 10k agents batched on 1 B200"; a
*step* is one simulated week of all replicates (all 17 modules).  For N > 1 every rank holds its own 1000 replicates
(weak scaling; replicates are independent, the only exchange is the final gather of the targets matrix).
The reference's R modules cannot run here (R absent, module source not in /root/reference: SURVEY §8c), so the
reference arm and `cpu_baseline` time the in-repo C restatement (oracle/, kind "port"), one single-threaded process per
replicate per core like the reference's SLURM arrays.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "agent-weeks/sec", "agent-weeks/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=260)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="xgc", choices=["xgc", "reference"])
    ap.add_argument("--replicates", type=int, default=1000, help="replicates per GPU (weak scaling)")
    ap.add_argument("--total-replicates", type=int, default=0,
                    help="strong scaling: this many replicates in total, sharded over the GPUs (north_star: the 1000-replicate sweep on 8 GPUs = 125 per GPU)")
    ap.add_argument("--mode", default="fast", choices=["fast", "strict"],
                    help="native-draw contract of the CUDA arm: fast = replicate-resident kernel (statistical equivalence), strict = bit-exact kernels")
    ap.add_argument("--agents", type=int, default=10000)
    ap.add_argument("--e2e-steps", type=int, default=24)
    ap.add_argument("--e2e-separate-calls", action="store_true", help="fast mode: the week as separate ABI calls instead of one xgc_week_hostnet call")
    ap.add_argument("--e2e-pack-by-cost", action="store_true", help="e2e handles hold replicates of similar cost instead of blocks of consecutive replicate ids")
    ap.add_argument("--e2e-group-size", type=int, default=0, help="replicates per e2e handle (default: R split evenly over --e2e-groups handles)")
    ap.add_argument("--e2e-no-span", action="store_true", help="fast mode: two xgc_step calls per week instead of xgc_step_span")
    ap.add_argument("--e2e-groups", type=int, default=0, help="handles (streams) the e2e arm splits this GPU's replicates into")
    ap.add_argument("--profile-steps", type=int, default=6)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=1971)
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_config(args, parallelism, mode=None):
    c = {"workload": f"calibration sweep (BASELINE configs[1]): {args.replicates} replicates x {args.agents} agents per GPU, "
                     "per-replicate LHS parameters over the reference prior box, all 17 weekly modules, one step = one week",
         "replicates_per_gpu": args.replicates, "agents_per_replicate": args.agents, "network": "on-device surrogate for simnet_msm",
         "parallelism": parallelism,
         "l2": f"no flush: every week sweeps the agent planes of all replicates ({args.replicates} x {args.agents} agents x ~100 B = "
               f"{args.replicates * args.agents * 100 / 1e9:.2f} GB per GPU against 126 MB of L2; smaller than L2 only for shapes other than the default); the resident (fast) "
               "kernel stages each replicate in shared memory for up to 64 weeks by design, so its HBM traffic is far below the per-week model"}
    if args.total_replicates:
        c["total_replicates"] = args.total_replicates
    if mode:
        c["mode"] = mode
    return c


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement on all host cores (rank 0 only)
# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    if rank != 0:
        return
    from oracle import cpu_bench
    from xgc_msm_b200.params import MODULE_BIT
    mask = sum(MODULE_BIT.values())
    cores = os.cpu_count() or 1
    reps = cores * 8                                  # bounded sample: 8 replicates per core per step (keeps the per-step
                                                      # synchronisation of the worker pool below 1 % of a step)
    pool = cpu_bench.PersistentPool(args.agents, reps, args.seed, mask, cores)
    at = 2
    for _ in range(args.warmup):
        pool.step(at); at += 1
    aw = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        aw += pool.step(at); at += 1
    dt = time.perf_counter() - t0
    pool.close()
    v = aw / dt
    sample = f"{reps} of the workload's {args.replicates} replicates x {args.agents} agents, one week per step, {cores} single-threaded processes"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+int",
        "data": "synthetic", "config": workload_config(args, "cpu: one process per replicate per core"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "R and the EpiModelHIV modules are not available; this arm times oracle/ (plain-C restatement, -O2, scalar)"}), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# helpers for the CUDA arm
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: an in-process NVML sampling thread (one sample taken
    synchronously at start and at stop, so even a very short region has some), nvidia-smi polling as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index, uuid=None):
        self.sm, self.mx, self.reasons = [], [], set()
        self.p = self.th = self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                self.dev = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode()) if uuid else pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            except Exception:
                self.dev = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
            self._sample()
            import threading
            self._stop = threading.Event()
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nv = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))); self.mx.append(self.max_sm)
        try:
            bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev))
        except Exception:
            bits = 0
        for bit, name in self.REASONS.items():
            if bits & bit:
                self.reasons.add(name)

    def _loop(self):
        while not self._stop.wait(0.01):
            try:
                self._sample()
            except Exception:
                return

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.nv is not None:
            self._stop.set(); self.th.join(timeout=2)
            try:
                self._sample()
            except Exception:
                pass
        elif self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
            self.f.flush(); self.f.seek(0)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in self.f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    self.sm.append(float(c[1])); self.mx.append(float(c[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, c[5:9]):
                    if val.lower().startswith("active"):
                        self.reasons.add(nm)
            os.unlink(self.f.name)
        if self.sm:
            out.update(sm_mhz=float(np.median(self.sm)), sm_max_mhz=float(max(self.mx)), reasons=sorted(self.reasons), samples=len(self.sm))
        return out


def week_stats(h, K, at):
    """Population fractions of week `at` over all replicates of this rank (inputs of roofline.kernel_bytes)."""
    c = h.counters_all(at).astype(np.float64).sum(0)
    G = K.NCELL
    grp = lambda g: c[g * G:(g + 1) * G].sum()
    n = grp(K.K_NUM)
    incid = grp(K.K_INCID_RGC) + grp(K.K_INCID_UGC) + grp(K.K_INCID_PGC) + grp(K.K_INCID_HIV)
    return {"n": n, "f_hiv": grp(K.K_INUM) / n, "f_dx": grp(K.K_INUM_DX) / n, "f_gc": grp(K.K_INUM_GC) / n, "f_new": incid / n,
            "edges": float(c[K.K_NEDGES:K.K_NEDGES + 3].sum())}


def host_network_pool(R, e_cap, n_lo, rng, torch, n_edges=None, layers=(0, 1, 2)):
    """Pinned host edge lists standing for what host-side simnet_msm returns each week: [R][2][ne] 1-based positions.
    n_edges: edges per layer (default: the layer's design equilibrium); the bench passes the counts the device-resident
    arm actually holds, so that both arms process the same number of edges."""
    bufs = []
    for l in layers:
        ne = int(n_edges[l]) if n_edges is not None else int(e_cap[l] / 1.35) - 100
        ne = max(1, min(ne, int(e_cap[l])))            # buffers hold exactly ne edges
        a = rng.integers(1, n_lo, size=(R, ne), dtype=np.int32)
        b = rng.integers(1, n_lo - 1, size=(R, ne), dtype=np.int32)
        b = np.where(b >= a, b + 1, b)
        el = np.stack([np.minimum(a, b), np.maximum(a, b)], axis=1).astype(np.int32)          # [R,2,e_cap]
        t_el = torch.from_numpy(el).pin_memory()
        t_ne = torch.full((R,), ne, dtype=torch.int32).pin_memory()
        t_st = torch.from_numpy((1 - rng.geometric(1 / 100.0, size=(R, ne))).astype(np.int32)).pin_memory()
        bufs.append((t_el, t_ne, t_st, ne))
    return bufs


def host_delta_pool(R, n_lo, ne, dur, rng, torch, weeks=8):
    """Pinned weekly deltas of ONE persistent layer for `weeks` weeks (cycled): packed int32 rows [R][stride] =
    {n_removed, n_added, removed indices (ascending, 0-based), tails, heads, start weeks} -- what a host-side simnet_msm has to
    send when the device keeps the layer: ~ne/dur dissolved ties and as many (+2: ties lost to departures) formed ones."""
    out = []
    lam = ne / dur
    for _ in range(weeks):
        nrem = rng.poisson(lam, size=R).astype(np.int64)
        nadd = nrem + 2
        stride = int(2 + nrem.max() + 3 * nadd.max())
        d = np.zeros((R, stride), np.int32)
        for r in range(R):
            k, m = int(nrem[r]), int(nadd[r])
            rem = np.sort(rng.choice(int(ne * 0.9), size=k, replace=False)).astype(np.int32)
            a = rng.integers(1, n_lo, size=m); b = rng.integers(1, n_lo - 1, size=m); b = np.where(b >= a, b + 1, b)
            d[r, 0], d[r, 1] = k, m
            d[r, 2:2 + k] = rem
            d[r, 2 + k:2 + k + m] = np.minimum(a, b); d[r, 2 + k + m:2 + k + 2 * m] = np.maximum(a, b)
            d[r, 2 + k + 2 * m:2 + k + 3 * m] = 0            # start week: filled by the device (has_start = 0 would do; kept in the row for the byte count)
        out.append(torch.from_numpy(d).pin_memory())
    return out


def bind_to_gpu_numa_node(torch, local_rank):
    """Several ranks share one host: keep this rank's threads, and therefore its page-locked staging buffers (first
    touch), on the NUMA node its GPU hangs off, so that the weekly uploads of 8 ranks do not cross the socket link.
    Best effort: returns the node or None."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) >= 2:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def run_xgc(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from xgc_msm_b200 import roofline, workload
    from xgc_msm_b200.abi import Handle, K

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the xgc arm has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rid_base = rank * args.replicates
    if args.total_replicates:                          # strong scaling: the sweep is split over the ranks (shard.shard_range)
        from xgc_msm_b200.shard import shard_range
        rid_base, args.replicates = shard_range(args.total_replicates, rank, world)
    R = args.replicates
    spec = workload.SweepSpec(n_agents=args.agents, seed=args.seed)
    W, Kt, Pn, En = max(args.warmup, 3), args.steps, args.profile_steps, args.e2e_steps
    t_cap = 2 + W + Kt + Pn + En + 8
    h = Handle(R, spec.n_cap, spec.e_cap, t_cap, seed=args.seed, device=local_rank, replicate_offset=rid_base,
               compact_every=spec.weeks_between_compactions)
    workload.load_sweep(h, spec, rid0=rid_base)
    mode = args.mode
    if mode == "fast":
        try:
            h.set_mode("fast")
        except Exception as ex:                         # a replicate that does not fit one SM's shared memory runs on the strict kernels
            mode = f"strict (fast unavailable: {ex})"
    stream = torch.cuda.ExternalStream(h.stream(), device=torch.device("cuda", local_rank))

    # ---- kernel-resident throughput: W warm-up weeks, then exactly K timed weeks (CUDA-graph replay, no host sync)
    at = 2
    h.run(at, at + W - 1); at += W
    barrier()
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)      # robust to CUDA_VISIBLE_DEVICES remapping
    except Exception:
        uuid = None
    clocks = ClockSampler(local_rank, uuid)
    l0 = h.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    h.run(at, at + Kt - 1)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = h.kernel_launches() - l0
    clk = clocks.stop()
    t_first, t_last = at, at + Kt - 1
    at += Kt
    G = K.NCELL
    aw = 0.0
    for t in range(t_first, t_last + 1):
        aw += float(h.counters_all(t)[:, K.K_NUM * G:(K.K_NUM + 1) * G].sum())
    if world > 1:
        tt = torch.tensor([ms], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
        ta = torch.tensor([aw], device="cuda", dtype=torch.float64); dist.all_reduce(ta); aw = float(ta.item())
    value = aw / (ms * 1e-3)

    # ---- per-kernel device times (CUDA events around every launch on the launching stream) -> roofline of the top kernel
    h.profile(True)
    if mode == "fast":
        # one resident launch over Pn weeks (no compaction boundary inside: at .. at + Pn - 1 lies between multiples of 64)
        while at % spec.weeks_between_compactions > spec.weeks_between_compactions - Pn or at % spec.weeks_between_compactions == 0:
            h.profile(False); h.run(at, at); at += 1; h.profile(True)
        h.run(at, at + Pn - 1); at += Pn
    else:
        for _ in range(Pn):
            h.step(at); at += 1
    prof = h.profile_read()
    h.profile(False)
    stats = week_stats(h, K, at - 1)
    kb = roofline.kernel_bytes(stats)
    per = {k: v[0] / max(v[1], 1) * (v[1] / Pn) for k, v in prof.items()}          # ms per week per kernel
    step_ms_prof = sum(per.values())
    peak, peak_src = measured_peak()
    if mode == "fast":
        # the dominant (only) kernel of the step is k_week_fast: one launch = Pn weeks of every replicate.  Algorithmic bytes per
        # launch = Pn x the per-week bytes of the packed layout (what a week-at-a-time implementation must
        # stream; SURVEY 8(d) x agent-weeks) -- the resident kernel keeps most of them on chip, see `traffic`
        top = "k_week_fast"
        kb = dict(kb); kb[top] = roofline.step_bytes(stats) * Pn
    else:
        top = max((k for k in per if k in kb), key=lambda k: per[k])
    top_ms = prof[top][0] / prof[top][1]
    achieved = kb[top] / (top_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": kb[top], "launch_ms": top_ms, "share_of_step": per[top] / step_ms_prof,
            "bytes_model": "per-week HBM bytes of the packed SoA layout (xgc_msm_b200/roofline.py), per agent-week: "
                           f"{roofline.step_bytes(stats) / max(stats['n'], 1):.0f} B (SURVEY 8(d) first cut: 128 B)",
            "kernels_ms_per_step": {k: round(v, 4) for k, v in sorted(per.items(), key=lambda kv: -kv[1])},
            "step": {"algorithmic_bytes": roofline.step_bytes(stats), "achieved": roofline.step_bytes(stats) / (ms / Kt * 1e-3) / 1e9,
                     "frac": roofline.step_bytes(stats) / (ms / Kt * 1e-3) / 1e9 / peak}}
    # what actually bounds the step: issue slots.  warp instructions per launch come from the committed ncu
    # capture, the time is this run's; peak = SMs x 4 schedulers x 1 warp instruction per clock at the sampled SM clock
    wi = os.path.join(ROOT, "profiles", "warp_inst.json")
    if os.path.exists(wi) and clk.get("sm_mhz"):
        try:
            w = json.load(open(wi))
            per_week = float(sum(w[k] * (prof[k][1] / Pn) for k in per if k in w))
            sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
            peak_i = sms * 4 * clk["sm_mhz"] * 1e6
            roof["issue_slots"] = {"warp_inst_per_step": per_week, "achieved_ginst_s": per_week / (ms / Kt * 1e-3) / 1e9,
                                   "peak_ginst_s": peak_i / 1e9, "frac": per_week / (ms / Kt * 1e-3) / peak_i, "source": w.get("_source")}
        except Exception:
            pass
    tr = os.path.join(ROOT, "profiles", "traffic.json")          # ncu dram bytes per launch, committed with the profile
    if os.path.exists(tr):
        try:
            roof["traffic"] = json.load(open(tr)).get(top)
        except Exception:
            pass

    # ---- end to end through the C ABI with HOST buffers: the weekly R <-> device contract with simnet_msm on the host
    #      (pre-network modules -> living counts to host -> three edge lists from pinned host memory -> remaining
    #       modules -> the week's tallies back to host).  The replicates of this GPU are served as `--e2e-groups` handles
    #      (each with its own stream and host thread), so the edge-list upload of one group overlaps the kernels of the
    #      others; every group still goes through the full serial contract every week.
    from xgc_msm_b200.params import control_msm
    # default: 8 host threads per GPU, fewer when the ranks of this node would oversubscribe its cores
    Ge = args.e2e_groups if args.e2e_groups > 0 else min(8, max(2, (os.cpu_count() or 8) // max(world, 1)))
    if args.e2e_groups <= 0 and mode == "fast":
        # the resident kernel runs one CTA per replicate: groups of a little under one CTA per SM make every call one wave and leave
        # ~20 SMs to the small host-network kernels, whose CTAs cannot share an SM with a resident CTA (8 x 125: 0.849 ms per week,
        # 7 x 143: 0.865, 10 x 100: 0.854)
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        Ge = min(Ge, max(2, -(-R // max(sms - 20, 1))))
    Ge = max(1, min(Ge, R))
    g_size = [R // Ge + (1 if gi < R % Ge else 0) for gi in range(Ge)]
    if args.e2e_group_size > 0:                                   # explicit handle size (the last handle takes the remainder)
        Ge = -(-R // args.e2e_group_size)
        g_size = [min(args.e2e_group_size, R - gi * args.e2e_group_size) for gi in range(Ge)]
    g_off = [sum(g_size[:gi]) for gi in range(Ge)]
    # one wave or less per GPU (the 8-GPU split of the 1000-replicate sweep: 125 per GPU): the week is the latency of ONE handle's round
    # trip, and there the resident kernel applying the queued edge-list operations itself is worth 5 %; with several
    # waves per GPU the standalone kernels are faster
    if mode == "fast" and R <= torch.cuda.get_device_properties(local_rank).multi_processor_count:
        os.environ.setdefault("XGC_FUSE_HOSTNET", "1")
    groups = []
    # --e2e-pack-by-cost: handles of replicates of SIMILAR cost (ranked by the SM cycles the device-resident arm above spent on each:
    # xgc_replicate_cost; results do not depend on the packing: xgc_set_replicate_ids, Philox is keyed by the global id).  A call of the
    # resident kernel is one wave and lasts as long as its slowest replicate (CTA times 90 .. 130 us around a mean of 109), but the
    # measured week does not get shorter (0.948 vs 0.947 ms): the other handles' kernels already fill the tails.  Off by default.
    pack = None
    if mode == "fast" and args.e2e_pack_by_cost and Ge > 1:
        cost = h.replicate_cost()
        if cost.max() > 0:
            pack = np.argsort(cost, kind="stable")
    for gi in range(Ge):
        hg = Handle(g_size[gi], spec.n_cap, spec.e_cap, t_cap, seed=args.seed, device=local_rank, replicate_offset=rid_base + g_off[gi],
                    compact_every=spec.weeks_between_compactions)
        if pack is None:
            workload.load_sweep(hg, spec, rid0=rid_base + g_off[gi])
        else:
            workload.load_sweep(hg, spec, ids=rid_base + pack[g_off[gi]:g_off[gi] + g_size[gi]])
        if mode == "fast":
            hg.set_mode("fast")
        hg.run(2, at - 1)                                   # same point of the epidemic as the kernel-resident arm
        hg.set_control(control_msm(network_mode="host").flags)
        groups.append(hg)
    rng = np.random.default_rng(args.seed + rank)
    nmin = min(int(hg.num_agents_all().min()) for hg in groups)
    ne_dev = h.counters_all(t_last)[:, K.K_NEDGES:K.K_NEDGES + 3].astype(np.float64).mean(0)     # edges per layer the resident arm held
    # one-off partnerships are new every week: cycle through several one-off layers
    oneoff = [host_network_pool(R, spec.e_cap, max(int(nmin * 0.97), 16), rng, torch, n_edges=np.rint(ne_dev), layers=(2,))[0] for _ in range(8)]
    pre = K.M_AGING | K.M_DEPARTURE | K.M_ARRIVAL | K.M_HIVTEST | K.M_HIVTX | K.M_HIVPROGRESS | K.M_HIVVL
    post = K.M_ALL & ~pre & ~K.M_RESIM_NETS
    n_host = torch.empty(R, dtype=torch.int32).pin_memory()
    c_host = torch.empty((R, K.NCOUNTER), dtype=torch.int32).pin_memory()
    # main / casual ties persist for years: the device keeps those layers and the host sends the week's toggles only
    # (xgc_update_edgelists_all); the one-off layer is replaced every week (xgc_set_edgelists_all, no start weeks)
    deltas = [host_delta_pool(R, max(int(nmin * 0.97), 16), float(ne_dev[l]), (250.0, 90.0)[l], rng, torch) for l in range(2)]
    h2d = int(np.mean([sum(deltas[l][w].numel() * 4 for l in range(2)) for w in range(len(deltas[0]))])) + oneoff[0][0].numel() * 4 + oneoff[0][1].numel() * 4
    d2h = n_host.numel() * 4 + c_host.numel() * 4
    n_np, c_np = n_host.numpy(), c_host.numpy().view(np.uint32)

    # fast mode: the acts..prevalence call of week t also runs the aging..hivvl group of week t + 1 (xgc_step_span: what the last
    # module wrapper of a week issues; netsim has nothing on the host between prevalence_msm(t) and aging_msm(t + 1)) -- one
    # resident launch per week; every week still makes the full round trip living counts -> host -> edge lists -> device
    span = mode == "fast" and not args.e2e_no_span
    pre_done = [-1] * Ge

    # raw host pointers of every group's slice of the weekly buffers (pinned), taken once: the timed loop only passes integers
    esz = 4
    dptr = [[[d.data_ptr() + g_off[gi] * d.shape[1] * esz for d in deltas[l]] for l in range(2)] for gi in range(Ge)]
    optr = [[(o[0].data_ptr() + g_off[gi] * 2 * o[0].shape[2] * esz, o[1].data_ptr() + g_off[gi] * esz, o[0].shape[2]) for o in oneoff] for gi in range(Ge)]
    cptr = [c_host.data_ptr() + g_off[gi] * K.NCOUNTER * esz for gi in range(Ge)]
    nptr = [n_host.data_ptr() + g_off[gi] * esz for gi in range(Ge)]

    def e2e_week(gi, t):
        """One week of the contract for group gi (what one host-side simnet_msm loop does); returns living agents."""
        hg, r0, Rg = groups[gi], g_off[gi], g_size[gi]
        if span and not args.e2e_separate_calls:
            # ONE ABI call a week (xgc_week_hostnet): toggles + one-off list in, rest of week t + start of week t + 1, tallies of week t
            # and the living counts for the next network step out
            if pre_done[gi] != t:
                hg.step(t, pre); hg.num_agents_all(n_np[r0:r0 + Rg])
            d0, d1 = deltas[0][t % len(deltas[0])], deltas[1][t % len(deltas[1])]
            oe, on, ost = optr[gi][t % len(oneoff)]
            hg.week_hostnet(t, post, pre, dptr[gi][0][t % len(deltas[0])], d0.shape[1], dptr[gi][1][t % len(deltas[1])], d1.shape[1], False,
                            oe, on, ost, cptr[gi], nptr[gi])
            pre_done[gi] = t + 1
            return float(c_np[r0:r0 + Rg, K.K_NUM * G:(K.K_NUM + 1) * G].sum())
        if pre_done[gi] != t:
            hg.step(t, pre)
        hg.num_agents_all(n_np[r0:r0 + Rg])                  # what host-side simnet_msm needs before it can return edges
        for l in range(2):
            d = deltas[l][t % len(deltas[l])]
            hg.update_edgelists_all(l, d[r0:r0 + Rg].data_ptr(), d.shape[1], has_start=False)
        t_el, t_ne, _t_st, _ne = oneoff[t % len(oneoff)]
        hg.set_edgelists_all(2, t_el[r0:r0 + Rg].data_ptr(), t_ne[r0:r0 + Rg].data_ptr(), 0, stride=t_el.shape[2])
        if span:
            hg.step_span(t, post, pre); pre_done[gi] = t + 1
        else:
            hg.step(t, post)
        hg.counters_all(t, c_np[r0:r0 + Rg])
        return float(c_np[r0:r0 + Rg, K.K_NUM * G:(K.K_NUM + 1) * G].sum())

    # one host thread per group, like the independent host processes of the reference's job array; ctypes drops the GIL
    # inside every ABI call, so one group's blocking device->host read does not stall the others' uploads and launches
    import threading
    aw_g = [0.0] * Ge
    errs = []

    def drive(gi, t_first, n_weeks):
        try:
            torch.cuda.set_device(local_rank)
            acc = 0.0
            for t in range(t_first, t_first + n_weeks):
                acc += e2e_week(gi, t)
            aw_g[gi] = acc
        except Exception as ex:                               # surface in the main thread
            errs.append(ex)

    def run_weeks(t_first, n_weeks):
        th = [threading.Thread(target=drive, args=(gi, t_first, n_weeks)) for gi in range(Ge)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        if errs:
            raise errs[0]
        return sum(aw_g)

    run_weeks(at, 3); at += 3
    if os.environ.get("XGC_DEBUG_COST") and mode == "fast":
        for hg in groups:
            hg.replicate_cost(reset=True)
    barrier()
    t0 = time.perf_counter()
    aw_e = run_weeks(at, En); at += En
    barrier()
    dt_e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt_e], device="cuda", dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX); dt_e = float(tt.item())
        ta = torch.tensor([aw_e], device="cuda", dtype=torch.float64); dist.all_reduce(ta); aw_e = float(ta.item())
    e2e_packing = "replicates of similar cost per handle (xgc_replicate_cost ranks, xgc_set_replicate_ids)" if pack is not None else "consecutive replicate ids per handle"
    e2e = {"value": aw_e / dt_e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": En,
           "ms_per_step": dt_e / En * 1e3, "groups": Ge, "packing": e2e_packing, "numa_node": numa,
           "path": (f"{Ge} handles x {g_size[0]} replicates (one host thread each), each week ONE call: xgc_week_hostnet = 2x xgc_update_edgelists_all (main / casual toggles) + "
                    "xgc_set_edgelists_all (one-off layer), pinned -> xgc_step_span(post of this week, pre of the next: one resident launch) -> xgc_get_counters_all + xgc_num_agents_all")
                   if span and not args.e2e_separate_calls else
                   f"{Ge} handles x {g_size[0]} replicates (one host thread each), each week: " + ("" if span else "xgc_step(pre) -> ") + "xgc_num_agents_all -> 2x xgc_update_edgelists_all "
                   "(main / casual toggles) + xgc_set_edgelists_all (one-off layer), pinned -> " +
                   ("xgc_step_span(post of this week, pre of the next: one resident launch)" if span else "xgc_step(post)") + " -> xgc_get_counters_all"}
    if os.environ.get("XGC_DEBUG_COST") and mode == "fast":      # how well does the resident arm's per-replicate cost predict the one-week calls'?
        c_res = h.replicate_cost().astype(np.float64)
        c_e2e = np.zeros(R)
        for gi, hg in enumerate(groups):
            idx = pack[g_off[gi]:g_off[gi] + g_size[gi]] if pack is not None else np.arange(g_off[gi], g_off[gi] + g_size[gi])
            c_e2e[idx] = hg.replicate_cost().astype(np.float64)
        gmax = [float(groups[gi].replicate_cost().max()) for gi in range(Ge)]
        print("cost: corr(resident, e2e) = %.3f; e2e CTA cycles per replicate-week: min %.3g mean %.3g max %.3g; sum of per-handle maxima / (mean x handles) = %.3f" %
              (np.corrcoef(c_res, c_e2e)[0, 1], c_e2e.min() / En, c_e2e.mean() / En, c_e2e.max() / En, sum(gmax) / (c_e2e.mean() * Ge)), file=sys.stderr)
    e2e_status = 0
    for hg in groups:
        e2e_status |= int(np.bitwise_or.reduce([hg.status(r) for r in range(0, hg.n_replicates, max(hg.n_replicates // 8, 1))]))
        hg.close()
    e2e["status_sample_or"] = e2e_status

    # ---- final exchange of the path: gather the [R x K] targets matrix (NCCL all_gather), outside the timed regions
    tg = torch.from_numpy(h.targets(t_last, min(Kt, 52))).cuda()
    if world > 1:
        parts = [torch.empty_like(tg) for _ in range(world)]
        dist.all_gather(parts, tg)
        tg = torch.cat(parts)
    status = [h.status(r) for r in range(0, R, max(R // 16, 1))]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_bench
        cores = os.cpu_count() or 1
        reps = cores * 2
        weeks = max(int(args.cpu_seconds * 0.9e6 / (2 * args.agents)), 4)        # ~0.9M agent-weeks/s/core for the C port
        aw_c, dt_c, cores, _ = cpu_bench.run_batch(args.agents, reps, weeks, args.seed, K.M_ALL, cores)
        cpu = {"value": aw_c / dt_c, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{reps} replicates x {args.agents} agents x {weeks} weeks of the same sweep, one single-threaded process per core "
                         f"(oracle/xgc_oracle.c, gcc -O2)",
               "reference_floor": "the reference's SLURM limits imply >= 5.8k-8.7k agent-weeks/s/core for the whole R loop incl. network resimulation (BASELINE.md)"}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": Kt, "warmup": W, "ms_per_step": ms / Kt,
            "higher_is_better": True, "scaling": "strong" if args.total_replicates else "weak", "vs_baseline": None,
            "dtype": "f32+int (fast native-draw path; injected draws: f64+int)" if mode == "fast" else "f64+int", "data": "synthetic",
            "config": workload_config(args, f"replicates sharded over {world} GPU(s), no data-path collective; final all_gather of targets", mode),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
            "targets_shape": list(tg.shape), "status_sample_or": int(np.bitwise_or.reduce(status)),
            "agent_weeks_timed": aw}), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit(f"bench.py --gpus {args.gpus} must be launched with torchrun (one process per GPU); see the module docstring")
    run_xgc(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
