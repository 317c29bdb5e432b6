#!/usr/bin/env python
"""Summarise an `ncu --set full` report: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/ncu_rNN
Writes <out>.md (key metrics per kernel, mean over captured launches) and updates profiles/traffic.json
(dram__bytes_read.sum + dram__bytes_write.sum per launch, consumed by bench.py as roofline.traffic) and
profiles/warp_inst.json (smsp__inst_executed.sum per launch, consumed by bench.py as roofline.issue_slots)."""
import collections, csv, io, json, os, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
M = {"ms": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
     "occ_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
     "winst": "smsp__inst_executed.sum", "thr_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
     "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "l2_hit": "lts__t_sector_hit_rate.pct",
     "stall_long_sb": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
     "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
     "stall_math": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
     "stall_short_sb": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "stall_branch": "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
     "stall_lg": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"}
ix = {k: hdr.index(v) for k, v in M.items() if v in hdr}
ki = hdr.index("Kernel Name")
def scale(k, v):
    u = units[ix[k]]
    v = float(v.replace(",", ""))
    if k == "ms": return v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1)
    if k in ("rd", "wr"): return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return v
agg = collections.OrderedDict()
for r in data:
    name = r[ki].split("(")[0].replace("void ", "")
    if name.startswith("k_edge2b<"):
        name = "k_edge2b_anal" if name.rstrip(">").split(",")[-1].strip() in ("1", "true") else "k_edge2b_oral_kiss_rim"
    name = name.split("<")[0]
    a = agg.setdefault(name, collections.defaultdict(list))
    for k in ix:
        try: a[k].append(scale(k, r[ix[k]]))
        except Exception: pass
mean = lambda x: sum(x) / len(x) if x else float("nan")
lines = [f"# ncu --set full summary ({os.path.basename(rep)})", "",
         "Mean over the captured launches of each kernel; `--clock-control none`. DRAM GB/s = (read+write)/duration.", "",
         "| kernel | n | ms | DRAM rd MB | DRAM wr MB | DRAM GB/s | dram % | sm % | occ % | regs | warp-inst M | thr/inst | issue % | L2 hit % | stall long_sb | wait | math | short_sb | branch | lg |",
         "|" + "---|" * 20]
traffic, winst = {}, {}
for k, a in agg.items():
    ms, rd, wr = mean(a["ms"]), mean(a["rd"]), mean(a["wr"])
    traffic[k] = rd + wr
    winst[k] = mean(a["winst"])
    lines.append(f"| {k} | {len(a['ms'])} | {ms:.4f} | {rd/1e6:.1f} | {wr/1e6:.1f} | {(rd+wr)/ms/1e6:.0f} | {mean(a['dram_pct']):.1f} | {mean(a['sm_pct']):.1f} | "
                 f"{mean(a['occ_pct']):.1f} | {mean(a['regs']):.0f} | {mean(a['winst'])/1e6:.1f} | {mean(a['thr_per_inst']):.1f} | {mean(a['issue_pct']):.1f} | {mean(a['l2_hit']):.1f} | "
                 f"{mean(a['stall_long_sb']):.2f} | {mean(a['stall_wait']):.2f} | {mean(a['stall_math']):.2f} | {mean(a['stall_short_sb']):.2f} | {mean(a['stall_branch']):.2f} | {mean(a['stall_lg']):.2f} |")
open(out + ".md", "w").write("\n".join(lines) + "\n")
tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
old = json.load(open(tp)) if os.path.exists(tp) else {}
old.update(traffic); old["_source"] = os.path.basename(out) + ".md"
json.dump(old, open(tp, "w"), indent=1)
winst["_source"] = os.path.basename(out) + ".md"
json.dump(winst, open(os.path.join(os.path.dirname(tp), "warp_inst.json"), "w"), indent=1)      # warp instructions per launch (issue-slot accounting)
print("\n".join(lines))
