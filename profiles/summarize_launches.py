#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python profiles/summarize_launches.py in.csv out.md "title" "command"."""
import collections, csv, sys
src, out, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]; kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[kn].split("(")[0].replace("void ", "")
    if name.startswith("k_edge2b<"):
        name = "k_edge2b_anal" if name.rstrip(">").split(",")[-1].strip() in ("1", "true", "(bool)1") else "k_edge2b_oral_kiss_rim"
    name = name.split("<")[0]
    ns = float(r[mv].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[mu], 1)
    agg.setdefault(name, []).append(ns)
tot = sum(sum(v) for v in agg.values())
with open(out, "w") as f:
    f.write(f"# {title}\n\nCommand: `{cmd}`\n(after the same command exited 0 without ncu). Per-launch times are cold-cache and serialised: compare SHARES with "
            "bench.py's `roofline.kernels_ms_per_step`.\n\n| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| {k} | {len(v)} | {sum(v) / 1e6:.3f} | {sum(v) / len(v) / 1e6:.4f} | {sum(v) / tot:.3f} |\n")
print(open(out).read())
