"""Per-source-line hot spots of one kernel from an ncu report captured with --import-source on:
python tools/src_hot.py report.ncu-rep kernel_regex [n] [by: samp|inst]"""
import csv, subprocess, sys, io
rep, k = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 14
by = 1 if (len(sys.argv) > 4 and sys.argv[4] == "inst") else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + k, "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
out, cur, hdr = [], None, None
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr and r[0] != "" and len(r) > ie:
        try: out.append((int(r[ie]), int(r[isamp]), cur, r[0], r[1].strip()[:100]))
        except ValueError: pass
tot = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
agg = {}
for a, s, f, l, src in out:
    v = agg.setdefault((f, l, src), [0, 0]); v[0] += a; v[1] += s
for (f, l, src), (a, s) in sorted(agg.items(), key=lambda kv: -kv[1][1 - by])[:n]:
    print(f"{100 * a / tot:5.1f}% inst {100 * s / ts:5.1f}% samp  {f}:{l}  {src}")
