"""Instructions executed / stall samples of one kernel summed over source-line ranges of one file (ncu --import-source on):
python tools/src_ranges.py report.ncu-rep kernel_regex file.cu name:lo-hi [name:lo-hi ...]"""
import csv, io, subprocess, sys
rep, k, fname = sys.argv[1:4]
ranges = [(a.split(":")[0], *map(int, a.split(":")[1].split("-"))) for a in sys.argv[4:]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + k, "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = hdr = None
acc = {n: [0, 0] for n, _, _ in ranges}; other = {}
tot = [0, 0]
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr and r[0] != "" and len(r) > ie:
        try: a, s, ln = int(r[ie]), int(r[isamp]), int(r[0])
        except ValueError: continue
        tot[0] += a; tot[1] += s
        hit = False
        if cur == fname:
            for n, lo, hi in ranges:
                if lo <= ln <= hi: acc[n][0] += a; acc[n][1] += s; hit = True; break
        if not hit:
            v = other.setdefault(cur if cur != fname else fname + ":other", [0, 0]); v[0] += a; v[1] += s
for n, (a, s) in acc.items(): print(f"{n:18s} {100*a/tot[0]:5.1f}% inst {100*s/tot[1]:5.1f}% samp")
for n, (a, s) in sorted(other.items(), key=lambda kv: -kv[1][0])[:8]: print(f"{n:18s} {100*a/tot[0]:5.1f}% inst {100*s/tot[1]:5.1f}% samp")
print("total inst", tot[0])
