"""Regions of an `ncu --page source --csv --print-source sass` dump with similar execution counts (first kernel copy only).
Usage: python tools/sass_regions.py dump.csv [min_share_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ia = hdr.index("Source"); ie = hdr.index("Instructions Executed"); it = hdr.index("Avg. Threads Executed")
data = []
for r in rows[2:]:
    try: data.append((int(r[ie]), r[ia].strip(), float(r[it])))
    except Exception: pass
# keep the first copy (listing repeats when the report holds several launches of the kernel)
for k in range(1, len(data)):
    if data[k][1] == data[0][1] and data[k + 1][1] == data[1][1] and k > 50:
        data = data[:k]; break
tot = sum(d[0] for d in data); mn = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print("instructions in kernel", len(data), "executed warp insts", tot)
i = 0
while i < len(data):
    j = i; n = data[i][0]
    while j + 1 < len(data) and abs(data[j + 1][0] - n) <= 0.05 * max(n, 1): j += 1
    s = sum(d[0] for d in data[i:j + 1])
    if 100 * s / tot >= mn:
        thr = sum(d[2] * d[0] for d in data[i:j + 1]) / max(s, 1)
        print(f"sass[{i:4d}..{j:4d}] len {j - i + 1:4d} exec {n:9d} share {100 * s / tot:5.1f}%  thr {thr:4.1f}  {data[i][1][:70]}")
    i = j + 1
