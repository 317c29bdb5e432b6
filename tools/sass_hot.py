"""Summarise an `ncu --page source --csv --print-source sass` dump: executed warp instructions by opcode and the
hottest contiguous regions.  Usage: python tools/sass_hot.py dump.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ia = hdr.index("Source"); ie = hdr.index("Instructions Executed"); it = hdr.index("Avg. Threads Executed"); iss = hdr.index("# Samples")
ops = collections.Counter(); samp = collections.Counter(); tot = 0; data = []
for r in rows[2:]:
    if len(r) <= ie: continue
    try: n = int(r[ie]); s = int(r[iss])
    except ValueError: continue
    src = r[ia].strip()
    tok = src.split()
    op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")
    op = op.split(".")[0]
    ops[op] += n; samp[op] += s; tot += n
    data.append((n, s, src, r[it]))
print("total warp insts", tot, "samples", sum(samp.values()))
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{op:12s} {n:12d} {100*n/tot:5.1f}%   samples {100*samp[op]/max(sum(samp.values()),1):5.1f}%")
