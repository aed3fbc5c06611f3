"""Aggregates an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line: share of executed warp
instructions and of stall samples.  usage: python profiles/src_hot.py dump.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
h = next(r for r in rows if 'Instructions Executed' in r)
ie, ss, ln, src = h.index('Instructions Executed'), h.index('# Samples'), 0, 1
inst, smp, text = collections.Counter(), collections.Counter(), {}
cur = None
for r in rows:
    if len(r) <= ie or r is h:
        continue
    if r[ln].strip().isdigit():
        cur = int(r[ln]); text[cur] = r[src]
    try:
        a, b = int(r[ie] or 0), int(r[ss] or 0)
    except ValueError:
        continue
    if cur is not None and r[2]:      # a SASS row
        inst[cur] += a; smp[cur] += b
ti, ts = sum(inst.values()), sum(smp.values())
print(f"total warp instructions {ti}, samples {ts}")
for l, v in inst.most_common(top):
    print(f"{l:5d} {100 * v / ti:5.1f}% inst {100 * smp[l] / max(ts, 1):5.1f}% smp  {text[l].strip()[:120]}")
