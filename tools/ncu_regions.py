#!/usr/bin/env python
"""Stall-sample distribution of one kernel of an `ncu --set full --import-source on` report, cut into regions at
landmark SASS instructions (MMA issue, TMEM loads, barriers, global / shared accesses).
    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_regions.py src.csv <kernel index> [min share %]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
k = int(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
hdr = rows[start[k] + 1]
end = start[k + 1] if k + 1 < len(start) else len(rows)
data = rows[start[k] + 2:end]
idx = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[idx["# Samples"]]) for r in data)
print(rows[start[k]][1][:90], "samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
marks = re.compile(r"UTCHMMA|SYNCS|BAR\.|UTCBAR|STG|LDTM|LDG|EXIT|LDGSTS|LDGDEPBAR|DEPBAR|RED|ATOM|SHFL|STS|LDS")
cum = last = 0
agg = {h: 0 for h in stalls}
for i, r in enumerate(data):
    cum += int(r[idx["# Samples"]])
    for h in stalls:
        agg[h] += int(r[idx[h]])
    op = r[idx["Source"]].strip()
    if marks.search(op) and cum - last > thr / 100 * tot:
        top = sorted(agg.items(), key=lambda kv: -kv[1])[:2]
        print(f"{i:5d} {100 * cum / tot:6.1f}% (+{100 * (cum - last) / tot:5.1f})  {op[:58]:58s} {[(h[6:], v) for h, v in top]}")
        last = cum
        agg = {h: 0 for h in stalls}
