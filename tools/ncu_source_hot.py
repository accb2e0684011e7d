#!/usr/bin/env python
"""Per-source-line aggregation of an `ncu --page source --csv --print-source cuda,sass` dump:
samples (stall sampling) and executed warp instructions per CUDA source line, top N.
    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --launch-skip K --launch-count 1 > dump.csv
    python tools/ncu_source_hot.py dump.csv [N]"""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rd = csv.reader(open(path))
    cur_file, hdr = None, None
    agg = defaultdict(lambda: [0, 0, ""])   # (file, line) -> [samples, inst, src]
    stall = defaultdict(lambda: defaultdict(int))
    for r in rd:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        if r[0] not in ("", "-"):          # a CUDA source line header: totals for the line
            try:
                key = (cur_file, int(r[0]))
            except ValueError:
                continue
            a = agg[key]
            a[2] = r[1]
            try:
                a[0] += int(r[i_s]); a[1] += int(r[i_i])
                for i, h in stall_cols:
                    stall[key][h] += int(r[i] or 0)
            except ValueError:
                pass
    ts = sum(a[0] for a in agg.values()) or 1
    ti = sum(a[1] for a in agg.values()) or 1
    print(f"total samples {ts}, warp instructions {ti}")
    for (f, ln), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        st = sorted(stall[(f, ln)].items(), key=lambda kv: -kv[1])[:3]
        sts = " ".join(f"{h[6:]}:{v}" for h, v in st if v)
        print(f"{f}:{ln:<4d} samp {100 * s / ts:5.1f}% inst {100 * i / ti:5.1f}% [{sts}] {src.strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
