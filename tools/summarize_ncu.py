#!/usr/bin/env python
"""Summaries of ncu output for profiles/ (run here, no GPU needed).

    python tools/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/<name>.md
    python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep       > profiles/<name>.csv

`launches`: per-kernel count / total / share of the step from an `ncu --metrics gpu__time_duration.sum` launch list.
`full`: one row per profiled launch with the metrics the roofline needs (duration, DRAM bytes, L2 / tensor-pipe
utilisation, occupancy, registers) from an `ncu --set full` report.
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg",
]


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:90]


def launches(path):
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = short(r["Kernel Name"])
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + us)
        total += us
    print(f"| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t:.1f} | {100 * t / total:.1f}% |")
    print(f"| **all** | {sum(c for c, _ in agg.values())} | {total:.1f} | 100% |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in KEEP if c in idx]
    w = csv.writer(sys.stdout)
    w.writerow(["kernel"] + [f"{c} [{units[idx[c]]}]" for c in cols])
    for r in data:
        w.writerow([short(r[idx["Kernel Name"]])] + [r[idx[c]] for c in cols])


def _bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def traffic(path):
    """JSON for profiles/roofline_traffic.json: per kernel name, DRAM bytes (read + write) and duration of its launch."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    r_, w_, t_ = idx["dram__bytes_read.sum"], idx["dram__bytes_write.sum"], idx["gpu__time_duration.sum"]
    kernels = {}
    for r in data:
        name = short(r[idx["Kernel Name"]])
        b = _bytes(r[r_], units[r_]) + _bytes(r[w_], units[w_])
        if name not in kernels or b > kernels[name]["dram_bytes"]:
            kernels[name] = {"dram_bytes": b, "dram_read": _bytes(r[r_], units[r_]), "dram_write": _bytes(r[w_], units[w_]),
                             "duration": float(r[t_].replace(",", "")), "duration_unit": units[t_]}
    print(json.dumps({"source": path.split("/")[-1], "how": "ncu --set full --clock-control none, one training step, "
                      "largest launch per kernel name", "kernels": kernels}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
