#!/usr/bin/env python
"""Human-readable digest of a bench.py JSON line (stdout of a gpurun job)."""
import json
import sys

d = None
for line in open(sys.argv[1]):
    line = line.strip()
    if line.startswith("{"):
        try:
            d = json.loads(line)
        except Exception:
            pass
if d is None:
    print("no JSON line in", sys.argv[1])
    sys.exit(0)
c = d.get("config", {})
print(f"N={d.get('n_gpus')} value {d.get('value', 0):.0f} {d.get('unit')}  {d.get('ms_per_step', 0):.3f} ms/step  "
      f"e2e {d.get('e2e', {}).get('value', 0):.0f} ({d.get('e2e', {}).get('ms_per_step', 0):.3f} ms)  spr {c.get('samples_per_ray')} "
      f"alive {c.get('alive_samples_per_ray')} refreshes {c.get('refreshes_in_timed_region')} launches {d.get('gpu_launches')}")
print("exchange:", c.get("grad_exchange"), "| host enqueue ms", d.get("host_enqueue_ms_per_step"), "| clocks", d.get("clocks"))
print("early_termination:", d.get("early_termination"))
print("phases:", d.get("phases_ms"))
r = d.get("roofline") or {}
print("roofline:", {k: v for k, v in r.items() if k != "all"})
if r.get("all"):
    print({k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in r["all"].items()})
print("render:", d.get("render"))
print("c5:", d.get("c5"))
print("cpu:", d.get("cpu_baseline"))
