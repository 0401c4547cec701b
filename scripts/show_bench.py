#!/usr/bin/env python
"""Print the interesting fields of a bench.py JSON line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
if d.get("impl") == "reference":
    print("reference arm: %.4e %s  (%s cores) %s" % (
        d["value"], d["unit"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["sample"]))
    sys.exit(0)
r = d.get("roofline") or {}
print("value %.4e %s | ms/step %.2f | e2e %s | launches %s | acc %.3f step %.4f" % (
    d["value"], d["unit"], d["ms_per_step"],
    ("%.4e" % d["e2e"]["value"]) if d.get("e2e") else None, d["gpu_launches"],
    d["acceptance_mean"], d["step_size"]))
if r:
    print("roofline: %s kernel_ms %.4f share %.3f | hbm %.0f GB/s frac %.3f | tensor %.1f TF frac %.3f" % (
        r["kernel"], r["kernel_ms_per_launch"], r["kernel_share_of_step"], r["achieved"],
        r["frac"], r["tensor"]["achieved"], r["tensor"]["frac"]))
print("clocks", d.get("clocks"))
print("cpu_baseline", d.get("cpu_baseline"))
