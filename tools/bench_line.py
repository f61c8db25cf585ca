#!/usr/bin/env python
"""One-line summary of a bench.py JSON line read from stdin (used by the GPU-box scripts)."""
import json
import sys

for line in sys.stdin:
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d["roofline"]
    iso = r.get("isolated_launch_ms", {})
    print("%s  step_us %.2f  iso_us %.2f  frac %.3f  frac_min %.3f  e2e_us %.2f  tris/s %.3e" % (
        d["config"]["workload"][:28], d["ms_per_step"] * 1e3, iso.get("l2_flushed_by_memset", 0) * 1e3, r["frac"],
        r["frac_min_traffic"], d["e2e"]["ms_per_step"] * 1e3, d["value"]))
