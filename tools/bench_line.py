#!/usr/bin/env python
"""Short summary of a bench.py JSON line read from stdin (used by the GPU-box scripts)."""
import json
import sys

for line in sys.stdin:
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d["roofline"]
    iso = r.get("isolated_launch_ms", {})
    print("%s  step_us %.2f  iso_us %.2f  frac(unique) %.3f  frac(8d) %.3f  e2e_us %.2f  tris/s %.3e  setup_s %.1f" % (
        d["config"]["workload"][:28], d["ms_per_step"] * 1e3, iso.get("l2_flushed_by_memset", 0) * 1e3, r["frac"],
        r["survey_8d_per_pair_form"]["frac"], d["e2e"]["ms_per_step"] * 1e3, d["value"], d.get("setup_s", 0)))
    m = d["measurement"]
    print("  regions_ms", [round(x, 4) for x in m["timed_regions_ms"]], "streams", m.get("streams"), "one-stream us/step", round(m.get("ms_per_step_one_stream", 0) * 1e3, 2),
          "long region (K=%s) us/step %.2f frac %.3f" % ((r.get("long_region") or {}).get("steps"), (r.get("long_region") or {}).get("kernel_ms", 0) * 1e3,
                                                         (r.get("long_region") or {}).get("frac", 0)))
    print("  traffic/step", r.get("traffic"), "unique bytes", r.get("bytes_per_launch"), r.get("traffic_unmeasured", ""),
          (r.get("traffic_detail") or {}).get("how", ""))
    e = d["e2e"]
    print("  e2e:", {k: round(v, 5) if isinstance(v, float) else v for k, v in e.items() if k != "what"})
    c3 = d.get("c3_strong")
    if c3:
        print("  c3_strong: n_gpus %d  step_us %.1f  tris/s %.3e  pairs/s %.3e  frac(unique) %.3f  setup_s %.1f  regions %s" % (
            c3["n_gpus"], c3["ms_per_step"] * 1e3, c3["value"], c3["pairs_per_s"], c3["roofline"]["frac"], c3["setup_s"],
            [round(x, 3) for x in c3["timed_regions_ms"]]))
    cb = d.get("cpu_baseline")
    if cb:
        print("  cpu: %.3e tris/s on %d cores, ms/step %.3f %s" % (cb["value"], cb["cores"], cb["ms_per_step"], cb.get("ms_per_step_min_max")))
    print("  clocks", d.get("clocks"))
