#!/usr/bin/env bash
# after a change to the world frame's chain of stream operations: the whole GPU suite, then the e2e legs (with and
# without programmatic dependent launches inside the frame's chain) and the warm-cache durations of a frame's kernels
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for pdl in 0 1 0 1; do
  export BQ2_PDL_WORLD=$pdl
  timeout 300 python bench.py --no-cpu --no-traffic --no-c3 --no-dropin 2>gpurun_out/tail_$pdl.err | tail -1 > gpurun_out/tail_$pdl.log
  echo "PDL=$pdl"; python - <<PY
import json
d = json.loads(open("gpurun_out/tail_$pdl.log").read())
e = d["e2e"]
print("  step_us %.2f  e2e_us %.2f  serial %.2f  2/4/8 in flight %.2f %.2f %.2f  launches/step %s" % (d["ms_per_step"]*1e3, e["ms_per_step"]*1e3, e["ms_per_step_serial"]*1e3, e["ms_per_step_2_in_flight"]*1e3, e["ms_per_step_4_in_flight"]*1e3, e["ms_per_step_8_in_flight"]*1e3, e["gpu_launches_per_step"]))
PY
done
unset BQ2_PDL_WORLD
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --spinup 0 --no-c3 --no-traffic --no-dropin --repeats 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k "regex:bq2_world_link|bq2_frame_tail|bq2_frame_warp" -c 1500 --csv --log-file gpurun_out/launches_warm.csv $CMD > gpurun_out/ncu_warm.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/launches_warm.csv") if l.startswith('"')))
h = rows[0]; k, v = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[k][:60]].append(float(r[v].replace(",", "")))
for n, x in d.items():
    x = sorted(x); print(f"{n:62s} n={len(x):4d} median {x[len(x)//2]/1e3:7.2f} us  min {x[0]/1e3:.2f} max {x[-1]/1e3:.2f}")
PY
