#!/usr/bin/env bash
# warm-cache durations of the kernels of an end-to-end frame (ncu without its cache flush), and a --set full capture of
# the record-building kernel
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --spinup 0 --no-c3 --no-traffic --no-dropin --repeats 3"
$CMD > gpurun_out/plain_link.log 2>&1; echo plain=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k "regex:bq2_world_link|bq2_frame_tail|bq2_frame_warp" -c 1500 --csv --log-file gpurun_out/launches_warm.csv $CMD > gpurun_out/ncu_warm.log 2>&1
echo warm=$?
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:bq2_world_link -s 60 -c 1 -o gpurun_out/prof_link -f $CMD > gpurun_out/ncu_link.log 2>&1
echo link=$?
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
