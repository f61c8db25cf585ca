#!/usr/bin/env bash
# A/B of library variants on the brush kernel: parity tests, then the kernel's warm duration at 256 and 4,096 pairs (ncu
# without cache flush; bench_brush's own figure is the end-to-end call)
cp berserkerquake2_b200/libbq2gpu.so /tmp/product.so
run() {
  echo -n "$1 tests: "; timeout 300 python -m pytest tests/test_gpu_brush.py -m gpu -x -q 2>&1 | tail -1
  timeout 200 python tools/bench_brush.py 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$1 e2e: %5d pairs %.1f us per call' % (d['pairs'], d['gpu_ms_per_call_e2e'] * 1e3))"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:bq2_brush --csv --log-file gpurun_out/brush_$1.csv python tools/bench_brush.py > /dev/null 2>&1
  python - <<PY
import csv
rows = list(csv.reader(l for l in open("gpurun_out/brush_$1.csv") if l.startswith('"')))
h = rows[0]; v = h.index("Metric Value")
x = [float(r[v].replace(",", "")) / 1e3 for r in rows[1:]]
a, b = sorted(x[3:24]), sorted(x[27:48])
print("$1 kernel: 256 pairs %.2f us   4,096 pairs %.2f us (medians of %d / %d launches)" % (a[len(a)//2], b[len(b)//2], len(a), len(b)))
PY
}
run base
for f in gpurun_variants/libbq2gpu_*.so; do
  v=${f#gpurun_variants/libbq2gpu_}; v=${v%.so}
  cp $f berserkerquake2_b200/libbq2gpu.so
  run $v
done
cp /tmp/product.so berserkerquake2_b200/libbq2gpu.so
