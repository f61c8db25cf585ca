#!/usr/bin/env bash
# diagnostic variants (results are WRONG on purpose: stores suppressed): timing only, no tests
cp berserkerquake2_b200/libbq2gpu.so /tmp/product.so
for f in gpurun_variants/libbq2gpu_*.so; do
  v=${f#gpurun_variants/libbq2gpu_}; v=${v%.so}
  cp $f berserkerquake2_b200/libbq2gpu.so
  timeout 300 python bench.py --no-cpu --no-traffic --no-dropin 2>gpurun_out/ab_$v.err | tail -1 > gpurun_out/ab_${v}_1.log
  echo -n "$v "; python tools/bench_line.py < gpurun_out/ab_${v}_1.log 2>/dev/null | grep -E "step_us|c3_strong" | cut -c1-200
done
cp /tmp/product.so berserkerquake2_b200/libbq2gpu.so
