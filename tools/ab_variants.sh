#!/usr/bin/env bash
# A/B on the GPU box: every library under gpurun_variants/libbq2gpu_*.so (built here with different -D switches) takes
# the place of the product library in turn: parity tests first, then the C2 bench line (and the C3 strong leg with
# "c3" as first argument).  The product library is put back at the end.
cp berserkerquake2_b200/libbq2gpu.so /tmp/product.so
for f in gpurun_variants/libbq2gpu_*.so; do
  v=${f#gpurun_variants/libbq2gpu_}; v=${v%.so}
  cp $f berserkerquake2_b200/libbq2gpu.so
  echo -n "$v tests: "; timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_golden.py tests/test_gpu_exhaustive.py -m gpu -x -q 2>&1 | tail -1
  extra="--no-c3"; [ "$1" = "c3" ] && extra=""
  for rep in 1 2; do
    timeout 300 python bench.py --no-cpu --no-traffic $extra 2>gpurun_out/ab_$v.err | tail -1 > gpurun_out/ab_${v}_$rep.log
    echo -n "$v "; python tools/bench_line.py < gpurun_out/ab_${v}_$rep.log | head -2 | cut -c1-220
    python tools/bench_line.py < gpurun_out/ab_${v}_$rep.log | grep c3_strong | cut -c1-200
  done
done
cp /tmp/product.so berserkerquake2_b200/libbq2gpu.so
