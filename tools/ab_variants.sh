#!/usr/bin/env bash
# A/B on the GPU box: every library under gpurun_variants/ (built here with different -D switches, see the loop in
# profiles/r01_history.md) takes the place of the product library in turn: parity tests first, then the C2 and
# C3-shard bench lines.  The product library is put back at the end.
cp berserkerquake2_b200/libbq2gpu.so /tmp/product.so
for f in gpurun_variants/libbq2gpu_*.so; do
  v=${f#gpurun_variants/libbq2gpu_}; v=${v%.so}
  cp $f berserkerquake2_b200/libbq2gpu.so
  echo -n "$v tests: "; timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -1
  for w in c2 c3shard; do
    echo -n "$v "; timeout 200 python bench.py --no-cpu --workload $w 2>/dev/null | tail -1 | python tools/bench_line.py
  done
done
cp /tmp/product.so berserkerquake2_b200/libbq2gpu.so
