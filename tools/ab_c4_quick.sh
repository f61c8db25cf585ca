#!/usr/bin/env bash
# quick A/B on the team kernel: the product library against gpurun_variants/libbq2gpu_*.so (BQ2_LIB), C4 only
for lib in berserkerquake2_b200/libbq2gpu.so gpurun_variants/libbq2gpu_*.so berserkerquake2_b200/libbq2gpu.so; do
  export BQ2_LIB=$lib
  echo "== $lib"
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exhaustive.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -1
  for a in --no-ntb --no-ntb ""; do
    timeout 200 python tools/bench_c4.py $a 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  %-12s alone %.1f  one stream %.1f  two streams %.1f us  frac %.3f' % ('$a' or 'ntb', d['kernel_ms']*1e3, d['back_to_back_ms_one_stream']*1e3, d['back_to_back_ms_two_streams']*1e3, d['frac_back_to_back_two_streams']))"
  done
done
