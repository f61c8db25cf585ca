#!/usr/bin/env bash
# A/B of library variants (gpurun_variants/libbq2gpu_*.so) on the team kernel (C4) and the C2 bench line; the product
# library runs first and last ("base"), and is put back at the end.
cp berserkerquake2_b200/libbq2gpu.so /tmp/product.so
run() {
  echo -n "$1 tests: "; timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_golden.py tests/test_gpu_exhaustive.py -m gpu -x -q 2>&1 | tail -1
  for rep in 1 2; do
    timeout 200 python tools/bench_c4.py --no-ntb 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 c4 shadow: alone %.1f  one stream %.1f  two streams %.1f us  frac %.3f' % (d['kernel_ms']*1e3, d['back_to_back_ms_one_stream']*1e3, d['back_to_back_ms_two_streams']*1e3, d['frac_back_to_back_two_streams']))"
    timeout 200 python tools/bench_c4.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 c4 ntb:    alone %.1f  one stream %.1f  two streams %.1f us  frac %.3f' % (d['kernel_ms']*1e3, d['back_to_back_ms_one_stream']*1e3, d['back_to_back_ms_two_streams']*1e3, d['frac_back_to_back_two_streams']))"
  done
  timeout 300 python bench.py --no-cpu --no-traffic --no-c3 --no-dropin 2>gpurun_out/ab_$1.err | tail -1 > gpurun_out/ab_$1.log
  echo -n "$1 "; python tools/bench_line.py < gpurun_out/ab_$1.log | head -1 | cut -c1-200
}
run base
for f in gpurun_variants/libbq2gpu_*.so; do
  v=${f#gpurun_variants/libbq2gpu_}; v=${v%.so}
  cp $f berserkerquake2_b200/libbq2gpu.so
  run $v
done
cp /tmp/product.so berserkerquake2_b200/libbq2gpu.so
run base2
