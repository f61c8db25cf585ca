#!/usr/bin/env bash
# One short GPU-box call: parity of the in-tree build on the world path first, then its C2 bench line next to the
# previous build's (gpurun_variants/libbq2gpu_old.so) on the same box, then as much of the rest of the GPU suite as
# the time allows.  usage (under gpurun): bash tools/ab_final.sh <seconds>
budget=${1:-110}
left() { echo $(( budget - SECONDS )); }
timeout 45 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_soak.py -m gpu -x -q \
  -k "world or entity_records or ring_soak or in_flight or overlapped or degenerate or boundary" 2>&1 | tail -4
echo "t=$SECONDS"
timeout 30 python bench.py --no-cpu > gpurun_out/bench_new.log 2> gpurun_out/bench_new.err; echo "bench_new=$? t=$SECONDS"
tail -1 gpurun_out/bench_new.log | python tools/bench_line.py
if [ $(left) -gt 28 ]; then
  cp berserkerquake2_b200/libbq2gpu.so /tmp/new.so; cp gpurun_variants/libbq2gpu_old.so berserkerquake2_b200/libbq2gpu.so
  timeout 28 python bench.py --no-cpu > gpurun_out/bench_old.log 2> gpurun_out/bench_old.err; echo "bench_old=$? t=$SECONDS"
  tail -1 gpurun_out/bench_old.log | python tools/bench_line.py
  cp /tmp/new.so berserkerquake2_b200/libbq2gpu.so
fi
r=$(left)
if [ $r -gt 12 ]; then
  timeout $(( r - 4 )) python -m pytest tests -m gpu -x -q \
    -k "not ratio_sequence and not (world or entity_records or ring_soak or in_flight or overlapped or degenerate or boundary)" 2>&1 | tail -4
fi
echo "t=$SECONDS"
