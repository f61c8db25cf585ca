#!/usr/bin/env bash
# One GPU-box cycle: parity tests, C2 + C3-shard bench lines, and (optionally) one ncu capture of the top kernel.
# usage (under gpurun): bash tools/gpu_cycle.sh <tag> [ncu]
tag=${1:-x}
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/bench_c2_$tag.log 2>&1; echo bench_c2=$?
timeout 300 python bench.py --workload c3shard --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c3_$tag.log 2>&1; echo bench_c3=$?
for f in gpurun_out/bench_c2_$tag.log gpurun_out/bench_c3_$tag.log; do
  tail -1 $f | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step %.4f  frac %.3f  frac_min %.3f  e2e_ms %.4f  tris/s %.3e  clocks %s' % (d['ms_per_step'], r['frac'], r['frac_min_traffic'], d['e2e']['ms_per_step'], d['value'], d['clocks']['sm_mhz']))" || tail -5 $f
done
if [ "$2" = "ncu" ]; then
  python bench.py --steps 3 --warmup 3 --no-cpu --spinup 0 > gpurun_out/plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:bq2_frame -s 4 -c 2 -o gpurun_out/prof_$tag \
      python bench.py --steps 3 --warmup 3 --no-cpu --spinup 0 > gpurun_out/ncu_$tag.log 2>&1; echo ncu=$?
fi
