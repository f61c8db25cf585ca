#!/usr/bin/env bash
# C2 bench lines of several builds of the library, alternating on ONE box (a fresh box's first run is the slow one,
# so every build is run twice, interleaved).  Builds: "cur" = the in-tree product library, any other name N =
# gpurun_variants/libbq2gpu_N.so (made here with `make lib EXTRA=... BUILD=build_N OUT=../../gpurun_variants/libbq2gpu_N.so`).
# A build may carry environment for bench.py after a colon: fl7:BQ2_BENCH_IN_FLIGHT=7
# usage (under gpurun): bash tools/ab_order.sh [-w workload] cur old fl7:BQ2_BENCH_IN_FLIGHT=7
w=c2; if [ "$1" = "-w" ]; then w=$2; shift 2; fi
cp berserkerquake2_b200/libbq2gpu.so /tmp/cur.so
for round in 1 2; do
  for spec in "$@"; do
    v=${spec%%:*}; envs=""; [ "$spec" != "$v" ] && envs=${spec#*:}
    if [ $v = cur ]; then cp /tmp/cur.so berserkerquake2_b200/libbq2gpu.so; else cp gpurun_variants/libbq2gpu_$v.so berserkerquake2_b200/libbq2gpu.so; fi
    env $envs timeout 60 python bench.py --no-cpu --workload $w > gpurun_out/ab_${v}_$round.log 2>/dev/null
    echo -n "$v #$round t=$SECONDS "; tail -1 gpurun_out/ab_${v}_$round.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); e=d['e2e']
    print('step_us %.2f iso_us %.2f e2e_us %.2f min/max %s py %.2f serial %.1f' % (d['ms_per_step']*1e3, d['roofline']['isolated_launch_ms']['l2_flushed_by_memset']*1e3, e['ms_per_step']*1e3, [round(x*1e3,2) for x in e['ms_per_step_min_max_rank0']], e['ms_per_step_python_driver']*1e3, e['ms_per_step_serial']*1e3))
except Exception as ex: print('no bench line:', ex)"
  done
done
cp /tmp/cur.so berserkerquake2_b200/libbq2gpu.so
