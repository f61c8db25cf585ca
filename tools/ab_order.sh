#!/usr/bin/env bash
# old / new / old / new C2 bench lines on one box (run order matters on a fresh box: its first run is the slow one)
cp berserkerquake2_b200/libbq2gpu.so /tmp/new.so
for v in old new old new; do
  if [ $v = old ]; then cp gpurun_variants/libbq2gpu_old.so berserkerquake2_b200/libbq2gpu.so; else cp /tmp/new.so berserkerquake2_b200/libbq2gpu.so; fi
  [ $SECONDS -gt 46 ] && break
  timeout 14 python bench.py --no-cpu > gpurun_out/bench_$v$SECONDS.log 2>/dev/null
  echo -n "$v t=$SECONDS "; tail -1 gpurun_out/bench_$v*.log | grep "^{" | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('e2e_us %.2f min/max %s py %.2f serial %.1f hostbuilt %.1f' % (e['ms_per_step']*1e3, [round(x*1e3,2) for x in e['ms_per_step_min_max_rank0']], e['ms_per_step_python_driver']*1e3, e['ms_per_step_serial']*1e3, e['ms_per_step_host_built_records']*1e3))"
done
cp /tmp/new.so berserkerquake2_b200/libbq2gpu.so
