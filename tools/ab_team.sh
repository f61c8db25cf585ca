cp berserkerquake2_b200/libbq2gpu.so /tmp/new.so
for mode in HEAD NEW0 NEW1; do
  case $mode in
    HEAD) cp gpurun_variants/libbq2gpu_HEAD.so berserkerquake2_b200/libbq2gpu.so; export BQ2_TEAM_DOUBLE=1;;
    NEW0) cp /tmp/new.so berserkerquake2_b200/libbq2gpu.so; export BQ2_TEAM_DOUBLE=0;;
    NEW1) cp /tmp/new.so berserkerquake2_b200/libbq2gpu.so; export BQ2_TEAM_DOUBLE=1;;
  esac
  if [ $mode != HEAD ]; then echo -n "$mode tests: "; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1; fi
  for i in 1 2; do
  echo -n "$mode "; timeout 200 python tools/bench_c4.py --no-ntb 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('shadow-only kernel_us %.1f frac %.3f' % (d['kernel_ms']*1e3, d['frac_of_measured_peak']))"
  echo -n "$mode "; timeout 200 python tools/bench_c4.py 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ntb kernel_us %.1f frac %.3f' % (d['kernel_ms']*1e3, d['frac_of_measured_peak']))"
  done
done
cp /tmp/new.so berserkerquake2_b200/libbq2gpu.so
