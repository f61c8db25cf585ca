#!/usr/bin/env bash
# End-of-round evidence pass on one B200 (under gpurun): GPU tests, smoke, the bench line with the driver's flags and
# with the defaults, the reference arm, the ncu launch list of the bench command, --set full captures of the kernels
# changed last (record-building kernel, brush kernel), the config-5 sweep.  usage: bash tools/gpu_final.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_$tag.log 2>&1; tail -4 gpurun_out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$tag.log 2> gpurun_out/bench_ref_$tag.err; echo ref=$?
tail -1 gpurun_out/bench_ref_$tag.log | cut -c1-300
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_driver_$tag.log 2> gpurun_out/bench_driver_$tag.err; echo driver_flags=$?
tail -1 gpurun_out/bench_driver_$tag.log | python tools/bench_line.py
( time timeout 600 python bench.py ) > gpurun_out/bench_default_$tag.log 2> gpurun_out/bench_default_$tag.err; echo default=$?
tail -1 gpurun_out/bench_default_$tag.log | python tools/bench_line.py
cp gpurun_out/traffic_c2_range.csv gpurun_out/traffic_c2_range_$tag.csv 2>/dev/null
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --spinup 0 --no-c3 --no-traffic --no-dropin --repeats 3"
$CMD > gpurun_out/plain_launches.log 2>&1; echo plain=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo launches=$?
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:bq2_world_link -s 60 -c 1 -o gpurun_out/prof_link_$tag -f $CMD > gpurun_out/ncu_link.log 2>&1
echo link=$?
timeout 200 python tools/bench_brush.py 2>&1 | tail -2 | cut -c1-420
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bq2_brush -s 52 -c 1 -o gpurun_out/prof_brush_$tag -f python tools/bench_brush.py > gpurun_out/ncu_brush.log 2>&1
echo brush=$?
timeout 200 python tools/serial_timeline.py 2>&1 | grep serial
timeout 900 python tools/sweep.py --out gpurun_out/sweep_1gpu_$tag.md > gpurun_out/sweep_1gpu_$tag.log 2>&1; echo sweep=$?
cat gpurun_out/sweep_1gpu_$tag.md | cut -c1-260
