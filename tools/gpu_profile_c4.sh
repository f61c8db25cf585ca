#!/usr/bin/env bash
# --set full captures of the team kernel on C4 (shadow only, and with the tangent-space outputs) and of the brush
# kernel, each after the same command has run without ncu.
tag=${1:-x}
mkdir -p gpurun_out
timeout 200 python tools/bench_c4.py --no-ntb --steps 10 2>&1 | tail -1 | cut -c1-330
timeout 200 python tools/bench_c4.py --steps 10 2>&1 | tail -1 | cut -c1-330
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bq2_frame_kernel -s 6 -c 1 -o gpurun_out/prof_c4_shadow_$tag -f python tools/bench_c4.py --no-ntb --steps 4 > gpurun_out/ncu_c4_shadow.log 2>&1
echo c4_shadow=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bq2_frame_kernel -s 6 -c 1 -o gpurun_out/prof_c4_ntb_$tag -f python tools/bench_c4.py --steps 4 > gpurun_out/ncu_c4_ntb.log 2>&1
echo c4_ntb=$?
timeout 200 python tools/bench_brush.py 2>&1 | tail -2 | cut -c1-260
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bq2_brush -s 26 -c 1 -o gpurun_out/prof_brush_$tag -f python tools/bench_brush.py > gpurun_out/ncu_brush.log 2>&1
echo brush=$?
ls -la gpurun_out/prof_c4_*_$tag.ncu-rep gpurun_out/prof_brush_$tag.ncu-rep
