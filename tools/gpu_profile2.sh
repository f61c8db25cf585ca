#!/usr/bin/env bash
# ncu --set full captures of the frame kernels (C2 warp kernel, C4 team kernel, brush kernel) + the e2e host trace.
tag=${1:-x}
mkdir -p gpurun_out
timeout 200 python tools/e2e_trace.py 2>&1 | tail -14
CMD2="python bench.py --steps 3 --warmup 3 --no-cpu --spinup 0 --no-c3 --no-traffic --no-dropin --repeats 1"
$CMD2 > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bq2_frame_warp -s 20 -c 2 -o gpurun_out/prof_c2_$tag -f $CMD2 > gpurun_out/ncu_c2.log 2>&1
echo c2full=$?
CMD4="python tools/bench_c4.py --no-ntb --steps 3"
$CMD4 > gpurun_out/plain_c4.log 2>&1; tail -1 gpurun_out/plain_c4.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:bq2_frame_kernel -s 2 -c 1 -o gpurun_out/prof_c4_$tag -f $CMD4 > gpurun_out/ncu_c4.log 2>&1
echo c4full=$?
CMD5="python tools/bench_brush.py"
$CMD5 > gpurun_out/plain_brush.log 2>&1; tail -4 gpurun_out/plain_brush.log
ncu --set full --clock-control none --import-source on -k regex:bq2_brush -s 2 -c 1 -o gpurun_out/prof_brush_$tag -f $CMD5 > gpurun_out/ncu_brush.log 2>&1
echo brushfull=$?
