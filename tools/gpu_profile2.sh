#!/usr/bin/env bash
# Round-2 profile pass on one B200 (under gpurun): the ncu launch list of the bench command and --set full captures of
# both instantiations of the warp kernel that the bench line uses (<0,0>: the device-timed region, host-built records;
# <1,0>: the end-to-end frames, device-built records), each only after the same command has run without ncu.
tag=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --spinup 0 --no-c3 --no-traffic --no-dropin --repeats 3"
$CMD > gpurun_out/plain_launches.log 2>&1; echo plain=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo launches=$?
ncu --set full --clock-control none --import-source on -k regex:bq2_frame_warp_kernel -s 30 -c 1 -o gpurun_out/prof_c2_00_$tag -f $CMD > gpurun_out/ncu_c2_00.log 2>&1
echo c2_00=$?
ncu --set full --clock-control none --import-source on -k "regex:bq2_frame_warp_kernel<1, 0>|bq2_frame_warp_kernel<\(bool\)1, \(bool\)0>" -s 40 -c 1 -o gpurun_out/prof_c2_10_$tag -f $CMD > gpurun_out/ncu_c2_10.log 2>&1
echo c2_10=$?
ls -la gpurun_out/prof_c2_*_$tag.ncu-rep gpurun_out/launches_$tag.csv
