#!/usr/bin/env bash
# Evidence pass on one B200 (under gpurun): the ncu launch list of the default bench command and one
# `--set full` capture of the C3-shard launch (output > L2, so dram__bytes is the real traffic).
tag=${1:-x}
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --spinup 0"
$CMD > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo launches=$?
CMD3="python bench.py --workload c3shard --steps 3 --warmup 3 --no-cpu --spinup 0"
$CMD3 > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bq2_frame -s 4 -c 1 -o gpurun_out/prof_c3_$tag $CMD3 > gpurun_out/ncu_c3.log 2>&1
echo c3full=$?
