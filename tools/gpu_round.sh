#!/usr/bin/env bash
# One evidence pass on a B200 (under gpurun): GPU tests, the bench lines, C4 / brush tools, then the ncu launch list of
# the default bench command and `--set full` captures of the fused kernel on C2 and on the C3 shard.
# usage: bash tools/gpu_round.sh <tag>
tag=${1:-x}
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
for w in c2 c3shard; do
  timeout 300 python bench.py --no-cpu --workload $w > gpurun_out/bench_${w}_$tag.log 2> gpurun_out/bench_${w}_$tag.err
  tail -1 gpurun_out/bench_${w}_$tag.log | python tools/bench_line.py
done
timeout 200 python tools/bench_c4.py --no-ntb 2>&1 | tail -1 | cut -c1-400
timeout 200 python tools/bench_c4.py 2>&1 | tail -1 | cut -c1-400
timeout 200 python tools/bench_brush.py 2>&1 | tail -4
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --spinup 0"
$CMD > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo launches=$?
CMD2="python bench.py --steps 3 --warmup 3 --no-cpu --spinup 0"
ncu --set full --clock-control none --import-source on -k regex:bq2_frame_warp -s 20 -c 2 -o gpurun_out/prof_c2_$tag $CMD2 > gpurun_out/ncu_c2.log 2>&1
echo c2full=$?
CMD3="python bench.py --workload c3shard --steps 3 --warmup 3 --no-cpu --spinup 0"
$CMD3 > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bq2_frame -s 4 -c 1 -o gpurun_out/prof_c3_$tag $CMD3 > gpurun_out/ncu_c3.log 2>&1
echo c3full=$?
CMD4="python tools/bench_c4.py --no-ntb --steps 3"
ncu --set full --clock-control none --import-source on -k regex:bq2_frame_kernel -s 2 -c 1 -o gpurun_out/prof_c4_$tag $CMD4 > gpurun_out/ncu_c4.log 2>&1
echo c4full=$?
timeout 300 python bench.py > gpurun_out/bench_default_$tag.log 2> gpurun_out/bench_default_$tag.err; echo default=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.log 2> gpurun_out/bench_ref_$tag.err; echo ref=$?
tail -1 gpurun_out/bench_ref_$tag.log | cut -c1-300
