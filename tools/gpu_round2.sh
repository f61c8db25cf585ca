#!/usr/bin/env bash
# Round 2 evidence pass on one B200 (under gpurun): GPU tests, the default bench line (with the in-run DRAM traffic
# probe and the C3 strong leg), the reference arm, the PDL timeline.  usage: bash tools/gpu_round2.sh <tag> [quick]
tag=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_$tag.log 2>&1; tail -4 gpurun_out/pytest_$tag.log
( time timeout 600 python bench.py ) > gpurun_out/bench_default_$tag.log 2> gpurun_out/bench_default_$tag.err; echo default=$?
tail -1 gpurun_out/bench_default_$tag.log | python tools/bench_line.py
tail -5 gpurun_out/bench_default_$tag.err
ls -la gpurun_out/traffic_* 2>/dev/null
if [ -f gpurun_variants/libbq2gpu_tl.so ]; then
  BQ2_LIB=gpurun_variants/libbq2gpu_tl.so timeout 200 python tools/timeline.py --out gpurun_out/timeline_$tag.md > gpurun_out/timeline_$tag.log 2>&1; echo timeline=$?
  tail -25 gpurun_out/timeline_$tag.log
  BQ2_LIB=gpurun_variants/libbq2gpu_tl.so timeout 200 python tools/timeline.py --streams 2 --out gpurun_out/timeline2_$tag.md > gpurun_out/timeline2_$tag.log 2>&1; echo timeline2=$?
  tail -22 gpurun_out/timeline2_$tag.log
fi
if [ "$2" != "quick" ]; then
  timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$tag.log 2> gpurun_out/bench_ref_$tag.err; echo ref=$?
  tail -1 gpurun_out/bench_ref_$tag.log | cut -c1-400
fi
