#!/bin/bash
# bench.py under torchrun at N GPUs exactly as the driver launches it (--steps 20 --warmup 5), both arms; then the
# config-5 sweep at N.  usage: bash tools/scale_run.sh N <tag> [nosweep]
N=${1:-2}; tag=${2:-x}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_ref_n${N}_$tag.log 2> gpurun_out/bench_ref_n${N}_$tag.err
tail -1 gpurun_out/bench_ref_n${N}_$tag.log | cut -c1-330
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_n${N}_$tag.log 2> gpurun_out/bench_n${N}_$tag.err
grep '^{' gpurun_out/bench_n${N}_$tag.log | tail -1 | python tools/bench_line.py
tail -4 gpurun_out/bench_n${N}_$tag.err
if [ "$3" != "nosweep" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
      tools/sweep.py --out gpurun_out/sweep_${N}gpu_$tag.md > gpurun_out/sweep_${N}gpu_$tag.log 2>&1
  cat gpurun_out/sweep_${N}gpu_$tag.md
fi
