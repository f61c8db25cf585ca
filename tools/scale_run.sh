#!/bin/bash
# bench.py under torchrun at N GPUs (as the driver launches it); prints the JSON line's key numbers
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --no-cpu > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
tail -1 gpurun_out/bench_n$N.log | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print('N=%d value %.3e ms %.4f e2e %.3e e2e_ms %.4f frac %.3f' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))"
