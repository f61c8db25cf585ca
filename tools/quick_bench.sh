#!/bin/bash
# quick A/B on the GPU box: parity first, then kernel time at C2 and at the C3 shard (three runs each)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -2
for w in c2 c2 c2 c3shard c3shard; do
  python bench.py --no-cpu --workload $w 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w', 'kernel_us %.2f' % (d['ms_per_step']*1e3), 'clean_us %.2f' % (d['roofline']['kernel_ms_l2_clean_lines']*1e3), 'e2e_us %.2f' % (d['e2e']['ms_per_step']*1e3))"
done
