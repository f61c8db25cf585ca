#!/usr/bin/env bash
# team kernel (C4) and brush kernel: parity tests of both, then their bench tools
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_brush.py tests/test_gpu_exhaustive.py tests/test_golden.py tests/test_gpu_fuzz.py tests/test_gpu_threads.py tests/test_gpu_soak.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do timeout 200 python tools/bench_c4.py --no-ntb 2>&1 | tail -1 | cut -c1-330; done
timeout 200 python tools/bench_c4.py 2>&1 | tail -1 | cut -c1-330
timeout 200 python tools/bench_brush.py 2>&1 | tail -2 | cut -c1-260
