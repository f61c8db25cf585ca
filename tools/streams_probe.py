#!/usr/bin/env python
"""C2 (and a C3 shard): K back-to-back steps on 1 stream (programmatic dependent launch) vs dealt onto 2 / 3 / 4 streams."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import berserkerquake2_b200 as bq2
from berserkerquake2_b200 import synth
import bench
for name, (n_ents, lights) in (("C2", (1024, 4)), ("C3 shard", (8192, 8))):
    ctx = bq2.Context(0)
    V, T = synth.sphere_dims(bench.SLICES, bench.RINGS)
    l2 = int(getattr(torch.cuda.get_device_properties(0), "L2_cache_size", 126 << 20))
    ring, _ = bench.ring_size(l2, n_ents, lights, V, T, cap=17 if n_ents == 1024 else 4)
    R = bench.build_ring(ctx, n_ents, lights, 0, n_ents, max(ring, 2), own_results=True)
    K = 100
    ctx.run_ring(R["kept"], 0, 60)
    for S in (1, 2, 3, 4, 8):
        ms = [ctx.run_ring(R["kept"], (i * K) % len(R["kept"]), K, timed=True, streams=S) for i in range(9)]
        print(f"{name}: {S} stream(s): {np.median(ms) / K * 1e3:.2f} us per step (min {min(ms) / K * 1e3:.2f})", flush=True)
    # results intact after the multi-stream runs
    for h in R["kept"]:
        ctx.free_kept(h)
    ctx.close()
