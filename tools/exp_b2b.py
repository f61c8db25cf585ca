import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import berserkerquake2_b200 as bq2
from berserkerquake2_b200 import synth, host
import bench
n_ents, lights = 1024, 4
ctx = bq2.Context(0)
nb = None; mids = []
for e in range(n_ents):
    blob = synth.md2(16, 17, 198, e)
    nb = host.md2_neighbours(blob) if nb is None else nb
    mids.append(ctx.upload_md2(blob, nb))
we, wl, pe, pl = bench.make_scene(n_ents, lights)
we["model"] = np.array(mids, np.int32)
ents, pairs = ctx.build_records(we, wl, pe, pl)
ctx.frame(ents, pairs)
stream = torch.cuda.ExternalStream(ctx.stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
for K in (1, 2, 4, 16, 200):
    ts = []
    for rep in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            flush.zero_(); a.record(stream)
            for _ in range(K): ctx.replay()
            b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / K)
    print("back-to-back K=%d: %.2f us per launch" % (K, np.median(ts) * 1e3), flush=True)
