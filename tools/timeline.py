#!/usr/bin/env python
"""Phase timeline of the warp kernel on C2 (needs `make -C berserkerquake2_b200/csrc TIMELINE=1`): warp 0 of every
CTA stamps %globaltimer at: 0 start, 1 TMA issued, 2 key-frames arrived, 3 lerp done (CTA barrier), 4 normals done
(CTA barrier), 5 facing done, 6 sides done, 7 caps done (first light of warp 0)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import berserkerquake2_b200 as bq2
from berserkerquake2_b200 import synth, host, _capi as c
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
for _ in range(5):
    with torch.cuda.stream(stream):
        flush.zero_(); ctx.replay()
torch.cuda.synchronize()
buf = np.zeros(n_ents * 8, np.uint64)
assert c.lib.bq2_debug_timeline(buf.ctypes.data_as(C.c_void_p), buf.size) == 0
t = buf.reshape(n_ents, 8).astype(np.int64)
t0 = t[:, 0].min()
rel = (t - t0) / 1000.0
names = ["start", "TMA issued", "frames arrived", "lerp done", "normals done", "facing done", "sides done", "caps done(light 0)"]
print("phase                 median us   p10     p90     max   (since the first CTA started)")
for i, n in enumerate(names):
    print(f"{n:20s} {np.median(rel[:, i]):8.2f} {np.percentile(rel[:, i], 10):7.2f} {np.percentile(rel[:, i], 90):7.2f} {rel[:, i].max():7.2f}")
d = np.diff(rel, axis=1)
print("per-CTA phase durations (median us):", " ".join(f"{names[i+1]}={np.median(d[:, i]):.2f}" for i in range(7)))
