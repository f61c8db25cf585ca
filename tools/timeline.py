#!/usr/bin/env python
"""Device-side timeline of back-to-back steps of the warp kernel on C2 -- the evidence for the programmatic-dependent-
launch overlap that ncu (which serialises launches) cannot show.  Needs the diagnostic build:

    make -C berserkerquake2_b200/csrc lib TIMELINE=1 BUILD=build_tl OUT=../../gpurun_variants/libbq2gpu_tl.so
    BQ2_LIB=gpurun_variants/libbq2gpu_tl.so python tools/timeline.py [--steps 12] [--out profiles/r02_c2_pdl_timeline.md]

Thread 0 of each of the 1,024 CTAs of every launch stamps %globaltimer (one clock for the whole GPU, 32 ns ticks on
B200 are typical) at: 0 CTA start, 1 TMA issued (entity record arrived), 2 key frames + triangle table in shared memory,
3 lerp done, 4 triangle normals done (= arrives at griddepcontrol.wait), 5 wait returned (the grid in front has completed
and flushed), 6 sides of warp 0's last light written, 7 warp 0 done.  Stamps of 16 consecutive launches are kept side by
side.  The same ring of resident frames as bench.py (own result arrays), launched from C by bq2_resident_run."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                            # noqa: E402,F401  (device properties)
import berserkerquake2_b200 as bq2                      # noqa: E402
from berserkerquake2_b200 import synth, _capi as c      # noqa: E402
import bench                                            # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--out", default=None)
ap.add_argument("--streams", type=int, default=1)
a = ap.parse_args()
assert hasattr(c.lib, "bq2_debug_timeline"), "not a TIMELINE build: set BQ2_LIB to the diagnostic library"
n_ents, lights = 1024, 4
ctx = bq2.Context(0)
V, T = synth.sphere_dims(bench.SLICES, bench.RINGS)
l2 = int(getattr(torch.cuda.get_device_properties(0), "L2_cache_size", 126 << 20))
ring, _ = bench.ring_size(l2, n_ents, lights, V, T)
R = bench.build_ring(ctx, n_ents, lights, 0, n_ents, ring, own_results=True)
K = min(a.steps, 14)
ctx.run_ring(R["kept"], 0, 3 * ring)                    # warm: clocks up, every frame's arrays touched
torch.cuda.synchronize()
seq0 = ctx.kernel_launches
ms = ctx.run_ring(R["kept"], 0, K, timed=True, streams=a.streams)
buf = np.zeros(16 * 1024 * 8, np.uint64)
assert c.lib.bq2_debug_timeline(buf.ctypes.data_as(C.c_void_p), buf.size) == 0
t = buf.reshape(16, 1024, 8).astype(np.int64)
rows = [t[(seq0 + i) & 15] for i in range(K)]           # launch i of the timed run
t0 = rows[0][:, 0].min()
us = lambda x: (x - t0) / 1000.0
lines = []
P = lines.append
P(f"# Round 2: back-to-back steps of `bq2_frame_warp_kernel<0,0>` on C2, device timeline (`tools/timeline.py`, TIMELINE build)")
P("")
P(f"{K} launches from C (`bq2_resident_run{'_streams' if a.streams > 1 else ''}`) on {a.streams} stream(s), with programmatic stream serialization on each, over the bench ring ({ring} resident frames, "
  f"own result arrays); CUDA events around the {K} launches: {ms * 1e3 / K:.2f} us per step.  `%globaltimer` stamps of thread 0 "
  "of all 1,024 CTAs per launch, microseconds since the first CTA of launch 0 started.")
P("")
P("| launch | first CTA starts | last CTA starts | median CTA reaches `griddepcontrol.wait` | first `wait` returns | last CTA done | "
  "CTAs started before the launch in front finished | ... that had staged + lerped + built normals before then | start-to-start |")
P("|---|---|---|---|---|---|---|---|---|")
prev_end = None
prev_start = None
for i, r in enumerate(rows):
    s0, s1 = us(r[:, 0].min()), us(r[:, 0].max())
    arrive = us(np.median(r[:, 4])); wret = us(r[:, 5].min()); end = us(r[:, 7].max())
    if prev_end is None:
        early = ready = "-"
    else:
        early = int((us(r[:, 0]) < prev_end).sum()); ready = int((us(r[:, 4]) < prev_end).sum())
    dss = "-" if prev_start is None else f"{s0 - prev_start:.2f}"
    P(f"| {i} | {s0:.2f} | {s1:.2f} | {arrive:.2f} | {wret:.2f} | {end:.2f} | {early} | {ready} | {dss} |")
    prev_end, prev_start = end, s0
P("")
starts = np.array([us(r[:, 0].min()) for r in rows]); ends = np.array([us(r[:, 7].max()) for r in rows])
P(f"Start-to-start distance of consecutive launches: median {np.median(np.diff(starts)):.2f} us; one launch from its first CTA start "
  f"to its last CTA done: median {np.median(ends - starts):.2f} us.  The difference is what the overlap hides: the CTAs of launch n+1 "
  "become resident while launch n drains, stage their key frames and table, lerp and build the triangle normals (shared memory "
  "only) and then sit in `griddepcontrol.wait` until the launch in front of them ON THEIR STREAM has completed and flushed -- no "
  "global write of a launch is issued before that." + (" On one stream that is launch n itself (column 5 >= column 6 of the row above)."
  if a.streams == 1 else f" On {a.streams} streams it is launch n+1-{a.streams}, which finished long ago: launch n+1 does not wait for launch n at all, "
  "the two run side by side and the SMs that launch n leaves idle in its tail are filled at once."))
d = a.streams                                       # launch i follows launch i - streams on its stream
chk = all(us(rows[i][:, 5].min()) >= ends[i - d] - 0.2 for i in range(d, K))
P("")
P(f"Order check (every launch's first `wait` return is not earlier than the last CTA end of the launch in front of it ON ITS STREAM, "
  f"launch i - {d}; 0.2 us stamp slack): {'holds' if chk else 'VIOLATED'}." + ("" if d == 1 else
  f"  With {d} streams launch i + 1 is NOT ordered behind launch i: its CTAs run their whole program while launch i drains (column 5 of a "
  "row lies before column 6 of the row above)."))
txt = "\n".join(lines)
print(txt)
if a.out:
    with open(a.out, "w") as f:
        f.write(txt + "\n")
