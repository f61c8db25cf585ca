#!/usr/bin/env python
"""Where the end-to-end step goes (host wall clock): bq2_build_records / bq2_frame_submit / bq2_frame_wait."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import berserkerquake2_b200 as bq2
from berserkerquake2_b200 import synth, host, _capi as c
import bench

n_ents, lights = 1024, 4
ctx = bq2.Context(0)
nb = None
mids = np.zeros(n_ents, np.int32)
for e in range(n_ents):
    blob = synth.md2(16, 17, 198, e)
    if nb is None:
        nb = host.md2_neighbours(blob)
    mids[e] = ctx.upload_md2(blob, nb)
we, wl, pe, pl = bench.make_scene(n_ents, lights)
we["model"] = mids
n_pairs = len(pe)
ents_h = np.zeros(n_ents, c.ENTITY_DTYPE); pairs_h = np.zeros(n_pairs, c.PAIR_DTYPE)
kept = C.c_int32(0); fd = c.FrameDesc(); fo = c.FrameOut()
t = np.zeros(3)
N = 300
for it in range(0 if os.environ.get('BQ2_WORLD_ONLY') else N + 20):
    t0 = time.perf_counter()
    c.lib.bq2_build_records(ctx._h, c.ptr(we), n_ents, c.ptr(wl), len(wl), c.ptr(pe), c.ptr(pl), n_pairs, None,
                            c.ptr(ents_h), c.ptr(pairs_h), C.byref(kept))
    t1 = time.perf_counter()
    fd.ents = ents_h.ctypes.data; fd.n_ents = n_ents; fd.pairs = pairs_h.ctypes.data; fd.n_pairs = kept.value
    fd.want = bq2.WANT_SHADOW; fd.index_stride = 0
    c.lib.bq2_frame_submit(ctx._h, C.byref(fd))
    t2 = time.perf_counter()
    c.lib.bq2_frame_wait(ctx._h, C.byref(fo))
    t3 = time.perf_counter()
    if it >= 20:
        t += (t1 - t0, t2 - t1, t3 - t2)
wd = ctx.world_desc(we, wl, pe, pl, None, bq2.WANT_SHADOW, 0)
tw = np.zeros(2)
for it in range(0 if os.environ.get('BQ2_WORLD_ONLY') else N + 20):
    t0 = time.perf_counter()
    c.lib.bq2_world_submit(ctx._h, C.byref(wd))
    t1 = time.perf_counter()
    c.lib.bq2_frame_wait(ctx._h, C.byref(fo))
    t2 = time.perf_counter()
    if it >= 20:
        tw += (t1 - t0, t2 - t1)
print("world path us per step: submit %.1f  wait %.1f  total %.1f" % (*(tw / N * 1e6), tw.sum() / N * 1e6))
print("us per step: build_records %.1f  frame_submit %.1f  frame_wait %.1f  total %.1f" % (*(t / N * 1e6), t.sum() / N * 1e6))
# pipelined, as bench.py's e2e leg runs it: three frames in flight (submit frame i+2, then wait for frame i)
for depth in ((3,) if os.environ.get('BQ2_WORLD_ONLY') else (2, 3)):
    tp = np.zeros(2)
    for _ in range(depth - 1):
        c.lib.bq2_world_submit(ctx._h, C.byref(wd))
    for it in range(N + 20):
        t0 = time.perf_counter()
        c.lib.bq2_world_submit(ctx._h, C.byref(wd))
        t1 = time.perf_counter()
        c.lib.bq2_frame_wait(ctx._h, C.byref(fo))
        t2 = time.perf_counter()
        if it >= 20:
            tp += (t1 - t0, t2 - t1)
    for _ in range(depth - 1):
        c.lib.bq2_frame_wait(ctx._h, C.byref(fo))
    print("world path, %d frames in flight, us per frame: submit %.1f  wait %.1f  total %.1f"
          % (depth, *(tp / N * 1e6), tp.sum() / N * 1e6))
