#!/usr/bin/env python
"""SURVEY 8f rows 1-2 as measurements: load-time precompute on the device vs the CPU restatement of the
reference loops (oracle, one core) on the same mesh.  The oracle is used here only as the CPU baseline and checker."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import berserkerquake2_b200 as bq2
from berserkerquake2_b200 import synth, host
import oracle_lib

o = oracle_lib.Oracle()
ctx = bq2.Context(0)
SC = float(np.float32(np.cos(np.deg2rad(45.0))))


def t(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); best = min(best, time.perf_counter() - t0)
    return best, r


rows = []
for name, (S, R, F) in {"MD2 player-size (V=258, T=512, 198 frames)": (16, 17, 198), "MD2 maximum (V=2048, T=4092, 40 frames)": (62, 34, 40)}.items():
    blob = synth.md2(S, R, F, 1); st = synth.md2_st(blob)
    c_nb, nb = t(lambda: o.md2_neighbours(blob), 1)
    h_nb, nb2 = t(lambda: host.md2_neighbours(blob))
    g_nb, nb3 = t(lambda: ctx.gpu_md2_neighbours(blob))
    assert np.array_equal(nb, nb2) and np.array_equal(nb, nb3)
    c_tg, (tt, bb) = t(lambda: o.md2_tangents_smooth(blob, st, SC), 1)
    g_tg, (_, gt, gb) = t(lambda: ctx.gpu_md2_tangents(blob, st, SC, True))
    assert np.array_equal(tt, gt) and np.array_equal(bb, gb)
    rows.append(dict(mesh=name, neighbours_cpu_reference_loop_ms=c_nb * 1e3, neighbours_host_sorted_edges_ms=h_nb * 1e3,
                     neighbours_gpu_ms_incl_copies=g_nb * 1e3, tangents_cpu_reference_loop_ms=c_tg * 1e3,
                     tangents_gpu_ms_incl_copies=g_tg * 1e3))
verts, idx, st = synth.md3_mesh(60, 69, 16, 3)
c_nb, nb = t(lambda: o.md3_neighbours(idx), 1)
g_nb, nb3 = t(lambda: ctx.gpu_md3_neighbours(idx))
assert np.array_equal(nb, nb3)
want = verts.copy(); c_tg, _ = t(lambda: o.md3_tangents(want, st, idx), 1)
got = verts.copy(); g_tg, _ = t(lambda: ctx.gpu_md3_tangents(got, st, idx))
assert np.array_equal(got["tangent"], want["tangent"]) and np.array_equal(got["binormal"], want["binormal"])
rows.append(dict(mesh="MD3 mesh near the limits (V=4082, T=8160, 16 frames)", neighbours_cpu_reference_loop_ms=c_nb * 1e3,
                 neighbours_gpu_ms_incl_copies=g_nb * 1e3, tangents_cpu_reference_loop_ms=c_tg * 1e3, tangents_gpu_ms_incl_copies=g_tg * 1e3))
for r in rows:
    print(json.dumps(r))
