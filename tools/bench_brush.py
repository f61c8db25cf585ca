#!/usr/bin/env python
"""SURVEY 8f row 4 as a measurement: brush-model volumes, N (brush entity, light) pairs in one call.
End-to-end wall clock of bq2_brush_volumes (pair records + lit flags H2D, kernel, counts D2H, sync) and the
reference's R_DrawBrushModelVolumes on one host core over the same pairs (checker + CPU baseline only)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import berserkerquake2_b200 as bq2
from berserkerquake2_b200 import _capi as c, host
import oracle_lib
from scenes import BrushModel

ctx = bq2.Context(0)
o = oracle_lib.Oracle(); ref = oracle_lib.Reference.load()
rng = np.random.RandomState(1)
models = [BrushModel(16, 12, s) for s in range(8)]                     # ~220 polygons, ~850 vertex slots each
ids = [ctx.upload_brush(m.numverts, m.verts, m.neighbours, m.fence) for m in models]
for n_pairs in (256, 4096):
    pairs, lits, meta = [], [], []
    for k in range(n_pairs):
        m = models[k % 8]
        origin = rng.randn(3) * 50; angles = np.zeros(3) if k % 3 else rng.uniform(-180, 180, 3)
        lw = origin + rng.randn(3) * 120
        lm = host.point_to_model(lw, origin, angles)
        lit = m.lit(lm)
        pairs.append((ids[k % 8], tuple(lm), np.float32(9600.0))); lits.append(lit); meta.append((m, lit, origin, angles, lw, lm))
    pr = np.array(pairs, c.BRUSH_PAIR_DTYPE); li = np.concatenate(lits)
    out = c.BrushOut()
    import ctypes as C
    for _ in range(3):
        c.lib.bq2_brush_volumes(ctx._h, c.ptr(pr), n_pairs, c.ptr(li), C.byref(out))
    t0 = time.perf_counter(); R = 20
    for _ in range(R):
        c.lib.bq2_brush_volumes(ctx._h, c.ptr(pr), n_pairs, c.ptr(li), C.byref(out))
    gpu_ms = (time.perf_counter() - t0) * 1e3 / R
    li_pin = ctx.host_array(li)                                        # the lit bytes in page-locked memory of the ctx
    for _ in range(3):
        c.lib.bq2_brush_volumes(ctx._h, c.ptr(pr), n_pairs, c.ptr(li_pin), C.byref(out))
    t0 = time.perf_counter()
    for _ in range(R):
        c.lib.bq2_brush_volumes(ctx._h, c.ptr(pr), n_pairs, c.ptr(li_pin), C.byref(out))
    gpu_ms_pinned = (time.perf_counter() - t0) * 1e3 / R
    vb = sum(int(out.counts[2 * k]) for k in range(n_pairs)); ib = sum(int(out.counts[2 * k + 1]) for k in range(n_pairs))
    sample = meta[:256]
    t0 = time.perf_counter()
    for m, lit, origin, angles, lw, lm in sample:
        (ref.brush_volume(m, lit, origin, angles, lw, 300.0) if ref else o.brush_volume(m, lit, lm, 300.0))
    cpu_ms_per_pair = (time.perf_counter() - t0) * 1e3 / len(sample)
    m, lit, origin, angles, lw, lm = meta[5]
    got = ctx.brush_volumes(pr[5:6], lit)[0]; want = o.brush_volume(m, lit, lm, 300.0)
    assert np.array_equal(got[0].view(np.uint32), want[0].view(np.uint32)) and np.array_equal(got[1], want[1])
    bytes_out = vb * 24 + ib * 4
    print(json.dumps({"pairs": n_pairs, "gpu_ms_per_call_e2e": gpu_ms, "gpu_ms_per_call_e2e_lit_page_locked": gpu_ms_pinned, "pairs_per_s": n_pairs / (gpu_ms * 1e-3),
                      "vertex_pairs": vb, "indices": ib, "output_GBps_e2e": bytes_out / (gpu_ms * 1e-3) / 1e9,
                      "cpu_ms_per_pair_1core_incl_python_marshalling": cpu_ms_per_pair,
                      "cpu_kind": "reference" if ref else "port"}))
