#!/usr/bin/env python
"""BASELINE config 2 WITH the tangent-space basis: 1,024 MD2 entities x 4 lights, N/T/B rows per entity and LightP / tsH
rows per pair next to the shadow volumes (smooth-vector MD2s: per-vertex rows).  Tangent / binormal bytes are synthetic
anorm indices (the kernel only decodes them).  One launch alone after an L2 flush, CUDA events on the ctx stream.

  python tools/bench_c2_ntb.py [--entities 1024] [--lights 4] [--steps 50] [--flat]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--entities", type=int, default=1024)
    ap.add_argument("--lights", type=int, default=4)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--flat", action="store_true", help="non-smooth MD2: per-triangle T/B/N bytes, per-corner LightP/tsH")
    a = ap.parse_args()
    import torch
    import berserkerquake2_b200 as bq2
    from berserkerquake2_b200 import synth, host
    ctx = bq2.Context(0)
    V, T = synth.sphere_dims(bench.SLICES, bench.RINGS)
    F = bench.FRAMES
    rng = np.random.RandomState(1)
    nb, mids = None, []
    rows = T if a.flat else V
    for e in range(a.entities):
        blob = synth.md2(bench.SLICES, bench.RINGS, F, e)
        nb = host.md2_neighbours(blob) if nb is None else nb
        tb = rng.randint(0, 162, (F, rows)).astype(np.uint8); bn = rng.randint(0, 162, (F, rows)).astype(np.uint8)
        nm = rng.randint(0, 162, (F, rows)).astype(np.uint8) if a.flat else None
        mids.append(ctx.upload_md2(blob, nb, tb, bn, nm, smooth=not a.flat))
    we, wl, pe, pl = bench.make_scene(a.entities, a.lights)
    we["model"] = np.array(mids, np.int32)
    view = np.array([100.0, -50.0, 80.0], np.float32)
    ents, pairs = ctx.build_records(we, wl, pe, pl, view_world=view)
    stream = None
    flush = torch.empty(bench.L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda:0")
    out = {}
    for name, want in (("shadow", bq2.WANT_SHADOW), ("shadow+ntb", bq2.WANT_SHADOW | bq2.WANT_NTB),
                       ("shadow+ntb+tslight", bq2.WANT_SHADOW | bq2.WANT_NTB | bq2.WANT_TSLIGHT)):
        ctx.frame(ents, pairs, want=want)
        stream = torch.cuda.ExternalStream(ctx.stream)
        ts = []
        for i in range(a.steps + 5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                flush.zero_(); e0.record(stream); ctx.replay(); e1.record(stream)
            torch.cuda.synchronize()
            if i >= 5:
                ts.append(e0.elapsed_time(e1))
        out[name] = float(np.median(ts)) * 1e3
    # end to end from host arrays (bq2_world_submit + bq2_frame_wait, the C frame loop of bench.py's e2e leg): frames with
    # per-vertex rows go through the device-built route, a non-smooth MD2 falls back to host-built records
    import ctypes as C
    from berserkerquake2_b200 import _capi as c
    loop = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(bq2.__file__)), "libbq2loop.so"))
    loop.bq2_frame_loop.restype = C.c_int
    loop.bq2_frame_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    pin = [ctx.host_array(x) for x in (we, wl, pe, pl)]
    e2e = {}
    for name, want in (("shadow", bq2.WANT_SHADOW), ("shadow+ntb", bq2.WANT_SHADOW | bq2.WANT_NTB),
                       ("shadow+ntb+tslight", bq2.WANT_SHADOW | bq2.WANT_NTB | bq2.WANT_TSLIGHT)):
        wd = ctx.world_desc(*pin, view, want, 0)
        sec, tot = C.c_double(0), C.c_uint64(0)
        for depth in (16, 1):
            assert loop.bq2_frame_loop(ctx._h, C.byref(wd), 100, depth, C.byref(sec), C.byref(tot)) == 0, c.lib.bq2_last_error(ctx._h)
            rs = []
            for _ in range(9):
                assert loop.bq2_frame_loop(ctx._h, C.byref(wd), 200, depth, C.byref(sec), C.byref(tot)) == 0
                rs.append(sec.value / 200 * 1e6)
            e2e[f"{name}, {depth} in flight"] = float(np.median(rs))
    print(json.dumps({"workload": f"{a.entities} MD2 entities x {a.lights} lights (V={V}, T={T}), {'flat' if a.flat else 'smooth'} basis",
                      "kernel_us": out, "e2e_us_per_frame": e2e}))
    ctx.close()


if __name__ == "__main__":
    main()
