#!/usr/bin/env python
"""Host time per phase of bq2_world_submit on C2 with the frame loop in C, four frames in flight (BQ2_TRACE=host: laps
printed by bq2_destroy), and the loop's own wall clock per frame beside them."""
import ctypes as C
import os
import sys

os.environ["BQ2_TRACE"] = "host"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                       # noqa: E402
import berserkerquake2_b200 as bq2                      # noqa: E402
from berserkerquake2_b200 import synth, host, _capi as c  # noqa: E402
import bench                                            # noqa: E402

n_ents, lights = 1024, 4
ctx = bq2.Context(0)
blobs = [synth.md2(16, 17, 198, e) for e in range(n_ents)]
nb = host.md2_neighbours(blobs[0])
mids = np.array([ctx.upload_md2(b, nb) for b in blobs], np.int32)
we, wl, pe, pl = bench.make_scene(n_ents, lights)
we["model"] = mids
pin = [ctx.host_array(x) for x in (we, wl, pe, pl)]
wd = ctx.world_desc(*pin, None, bq2.WANT_SHADOW, 0)
loop = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(bq2.__file__)), "libbq2loop.so"))
loop.bq2_frame_loop.restype = C.c_int
loop.bq2_frame_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
sec, tot = C.c_double(0), C.c_uint64(0)
# the same entities with only 4 pairs: the host does the same per-entity work, the GPU next to nothing -- what is left is
# the host + driver cost of a frame
pin4 = [pin[0], pin[1], ctx.host_array(pe[:4]), ctx.host_array(pl[:4])]
wd4 = ctx.world_desc(*pin4, None, bq2.WANT_SHADOW, 0)
keep = [ctx._keep]
for name, d in (("C2", wd), ("C2 entities, 4 pairs only", wd4)):
    for depth in [int(x) for x in os.environ.get('DEPTHS', '4,1').split(',')]:
        loop.bq2_frame_loop(ctx._h, C.byref(d), 200, depth, C.byref(sec), C.byref(tot))
        rs = []
        for _ in range(15):
            loop.bq2_frame_loop(ctx._h, C.byref(d), 400, depth, C.byref(sec), C.byref(tot))
            rs.append(sec.value / 400 * 1e6)
        print(f"{name}: C frame loop, {depth} in flight: median {np.median(rs):.2f} us per frame (min {min(rs):.2f}, max {max(rs):.2f})")
ctx.close()
