#!/usr/bin/env python
"""One C2 frame at a time through the public ABI: host time inside bq2_world_submit and inside bq2_frame_wait (median
over 2,000 frames, perf_counter around the two ctypes calls), and -- in a second context with BQ2_TRACE=1, direct
launches with CUDA events between the frame's stream operations -- the GPU time of each stage."""
import ctypes as C
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                       # noqa: E402


def main():
    import berserkerquake2_b200 as bq2
    from berserkerquake2_b200 import synth, host, _capi as c
    import bench
    n_ents, lights = 1024, 4
    ctx = bq2.Context(0)
    blobs = [synth.md2(16, 17, 198, e) for e in range(n_ents)]
    nb = host.md2_neighbours(blobs[0])
    mids = np.array([ctx.upload_md2(b, nb) for b in blobs], np.int32)
    we, wl, pe, pl = bench.make_scene(n_ents, lights)
    we["model"] = mids
    pin = [ctx.host_array(x) for x in (we, wl, pe, pl)]
    wd = ctx.world_desc(*pin, None, bq2.WANT_SHADOW, 0)
    fo = c.FrameOut()
    sub, wai = [], []
    for i in range(2300):
        t0 = time.perf_counter()
        a = c.lib.bq2_world_submit(ctx._h, C.byref(wd))
        t1 = time.perf_counter()
        b = c.lib.bq2_frame_wait(ctx._h, C.byref(fo))
        t2 = time.perf_counter()
        assert a == 0 and b == 0
        if i >= 300:
            sub.append((t1 - t0) * 1e6); wai.append((t2 - t1) * 1e6)
    print("serial C2 frame (%s): bq2_world_submit %.1f us  bq2_frame_wait %.1f us  sum %.1f us (medians; perf_counter around ctypes calls)"
          % (os.environ.get("BQ2_TRACE", "no trace"), np.median(sub), np.median(wai), np.median(np.array(sub) + np.array(wai))))
    ctx.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        main()
    else:
        for env in ({}, {"BQ2_TRACE": "1"}, {"BQ2_NO_GRAPH": "1"}):
            e = dict(os.environ); e.update(env)
            print("#", env or "default")
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=e, capture_output=True, text=True)
            print(r.stdout.strip()); print("\n".join(r.stderr.strip().splitlines()[-3:]))
