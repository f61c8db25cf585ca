#!/usr/bin/env python
"""BASELINE config 5: stress sweep over the number of (entity, light) pairs on this rank's GPU.

  python tools/sweep.py [--lights 8] [--models 4096] [--out gpurun_out/sweep.md]
  torchrun --nproc-per-node N tools/sweep.py ...        (weak scaling: every rank runs the same sizes)

For each size: kernel time (records resident, L2 flushed between replays, CUDA events on the ctx stream),
pipelined end-to-end time through the C ABI, and the SURVEY 8(d) roofline fraction.  Entities cycle through a
pool of distinct models (a distinct model per entity up to --models)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lights", type=int, default=8)
    ap.add_argument("--models", type=int, default=4096)
    ap.add_argument("--pairs", type=int, nargs="*", default=[1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.md"))
    a = ap.parse_args()
    import torch
    import berserkerquake2_b200 as bq2
    from berserkerquake2_b200 import synth, host, shard, _capi as c
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = bq2.Context(local)
    V, T = synth.sphere_dims(bench.SLICES, bench.RINGS)
    nb, mids = None, []
    for m in range(a.models):
        blob = synth.md2(bench.SLICES, bench.RINGS, bench.FRAMES, rank * a.models + m)
        nb = host.md2_neighbours(blob) if nb is None else nb
        mids.append(ctx.upload_md2(blob, nb))
    mids = np.array(mids, np.int32)
    flush = torch.empty(bench.L2_FLUSH_BYTES, dtype=torch.uint8, device=f"cuda:{local}")
    peak, _ = bench.peaks()
    rows = []
    for n_pairs in a.pairs:
        n_ents = max(1, n_pairs // a.lights)
        we, wl, pe, pl = bench.make_scene(n_ents, a.lights)
        we["model"] = mids[np.arange(n_ents) % len(mids)]
        ents, pairs = ctx.build_records(we, wl, pe, pl)
        res = ctx.frame(ents, pairs)
        stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))   # of THIS frame's state
        tot = int(res.total_indices)
        steps = int(max(5, min(200, 2e9 // max(tot, 1))))
        for _ in range(3):
            with torch.cuda.stream(stream):
                flush.zero_(); ctx.replay()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for e0, e1 in evs:
            with torch.cuda.stream(stream):
                flush.zero_(); e0.record(stream); ctx.replay(); e1.record(stream)
        torch.cuda.synchronize()
        k_ms = float(np.median([e0.elapsed_time(e1) for e0, e1 in evs]))
        # end to end like bench.py: bq2_world_submit + bq2_frame_wait from host arrays
        pin = [ctx.host_array(x) for x in (we, wl, pe, pl)]       # page-locked caller arrays (bq2_host_alloc)
        wd = ctx.world_desc(*pin, None, bq2.WANT_SHADOW, 0)
        fo = c.FrameOut()
        k = max(6, steps // 4)
        depth = 4 if n_pairs <= (1 << 18) else 2            # frames in flight (each owns its result arrays: memory)
        for _ in range(depth - 1):
            assert c.lib.bq2_world_submit(ctx._h, C.byref(wd)) == 0
        for _ in range(3 * depth):                          # every frame state allocates and captures its graph
            assert c.lib.bq2_world_submit(ctx._h, C.byref(wd)) == 0 and c.lib.bq2_frame_wait(ctx._h, C.byref(fo)) == 0
        t0 = time.perf_counter()
        for _ in range(k):
            c.lib.bq2_world_submit(ctx._h, C.byref(wd)); c.lib.bq2_frame_wait(ctx._h, C.byref(fo))
        e2e_ms = (time.perf_counter() - t0) * 1e3 / k
        for _ in range(depth - 1):
            c.lib.bq2_frame_wait(ctx._h, C.byref(fo))
        assert fo.total_indices == tot
        k_ms, e2e_ms = shard.max_over_ranks([k_ms, e2e_ms], f"cuda:{local}", dist)
        alg = bench.algorithmic_bytes(n_ents, len(pe), V, T, tot)
        mint = bench.min_traffic_bytes(n_ents, len(pe), V, T, tot)
        rows.append(dict(pairs_per_gpu=len(pe), entities_per_gpu=n_ents, lights=a.lights, n_gpus=world, kernel_ms=k_ms,
                         pairs_per_s=world * len(pe) / (k_ms * 1e-3), tris_per_s=world * tot / 3 / (k_ms * 1e-3),
                         lerped_verts_per_s=world * n_ents * V / (k_ms * 1e-3),
                         frac=alg / (k_ms * 1e-3) / 1e9 / peak, frac_min_traffic=mint / (k_ms * 1e-3) / 1e9 / peak,
                         e2e_ms=e2e_ms, e2e_pairs_per_s=world * len(pe) / (e2e_ms * 1e-3)))
        if rank == 0:
            print(json.dumps(rows[-1]), flush=True)
    if rank == 0:
        with open(a.out, "w") as f:
            f.write(f"| pairs/GPU | entities x lights | GPUs | kernel ms | pairs/s | sv tris/s | lerped verts/s | roofline frac (8d) | frac (min traffic) | e2e ms | e2e pairs/s |\n")
            f.write("|---|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write(f"| {r['pairs_per_gpu']} | {r['entities_per_gpu']} x {r['lights']} | {r['n_gpus']} | {r['kernel_ms']:.4f} | {r['pairs_per_s']:.3e} | "
                        f"{r['tris_per_s']:.3e} | {r['lerped_verts_per_s']:.3e} | {r['frac']:.3f} | {r['frac_min_traffic']:.3f} | {r['e2e_ms']:.3f} | {r['e2e_pairs_per_s']:.3e} |\n")
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
