#!/usr/bin/env python
"""BASELINE config 5: stress sweep over the number of (entity, light) pairs on this rank's GPU, the reference's CPU
path on the box's host cores beside it, and EVERY pair of every size checked against the reference.

  python tools/sweep.py [--lights 8] [--models 4096] [--out gpurun_out/sweep.md]
  torchrun --nproc-per-node N tools/sweep.py ...        (weak scaling: every rank runs the same sizes on its own models)

For each size, per GPU:
  kernel ms       ONE launch between two CUDA events on the ctx stream after an L2 flush (median); the small sizes are
                  launch-latency bound and say so
  back-to-back ms the same frame replayed K times from C with programmatic dependent launch (bq2_resident_run), per step;
                  sizes whose inputs + outputs fit the L2 run from L2 here -- marked
  e2e ms          pipelined bq2_world_submit + bq2_frame_wait from page-locked host arrays
  parity          the device's per-pair hash (bq2_hash_results) against the unmodified reference run over all host cores
                  (tests/oracle_lib.reference_pair_hashes): pairs that differ / pairs (rank 0 checks its own shard)
  CPU             the reference's R_DrawAliasShadowVolume over the same pair list, one process per core (rank 0)
Entities cycle through a pool of distinct models (a distinct model per entity up to --models)."""
import argparse
import ctypes as C
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                                            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lights", type=int, default=8)
    ap.add_argument("--models", type=int, default=4096)
    ap.add_argument("--pairs", type=int, nargs="*", default=[1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20])
    ap.add_argument("--no-cpu", action="store_true", help="skip the reference / parity columns")
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
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = bq2.Context(local)
    V, T = synth.sphere_dims(bench.SLICES, bench.RINGS)
    threads = max(1, len(os.sched_getaffinity(0)) // world)
    with ThreadPoolExecutor(max_workers=threads) as pool:
        blobs = list(pool.map(lambda m: synth.md2(bench.SLICES, bench.RINGS, bench.FRAMES, rank * a.models + m), range(a.models)))
    nb = host.md2_neighbours(blobs[0])
    mids = np.array([ctx.upload_md2(b, nb) for b in blobs], np.int32)
    flush = torch.empty(bench.L2_FLUSH_BYTES, dtype=torch.uint8, device=f"cuda:{local}")
    l2 = int(getattr(torch.cuda.get_device_properties(local), "L2_cache_size", 126 << 20))
    peak, _ = bench.peaks()
    ref_ok = rank == 0 and not a.no_cpu and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libbq2ref.so"))
    if ref_ok:
        import oracle_lib
    rows = []
    for n_pairs in a.pairs:
        n_ents = max(1, n_pairs // a.lights)
        we, wl, pe, pl = bench.make_scene(n_ents, a.lights)
        model_of = np.arange(n_ents) % len(mids)
        we["model"] = mids[model_of]
        ents, pairs = ctx.build_records(we, wl, pe, pl)
        res = ctx.frame(ents, pairs)
        stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))   # of THIS frame's state
        tot = int(res.total_indices)
        # ---- parity of every pair + the CPU column (rank 0, its own shard) ---------------------------------------
        bad = cpu_ms = cpu_cores = None
        if ref_ok:
            we_ref = we.copy(); we_ref["model"] = model_of
            got = ctx.hash_results(len(pe))
            want, ref_total, secs, cpu_cores = oracle_lib.reference_pair_hashes(blobs, [nb] * len(blobs), we_ref, wl, pe, pl,
                                                                                timed_passes=3 if n_pairs <= (1 << 16) else 1,
                                                                                want_seconds=True)
            bad = int((got != want).sum()); cpu_ms = secs * 1e3
            assert ref_total == tot
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
        kept = ctx.keep()                                   # back to back from C, programmatic dependent launch
        ctx.run_ring([kept], 0, 5)
        b2b_ms = float(np.median([ctx.run_ring([kept], 0, steps, timed=True) / steps for _ in range(5)]))
        ctx.free_kept(kept)
        uniq = bench.min_traffic_bytes(n_ents, len(pe), V, T, tot)
        in_l2 = uniq < 0.8 * l2
        # end to end like bench.py: bq2_world_submit + bq2_frame_wait from host arrays
        pin = [ctx.host_array(x) for x in (we, wl, pe, pl)]       # page-locked caller arrays (bq2_host_alloc)
        wd = ctx.world_desc(*pin, None, bq2.WANT_SHADOW, 0)
        fo = c.FrameOut()
        k = max(6, steps // 4)
        depth = 8 if n_pairs <= (1 << 16) else 4 if n_pairs <= (1 << 18) else 2   # frames in flight (each owns its result arrays: memory)
        def ok(rc):
            if rc != 0:
                raise RuntimeError(f"rank {rank}, {n_pairs} pairs: status {rc}: {c.lib.bq2_last_error(ctx._h)}")
        for _ in range(depth - 1):
            ok(c.lib.bq2_world_submit(ctx._h, C.byref(wd)))
        for _ in range(3 * depth + 18):                     # every frame state allocates and captures its graph
            ok(c.lib.bq2_world_submit(ctx._h, C.byref(wd))); ok(c.lib.bq2_frame_wait(ctx._h, C.byref(fo)))
        k = max(k, 4 * depth, 24)                           # a window much longer than the pipeline is deep, three of them
        wins = []
        for _ in range(3):
            t0 = time.perf_counter()
            for _ in range(k):
                ok(c.lib.bq2_world_submit(ctx._h, C.byref(wd))); ok(c.lib.bq2_frame_wait(ctx._h, C.byref(fo)))
            wins.append((time.perf_counter() - t0) * 1e3 / k)
        e2e_ms = float(np.median(wins))
        for _ in range(depth - 1):
            ok(c.lib.bq2_frame_wait(ctx._h, C.byref(fo)))
        assert fo.total_indices == tot
        for x in pin:
            c.lib.bq2_host_free(ctx._h, x.ctypes.data)
        k_ms, b2b_ms, e2e_ms = shard.max_over_ranks([k_ms, b2b_ms, e2e_ms], f"cuda:{local}", dist)
        rows.append(dict(pairs_per_gpu=len(pe), entities_per_gpu=n_ents, lights=a.lights, n_gpus=world, kernel_ms=k_ms,
                         b2b_ms=b2b_ms, in_l2=bool(in_l2),
                         pairs_per_s=world * len(pe) / (b2b_ms * 1e-3), tris_per_s=world * tot / 3 / (b2b_ms * 1e-3),
                         lerped_verts_per_s=world * n_ents * V / (b2b_ms * 1e-3),
                         frac_isolated=uniq / (k_ms * 1e-3) / 1e9 / peak, frac_b2b=uniq / (b2b_ms * 1e-3) / 1e9 / peak,
                         e2e_ms=e2e_ms, e2e_pairs_per_s=world * len(pe) / (e2e_ms * 1e-3),
                         parity_bad=bad, cpu_ms=cpu_ms, cpu_cores=cpu_cores,
                         cpu_pairs_per_s=None if cpu_ms is None else len(pe) / (cpu_ms * 1e-3)))
        if rank == 0:
            print(json.dumps(rows[-1]), flush=True)
    if rank == 0:
        with open(a.out, "w") as f:
            f.write(f"BASELINE config 5 on {world} GPU(s) (weak: every GPU runs the listed size on its own models); fractions are of "
                    f"the measured HBM peak ({peak:.0f} GB/s) on the launch's UNIQUE bytes; `*` = inputs + outputs fit the L2, the "
                    "back-to-back figure runs from L2 and is not an HBM fraction.  CPU = the reference on rank 0's host cores over "
                    "ONE GPU's pair list; parity = pairs whose device hash differs from the reference's / pairs checked (rank 0).\n\n")
            f.write("| pairs/GPU | entities x lights | GPUs | 1 launch alone (ms) | back to back (ms/step) | pairs/s (all GPUs) | sv tris/s | "
                    "frac alone | frac back to back | e2e ms/frame | e2e pairs/s | reference CPU (ms, cores) | CPU pairs/s | one GPU vs CPU | parity |\n")
            f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                star = "*" if r["in_l2"] else ""
                cpu = "-" if r["cpu_ms"] is None else f"{r['cpu_ms']:.2f} ({r['cpu_cores']})"
                cps = "-" if r["cpu_ms"] is None else f"{r['cpu_pairs_per_s']:.3e}"
                ratio = "-" if r["cpu_ms"] is None else f"{r['cpu_ms'] / r['b2b_ms']:.0f}x"
                par = "-" if r["parity_bad"] is None else f"{r['parity_bad']} / {r['pairs_per_gpu']}"
                f.write(f"| {r['pairs_per_gpu']} | {r['entities_per_gpu']} x {r['lights']} | {r['n_gpus']} | {r['kernel_ms']:.4f} | {r['b2b_ms']:.4f}{star} | "
                        f"{r['pairs_per_s']:.3e} | {r['tris_per_s']:.3e} | {r['frac_isolated']:.3f} | {r['frac_b2b']:.3f}{star} | "
                        f"{r['e2e_ms']:.3f} | {r['e2e_pairs_per_s']:.3e} | {cpu} | {cps} | {ratio} | {par} |\n")
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
