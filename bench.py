#!/usr/bin/env python
"""bench.py -- the hot path (MD2 keyframe lerp -> light-facing -> silhouette / cap / extrusion index
buffers) on BASELINE.json's configs[1]: 1,024 synthetic MD2 entities x 4 lights per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU)
  python bench.py --impl reference [...]                          the reference's own CPU code, all host cores

One "step" = one pass of the path over the whole batch = ONE launch of the fused kernel.
  value     shadow-volume triangles/s (sum index_count / 3 / time) over EXACTLY K steps launched back to back between one
            pair of CUDA events on the launching stream, max over ranks.  Inputs resident in HBM and larger than L2:
            the steps cycle through a ring of resident frames (bq2_frame_keep), each drawing on its own device copy of
            the models, so a cycle reads more key-frame / triangle-table / record bytes than the L2 holds (no flush).
            The replays are launched with programmatic dependent launch: stream order of every write is kept, the
            read-only front of the next step runs under the tail of the current one.
  e2e       the same metric through the C ABI from HOST arrays: bq2_world_submit (host: validation, numbering,
            AngleVectors; H2D; record-building kernel + fused kernel; D2H of the per-pair counts) + bq2_frame_wait,
            four frames in flight, the loop itself in C (csrc/bq2_frame_loop.c), wall clock over K iterations
  roofline  SURVEY.md 8(d) algorithmic bytes per launch / (timed region / K), vs MEASURED_PEAKS.json; the time of ONE
            launch alone after an L2 flush is reported beside it (isolated_launch_ms)
  cpu_baseline  oracle/_ref (the unmodified reference, compiled headless) on the host cores, N=1 only

The oracle (oracle/) is loaded ONLY by the cpu_baseline / --impl reference legs below, as the thing that
is reported beside the GPU number -- never by the measured GPU path.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (entities per GPU, lights per entity, description)
    "c2": (1024, 4, "BASELINE configs[1]: 1,024 synthetic MD2 entities (V=258,T=512,198 frames) x 4 lights"),
    "c3shard": (8192, 8, "one GPU's share of BASELINE configs[2]: 8,192 entities x 8 lights"),
}
SLICES, RINGS, FRAMES = 16, 17, 198
RADIUS = 300.0
L2_FLUSH_BYTES = 256 << 20
E2E_REPEATS = 9            # the K-frame wall-clock measurement is repeated at least this often, and for at least
E2E_MIN_SECONDS = 0.5      # this long (nine 100-frame runs last 20 ms: one hiccup of the host would cover them all);
E2E_MAX_REPEATS = 999      # the median run is reported, min/max alongside
E2E_IN_FLIGHT = int(os.environ.get("BQ2_BENCH_IN_FLIGHT", "4"))   # frames the e2e leg keeps in flight (the library allows
                           # four; results of each stay separate; the variable is for experiment builds of the library)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(n_ents, n_pairs, V, T, total_indices):
    """per entity read 2(40+4V)+36, write 12V; per pair read 6T+12T+16, write
    12V + 4*ceil(T/32) + 4*I + 4 (I = measured index count)."""
    per_ent = 2 * (40 + 4 * V) + 36 + 12 * V
    per_pair = 18 * T + 16 + 12 * V + 4 * ((T + 31) // 32) + 4
    return n_ents * per_ent + n_pairs * per_pair + 4 * total_indices


def min_traffic_bytes(n_ents, n_pairs, V, T, total_indices):
    """What THIS layout has to move at minimum: per entity two 4V-byte frames + the 128-byte record + the
    12T-byte index/neighbour table (read once per entity, not per pair) + 12V out; per pair a 48-byte
    record in, 12V + 4*ceil(T/32) + 4 + 4*I out."""
    per_ent = 2 * 4 * V + 128 + 12 * T + 12 * V
    per_pair = 48 + 12 * V + 4 * ((T + 31) // 32) + 4
    return n_ents * per_ent + n_pairs * per_pair + 4 * total_indices


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.start, self.frozen = [], None, 0, False
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            if not self.frozen:
                self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        self.start = len(self.rows)

    def since_mark(self):
        return len(self.rows) - self.start

    def freeze(self):
        self.frozen = True

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.start:]:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for k, n in enumerate(names):
                if r[3 + k].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons),
                "how": "nvidia-smi -lms 100 from the start of the timed region, the same step kept running until "
                       "at least 4 samples were seen"}


# ------------------------------------------------------------------------------------------------------
def make_scene(n_ents, lights, first_global_ent=0, total_ents=None, step=0):
    """World entities / lights of this rank's shard of the global entity list (entity-major sharding)."""
    from berserkerquake2_b200 import synth
    import berserkerquake2_b200 as bq2
    total = total_ents or n_ents
    we, lw = synth.scene(total, lights, FRAMES, total, 12345, step)       # one distinct model per entity
    we, lw = we[first_global_ent:first_global_ent + n_ents].copy(), \
        lw[first_global_ent * lights:(first_global_ent + n_ents) * lights].copy()
    wl = np.zeros(len(lw), bq2.WORLD_LIGHT_DTYPE)
    wl["origin"] = lw
    wl["radius"] = RADIUS
    pe = np.repeat(np.arange(n_ents, dtype=np.int32), lights)
    pl = np.arange(n_ents * lights, dtype=np.int32)
    return we, wl, pe, pl


def run_gpu(args):
    import torch
    import berserkerquake2_b200 as bq2
    from berserkerquake2_b200 import synth, host, _capi as c

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n_ents, lights, desc = WORKLOADS[args.workload]
    first, _ = host.shard_entities(n_ents * world, rank, world)           # weak scaling: n_ents per GPU
    ctx = bq2.Context(local)
    t0 = time.time()
    V, T = synth.sphere_dims(SLICES, RINGS)
    # Inputs larger than L2 instead of a flush between steps: RING resident frames, each on its own device copy of the
    # models (a frame reads 2 key-frame rows + the triangle table per entity and its records), so that a cycle through
    # the ring reads more than the L2 holds and every launch finds its inputs in HBM.
    l2_bytes = int(getattr(torch.cuda.get_device_properties(local), "L2_cache_size", 126 << 20))
    step_input_bytes = n_ents * (2 * 4 * V + 12 * T + 128) + n_ents * lights * 48
    ring = int(min(32, max(2, -(-int(1.1 * l2_bytes) // step_input_bytes))))
    nb = None
    blobs = []
    for e in range(n_ents):
        blobs.append(synth.md2(SLICES, RINGS, FRAMES, first + e))
        if nb is None:
            nb = host.md2_neighbours(blobs[0])                             # same topology for every sphere
    mids = np.zeros((ring, n_ents), np.int32)
    model_bytes = 0
    for r in range(ring):
        for e in range(n_ents):
            mids[r, e] = ctx.upload_md2(blobs[e], nb)
            model_bytes += FRAMES * 4 * V + 12 * T
    del blobs
    setup_s = time.time() - t0

    kept_frames, ring_indices = [], []
    for r in range(ring):                              # frame r of the ring: scene step r on model copy r
        we_r, wl_r, pe, pl = make_scene(n_ents, lights, first, n_ents * world, step=r)
        we_r["model"] = mids[r]
        ents_r, pairs_r = ctx.build_records(we_r, wl_r, pe, pl)
        res = ctx.frame(ents_r, pairs_r, want=bq2.WANT_SHADOW)
        assert res.overflow_pairs == 0 and int(res.total_indices) > 0
        kept_frames.append(ctx.keep())                 # its records stay resident in HBM
        ring_indices.append(int(res.total_indices))
        if r == 0:
            we, wl, n_pairs = we_r, wl_r, len(pairs_r)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def steps_resident(k, start=0):
        """k steps back to back on the ctx stream: ONE fused launch each, inputs resident, cycling through the ring"""
        for i in range(k):
            ctx.replay_kept(kept_frames[(start + i) % ring])

    clocks = ClockSampler(local) if rank == 0 else None
    steps_resident(max(args.warmup, 3))
    t_spin = time.perf_counter()                       # idle GPUs sit at ~120 MHz: keep stepping until the
    while time.perf_counter() - t_spin < args.spinup:  # clocks have ramped, then start the timed region
        steps_resident(ring)
        torch.cuda.synchronize()
    barrier()
    if clocks:
        clocks.mark()
    launches0 = ctx.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    steps_resident(args.steps)                         # EXACTLY K steps
    ev1.record(stream)
    barrier()
    ms_total = float(ev0.elapsed_time(ev1))
    launches = ctx.kernel_launches - launches0
    total_indices = sum(ring_indices[i % ring] for i in range(args.steps))       # over the K timed steps
    if clocks:                                         # the timed region can be shorter than nvidia-smi's
        t_more = time.perf_counter()                   # period: keep the SAME steps running until it has
        while clocks.since_mark() < 4 and time.perf_counter() - t_more < 3.0:   # seen a few samples
            steps_resident(ring)
            torch.cuda.synchronize()
        clocks.freeze()
    clock_info = clocks.stop() if clocks else None     # (the sampler polls the driver: not during the wall-clock leg)
    # informational: ONE launch alone between two events, L2 flushed before it by a 256 MiB memset (dirty lines), or by
    # the memset followed by a read of a second buffer (clean lines) -- the latency of an isolated launch
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=f"cuda:{local}")
    flush2 = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=f"cuda:{local}")
    iso = {}
    for name, clean in (("dirty", False), ("clean", True)):
        evs2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 50))]
        for i, (a, b) in enumerate(evs2):
            with torch.cuda.stream(stream):
                flush.zero_()
                if clean:
                    flush2.sum()
                a.record(stream)
                ctx.replay_kept(kept_frames[i % ring])
                b.record(stream)
        torch.cuda.synchronize()
        iso[name] = float(np.median([a.elapsed_time(b) for a, b in evs2]))
    del flush2

    # ---- e2e: HOST records in through the C ABI, counts back, wall clock -----------------------------
    we_c = np.ascontiguousarray(we); wl_c = np.ascontiguousarray(wl)
    # the AngleVectors bases of the rotated entities travel in steps of 64 rows of 36 bytes (bq2_world_submit)
    n_rotated = int(np.count_nonzero((we_c["angles"] != 0).any(axis=1)))
    h2d_basis_bytes = min(((n_rotated + 63) & ~63) * 36, (n_ents * 36 + 15) & ~15)
    ents_h = np.zeros(n_ents, c.ENTITY_DTYPE); pairs_h = np.zeros(n_pairs, c.PAIR_DTYPE)
    kept = C.c_int32(0)
    fd = c.FrameDesc(); fo = c.FrameOut()

    # host arrays as the engine holds them, in page-locked memory from bq2_host_alloc (copied to the device from there)
    we_p, wl_p, pe_p, pl_p = (ctx.host_array(a) for a in (we_c, wl_c, pe, pl))
    wd = ctx.world_desc(we_p, wl_p, pe_p, pl_p, None, bq2.WANT_SHADOW, 0)
    wd_pageable = ctx.world_desc(we_c, wl_c, pe, pl, None, bq2.WANT_SHADOW, 0); wd_pageable._keep = ctx._keep

    def e2e_submit():
        if c.lib.bq2_world_submit(ctx._h, C.byref(wd)) != 0:
            raise RuntimeError(c.lib.bq2_last_error(ctx._h))

    def e2e_wait():
        if c.lib.bq2_frame_wait(ctx._h, C.byref(fo)) != 0:
            raise RuntimeError(c.lib.bq2_last_error(ctx._h))
        return fo.total_indices

    def e2e_run(k, timed=False):
        """k frames through the public C ABI from HOST data.  Each frame: bq2_world_submit (host: per-entity
        validation, numbering and AngleVectors -> pinned -> H2D; the record-building kernel + the fused kernel; D2H of
        the per-pair index counts and totals) and bq2_frame_wait.  E2E_IN_FLIGHT frames in flight (the library's
        documented mode: each frame has its own records, result arrays and stream, so frames i+1, i+2 are packed,
        copied and computed while frame i is on the GPU and every frame's results stay valid for the consumer).
        timed: the clock runs over k submit+wait iterations of the full pipeline (k frames in, k frames out)."""
        tot = 0
        depth = E2E_IN_FLIGHT
        for _ in range(depth - 1):
            e2e_submit()
        t0 = time.perf_counter()
        for _ in range(k):
            e2e_submit()
            tot += e2e_wait()
        dt = time.perf_counter() - t0
        for _ in range(depth - 1):
            e2e_wait()
        return (tot, dt) if timed else tot

    e2e_run(max(args.warmup, 12))      # every frame state of the ring has captured its graph before the timed frames
    barrier()
    _, e2e_py_s = e2e_run(args.steps, timed=True)                  # the same loop driven from Python (ctypes per call)
    # the frame loop as an engine runs it: C on top of the public ABI (csrc/bq2_frame_loop.c), same calls, same depth
    loop = C.CDLL(os.path.join(ROOT, "berserkerquake2_b200", "libbq2loop.so"))
    loop.bq2_frame_loop.restype = C.c_int
    loop.bq2_frame_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    sec, tot64 = C.c_double(0), C.c_uint64(0)

    def e2e_loop(k, desc=None):
        if loop.bq2_frame_loop(ctx._h, C.byref(desc or wd), k, E2E_IN_FLIGHT, C.byref(sec), C.byref(tot64)) != 0:
            raise RuntimeError(c.lib.bq2_last_error(ctx._h))
        return int(tot64.value), float(sec.value)

    e2e_loop(max(args.warmup, 12))
    barrier()
    launches_e2e0 = ctx.kernel_launches
    e2e_runs, t_e2e0 = [], time.perf_counter()                     # K frames each; the median run is reported
    while len(e2e_runs) < E2E_REPEATS or (time.perf_counter() - t_e2e0 < E2E_MIN_SECONDS and len(e2e_runs) < E2E_MAX_REPEATS):
        e2e_runs.append(e2e_loop(args.steps))
    e2e_repeats = len(e2e_runs)
    launches_e2e = (ctx.kernel_launches - launches_e2e0) / e2e_repeats - 2 * (E2E_IN_FLIGHT - 1)   # minus the pipeline fill
    e2e_indices, e2e_s = sorted(e2e_runs, key=lambda r: r[1])[e2e_repeats // 2]
    e2e_spread = [min(r[1] for r in e2e_runs) * 1e3 / args.steps, max(r[1] for r in e2e_runs) * 1e3 / args.steps]
    barrier()
    e2e_loop(max(args.warmup, 12), wd_pageable)                    # the same from pageable arrays (staged by the library)
    _, e2e_pageable_s = e2e_loop(args.steps, wd_pageable)
    t0 = time.perf_counter()                                      # one frame at a time (latency of a frame)
    for _ in range(args.steps):
        e2e_submit(); e2e_wait()
    e2e_serial_s = time.perf_counter() - t0
    # the older route for comparison: records built on the host (bq2_build_records + bq2_frame)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c.lib.bq2_build_records(ctx._h, c.ptr(we_c), n_ents, c.ptr(wl_c), len(wl_c), c.ptr(pe), c.ptr(pl), n_pairs,
                                None, c.ptr(ents_h), c.ptr(pairs_h), C.byref(kept))
        fd.ents = ents_h.ctypes.data; fd.n_ents = n_ents; fd.pairs = pairs_h.ctypes.data; fd.n_pairs = kept.value
        fd.want = bq2.WANT_SHADOW; fd.index_stride = 0
        c.lib.bq2_frame(ctx._h, C.byref(fd), C.byref(fo))
    e2e_host_s = time.perf_counter() - t0

    # ---- SURVEY 8(d) variant: every entity instantiates the SAME model (key frames come from L2 after first touch) --
    we_s = we.copy(); we_s["model"] = mids[0, 0]
    ents_s, pairs_s = ctx.build_records(we_s, wl, pe, pl)
    res_s = ctx.frame(ents_s, pairs_s, want=bq2.WANT_SHADOW)
    shared_indices = int(res_s.total_indices)
    stream_s = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    evs3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(3 + len(evs3)):
        with torch.cuda.stream(stream_s):
            flush.zero_()
            if i >= 3:
                evs3[i - 3][0].record(stream_s)
            ctx.replay()
            if i >= 3:
                evs3[i - 3][1].record(stream_s)
    torch.cuda.synchronize()
    shared_ms = float(np.median([a.elapsed_time(b) for a, b in evs3]))

    # ---- reduce: max time over ranks; NCCL all_gather of the per-shard counters (the only collective) --
    from berserkerquake2_b200 import shard
    dev = f"cuda:{local}"
    per_rank = shard.gather_counters([n_pairs * args.steps, n_ents * V * args.steps, total_indices // 3, total_indices,
                                      e2e_indices // 3], dev, dist)           # summed over the K timed steps
    g_pairs, g_verts, g_tris, g_indices, g_e2e_tris = [int(x) for x in per_rank.sum(0)]
    ms_total_max, e2e_ms_max = shard.max_over_ranks([ms_total, e2e_s * 1e3], dev, dist)

    if rank == 0:
        ms_per_step = ms_total_max / args.steps
        peak, peak_src = peaks()
        alg = algorithmic_bytes(n_ents, n_pairs, V, T, total_indices / args.steps)     # per launch (this rank's mean)
        mint = min_traffic_bytes(n_ents, n_pairs, V, T, total_indices / args.steps)
        k_ms = ms_per_step                                                 # one launch per step
        out = {
            "metric": "shadow-volume tris/s (fused MD2 lerp + light-facing + silhouette/cap/extrusion index "
                      "generation); lerped verts/s alongside",
            "value": g_tris / (ms_total_max * 1e-3), "unit": "tris/s",
            "lerped_verts_per_s": g_verts / (ms_total_max * 1e-3),
            "pairs_per_s": g_pairs / (ms_total_max * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc + " per GPU; one distinct model per entity; entity-major sharding, "
                                          "no data-path collective",
                       "entities_per_gpu": n_ents, "lights_per_entity": lights, "pairs_per_gpu": n_pairs,
                       "verts": V, "tris": T, "frames": FRAMES, "model_bytes_in_hbm_per_gpu": model_bytes,
                       "mean_indices_per_pair": total_indices / args.steps / n_pairs,
                       "l2": f"not flushed: inputs larger than L2 -- the K steps cycle through a ring of {ring} resident frames "
                             f"(scene steps 0..{ring - 1}), each on its own device copy of the {n_ents} models, so one cycle reads "
                             f"{ring * step_input_bytes / 1e6:.0f} MB of key frames, triangle tables and records (L2: "
                             f"{l2_bytes / 1e6:.0f} MB) besides streaming {ring} x the outputs through it",
                       "resident_frames": ring, "input_bytes_per_step": step_input_bytes,
                       "timing": "one pair of CUDA events on the ctx stream around the K back-to-back launches, max over ranks; "
                                 "the launches carry programmatic stream serialization: all K steps run in stream order and "
                                 "complete inside the events, the read-only front of step N+1 (record load, TMA staging, lerp, "
                                 "triangle normals) overlaps the tail of step N; "
                                 "roofline.isolated_launch_ms = one launch alone between two events after an L2 flush"},
            "e2e": {"value": g_e2e_tris / args.steps / (e2e_ms_max * 1e-3 / args.steps), "unit": "tris/s",
                    "ms_per_step": e2e_ms_max / args.steps, "gpu_launches_per_step": launches_e2e / args.steps,
                    "repeats": e2e_repeats, "ms_per_step_min_max_rank0": e2e_spread,
                    "ms_per_step_python_driver": e2e_py_s * 1e3 / args.steps,
                    "ms_per_step_pageable_inputs": e2e_pageable_s * 1e3 / args.steps,
                    "ms_per_step_serial": e2e_serial_s * 1e3 / args.steps,
                    "ms_per_step_host_built_records": e2e_host_s * 1e3 / args.steps,
                    "h2d_bytes_per_step": n_ents * 64 + 16 + len(wl_c) * 16 + n_pairs * 8 + h2d_basis_bytes, "d2h_bytes_per_step": 4 * n_pairs + 24,
                    "what": "per frame: bq2_world_submit (host: per-entity validation, numbering and AngleVectors into one 64-byte "
                            "row per entity that also carries the caller's entity fields; H2D of those rows, the bases of the "
                            "rotated entities and the caller's lights and pairs from the library's pinned buffer (the caller's "
                            "own arrays are page-locked memory from bq2_host_alloc); the record-building kernel + the "
                            "fused kernel; D2H of per-pair index counts + totals) + "
                            "bq2_frame_wait, wall clock over K frames (median of %d such runs) with %d frames in flight (each has its own records, "
                            "result arrays and stream; the clock runs over K submit+wait iterations of the full pipeline, K "
                            "frames in and K frames out); the loop itself is C on the public ABI (csrc/bq2_frame_loop.c), as "
                            "an engine would run it; ms_per_step_python_driver = the same calls made from Python; "
                            "ms_per_step_serial = one at a time; " % (e2e_repeats, E2E_IN_FLIGHT) +
                            
                            "vertex/index buffers stay in HBM for the draw"},
            "shared_model_variant": {
                "what": "SURVEY 8(d): the same frame with every entity instantiating ONE model; scored with the unique-"
                        "input form of the byte formula (the model's frames and tables counted once)",
                "ms_per_step": shared_ms, "tris_per_s": shared_indices / 3 / (shared_ms * 1e-3),
                "algorithmic_bytes_unique_input": algorithmic_bytes(n_ents, n_pairs, V, T, shared_indices)
                    - (n_ents - 1) * 2 * (40 + 4 * V) - (n_pairs - 1) * 18 * T,
                "frac": (algorithmic_bytes(n_ents, n_pairs, V, T, shared_indices) - (n_ents - 1) * 2 * (40 + 4 * V)
                         - (n_pairs - 1) * 18 * T) / (shared_ms * 1e-3) / 1e9 / peak},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (k_ms * 1e-3) / 1e9 / peak, "frac_of_spec_8000": alg / (k_ms * 1e-3) / 1e9 / 8000.0,
                         "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                         "kernel": "bq2_frame_warp_kernel", "kernel_ms": k_ms,
                         "isolated_launch_ms": {"l2_flushed_by_memset": iso["dirty"], "l2_flushed_then_read": iso["clean"],
                                                "what": "ONE launch between two events on an otherwise idle stream: adds "
                                                        "the launch and completion latency that back-to-back steps hide"},
                         "min_traffic_bytes_per_launch": mint,
                         "achieved_min_traffic": mint / (k_ms * 1e-3) / 1e9,
                         "frac_min_traffic": mint / (k_ms * 1e-3) / 1e9 / peak},
            "clocks": clock_info,
            "setup_s": setup_s,
        }
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                with open(prof) as f:
                    out["roofline"]["traffic"] = json.load(f).get(args.workload)
            except (OSError, ValueError):
                pass
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_arm(WORKLOADS["c2"][0], WORKLOADS["c2"][1], steps=None, warmup=1,
                                          budget_s=args.cpu_seconds)
        print(json.dumps(out), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own code (oracle/_ref/libbq2ref.so) or, failing that, the oracle port.
# The reference is single-threaded and keeps its state in globals, so "all host cores" = one worker
# PROCESS per core, each owning a contiguous slice of the pair list.
def cpu_worker(args):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from berserkerquake2_b200 import synth, host
    n_ents, lights = args.cpu_ents, args.cpu_lights
    lo, hi = args.cpu_lo, args.cpu_hi                                      # pair range of this worker
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libbq2ref.so")
    if os.path.exists(ref_so):
        lib = C.CDLL(ref_so); fn = lib.bq2ref_md2_world_bench; lib.bq2ref_init()
    else:
        lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so")); fn = lib.bq2o_md2_world_bench
    fn.restype = C.c_longlong
    fn.argtypes = [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 8
    we, wl, pe, pl = make_scene(n_ents, lights)
    pe, pl = pe[lo:hi], pl[lo:hi]
    n = hi - lo
    blobs, nb = {}, None
    for e in np.unique(pe):
        blobs[int(e)] = synth.md2(SLICES, RINGS, FRAMES, int(e))
        if nb is None:
            nb = host.md2_neighbours(blobs[int(e)])
    bp = (C.c_void_p * n)(*[blobs[int(e)].ctypes.data for e in pe])
    np_ = (C.c_void_p * n)(*[nb.ctypes.data] * n)
    a = dict(frames=we["frame"][pe].astype(np.int32), oldframes=we["oldframe"][pe].astype(np.int32),
             backlerps=we["backlerp"][pe].astype(np.float32), origins=np.ascontiguousarray(we["origin"][pe]),
             oldorigins=np.ascontiguousarray(we["oldorigin"][pe]), angles=np.ascontiguousarray(we["angles"][pe]),
             lights=np.ascontiguousarray(wl["origin"][pl]), radii=np.ascontiguousarray(wl["radius"][pl]))
    P = lambda x: x.ctypes.data_as(C.c_void_p)
    print("ready", flush=True)
    for line in sys.stdin:                              # "go N": N passes over this worker's slice, back to back
        w = line.split()
        if not w or w[0] != "go":
            break
        reps = int(w[1]) if len(w) > 1 else 1
        t0 = time.perf_counter()
        for _ in range(reps):
            tot = fn(bp, np_, n, P(a["frames"]), P(a["oldframes"]), P(a["backlerps"]), P(a["origins"]),
                     P(a["oldorigins"]), P(a["angles"]), P(a["lights"]), P(a["radii"]))
        print(json.dumps({"indices": int(tot), "s": time.perf_counter() - t0}), flush=True)


def cpu_arm(n_ents, lights, steps, warmup, budget_s=12.0, cores=None):
    """Each step = the whole batch (n_ents x lights pairs) once, split over one process per core."""
    cores = cores or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
    n_pairs = n_ents * lights
    cores = max(1, min(cores, n_pairs))
    kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libbq2ref.so")) else "port"
    procs = []
    for w in range(cores):
        lo, hi = n_pairs * w // cores, n_pairs * (w + 1) // cores
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--_cpu_worker", "--cpu-ents", str(n_ents),
                                       "--cpu-lights", str(lights), "--cpu-lo", str(lo), "--cpu-hi", str(hi)],
                                      stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, bufsize=1))
    for p in procs:
        assert p.stdout.readline().strip() == "ready"

    def run(reps):
        """reps passes on every worker at once, no synchronisation between passes (the most favourable way to run the
        reference in parallel); returns (seconds per pass = the slowest worker's time / reps, indices per pass)."""
        for p in procs:
            p.stdin.write(f"go {reps}\n"); p.stdin.flush()
        res = [json.loads(p.stdout.readline()) for p in procs]
        return max(r["s"] for r in res) / reps, sum(r["indices"] for r in res)

    t_warm = time.perf_counter()                        # host clocks ramp like the GPU's: at least half a second warm
    run(max(warmup, 1))
    while time.perf_counter() - t_warm < 0.5:
        run(max(warmup, 1))
    if steps is not None:
        k = steps
        s_per_step, tot = run(k)
    else:                                               # bounded sample: batches of 20 passes until the budget is spent
        times, k, t_begin = [], 0, time.perf_counter()
        while time.perf_counter() - t_begin < budget_s and k < 20000:
            dt, tot = run(20)
            times.append(dt); k += 20
        s_per_step = float(np.mean(times))
    for p in procs:
        p.stdin.write("quit\n"); p.stdin.flush(); p.wait(timeout=10)
    V, T = 258, 512
    return {"value": tot / 3 / s_per_step, "unit": "tris/s", "cores": cores, "kind": kind,
            "lerped_verts_per_s": n_pairs * V / s_per_step, "pairs_per_s": n_pairs / s_per_step,
            "ms_per_step": s_per_step * 1e3, "steps": k,
            "sample": f"{k} passes over the whole {n_ents} x {lights} batch ({n_pairs} pairs, {tot // 3} tris per pass), "
                      f"one process per core, each looping the reference's R_DrawAliasShadowVolume over its slice, no "
                      f"synchronisation between passes (time per pass = the slowest process's total / passes)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_ents, lights, desc = WORKLOADS["c2"]
    r = cpu_arm(n_ents, lights, steps=args.steps, warmup=max(args.warmup, 1))
    out = {"impl": "reference", "metric": "shadow-volume tris/s (fused MD2 lerp + light-facing + silhouette/cap/"
                                          "extrusion index generation); lerped verts/s alongside",
           "value": r["value"], "unit": "tris/s", "lerped_verts_per_s": r["lerped_verts_per_s"],
           "pairs_per_s": r["pairs_per_s"], "n_gpus": args.gpus, "steps": r["steps"], "warmup": max(args.warmup, 1),
           "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": desc + "; the reference's CPU path (R_DrawAliasShadowVolume per pair) on the host "
                                         "cores; the reference lerps once per PAIR, so lerped verts/s counts V per pair",
                      "entities": n_ents, "lights_per_entity": lights},
           "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": r["value"], "unit": "tris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--spinup", type=float, default=0.5, help="seconds of untimed stepping before the timed region")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--_cpu_worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-ents", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-lights", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-lo", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-hi", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args._cpu_worker:
        return cpu_worker(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    main()
