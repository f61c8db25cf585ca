#!/usr/bin/env python
"""bench.py -- the hot path (MD2 keyframe lerp -> light-facing -> silhouette / cap / extrusion index
buffers) on BASELINE.json's configs[1]: 1,024 synthetic MD2 entities x 4 lights per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU)
  python bench.py --impl reference [...]                          the reference's own CPU code, all host cores

One "step" = one pass of the path over the whole batch = ONE launch of the fused kernel.
  value     shadow-volume triangles/s (sum index_count / 3 / time) over EXACTLY K steps launched back to back FROM C
            (bq2_resident_run) between one pair of CUDA events on the launching stream, max over ranks; the K-step
            region is repeated (a barrier in front of each) and the median repeat is reported.  Inputs AND outputs are
            resident in HBM and larger than L2: the steps cycle through a ring of resident frames (bq2_frame_keep_own),
            each drawing on its own device copy of the models and writing its OWN result arrays, so a cycle reads and
            writes several times what the L2 holds (no flush).  The replays carry programmatic dependent launch:
            stream order of every write is kept, the read-only front of the next step runs under the tail of the
            current one.
  roofline  bound = HBM.  frac = UNIQUE bytes per launch (SURVEY 8(d) in its unique-input form: what the launch must
            read and write once) / (timed region / K) / MEASURED_PEAKS hbm_gbs; SURVEY 8(d)'s per-pair form (table
            bytes charged per pair) beside it; traffic = dram__bytes_read + dram__bytes_write per step, MEASURED IN
            THIS RUN by a child process under ncu's range replay over the same K back-to-back steps.
  e2e       the same metric through the C ABI from HOST arrays: bq2_world_submit (host: validation, numbering,
            AngleVectors; H2D; record-building kernel + fused kernel; a last kernel that stores the per-pair counts and
            totals to the page-locked result buffer) + bq2_frame_wait,
            sixteen frames in flight, the loop itself in C (csrc/bq2_frame_loop.c), wall clock over K iterations
  c3_strong BASELINE configs[2] as written: 65,536 entities x 8 lights sharded entity-major over the N GPUs
            (strong scaling), same timing rules
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
E2E_IN_FLIGHT = int(os.environ.get("BQ2_BENCH_IN_FLIGHT", "16"))  # frames the e2e leg keeps in flight (the library allows
                           # sixteen, bq2_max_frames_in_flight; the results of each stay separate); 8 / 4 / 2 / 1 beside it


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


METRIC = ("shadow-volume tris/s (fused MD2 lerp + light-facing + silhouette/cap/extrusion index generation); "
          "lerped verts/s alongside")


def rank_cpus(local, world):
    """The logical CPUs of the physical cores that belong to local rank `local` of `world`: the cores this process may
    use are grouped by /sys/.../topology/thread_siblings_list and dealt out in contiguous blocks."""
    cpus = sorted(os.sched_getaffinity(0))
    cores = {}
    for c in cpus:
        key = (c,)
        try:
            with open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list") as f:
                txt = f.read().strip()
            sib = set()
            for part in txt.split(","):
                lo, _, hi = part.partition("-")
                sib.update(range(int(lo), int(hi or lo) + 1))
            key = tuple(sorted(x for x in sib if x in cpus)) or (c,)
        except (OSError, ValueError):
            pass
        cores[key] = True
    phys = sorted(cores)
    per = max(1, len(phys) // max(world, 1))
    mine = phys[local * per:(local + 1) * per] or phys
    return sorted({c for core in mine for c in core})


def workload_config(name):
    """The keys that define the workload -- identical on both arms (the driver compares them)."""
    n_ents, lights, desc = WORKLOADS[name]
    return {"workload": desc + " per GPU; one distinct model per entity; entity-major sharding, no data-path collective",
            "entities_per_gpu": n_ents, "lights_per_entity": lights, "pairs_per_gpu": n_ents * lights,
            "verts": 258, "tris": 512, "frames": FRAMES}


def ring_size(l2_bytes, n_ents, lights, V, T, cap=32):
    """Resident frames in the ring: enough that one cycle READS more than the L2 holds (outputs come on top)."""
    step_input_bytes = n_ents * (2 * 4 * V + 12 * T + 128) + n_ents * lights * 48
    return int(min(cap, max(1, -(-int(1.1 * l2_bytes) // step_input_bytes)))), step_input_bytes


def build_ring(ctx, n_ents, lights, first, total_ents, ring, own_results, threads=8, chunk=1024):
    """`ring` resident frames of n_ents entities x lights (scene steps 0..ring-1 of the global entity list, this
    rank's entities [first, first + n_ents)), frame r drawing on its own device copy of the models and, with
    own_results, writing its own result arrays.  Models are generated by `threads` host threads (C generator, GIL
    released) a chunk at a time, so that 65,536 of them never sit in host memory at once."""
    import berserkerquake2_b200 as bq2
    from berserkerquake2_b200 import synth, host
    from concurrent.futures import ThreadPoolExecutor
    V, T = synth.sphere_dims(SLICES, RINGS)
    t0 = time.time()
    mids = np.zeros((ring, n_ents), np.int32)
    nb, model_bytes = None, 0
    with ThreadPoolExecutor(max_workers=max(1, threads)) as pool:
        for lo in range(0, n_ents, chunk):
            hi = min(n_ents, lo + chunk)
            blobs = list(pool.map(lambda e: synth.md2(SLICES, RINGS, FRAMES, first + e), range(lo, hi)))
            if nb is None:
                nb = host.md2_neighbours(blobs[0])                         # same topology for every sphere
            for r in range(ring):
                for e in range(lo, hi):
                    mids[r, e] = ctx.upload_md2(blobs[e - lo], nb)
                    model_bytes += FRAMES * 4 * V + 12 * T
            del blobs
    kept, ring_indices, scene0 = [], [], None
    for r in range(ring):                              # frame r of the ring: scene step r on model copy r
        we_r, wl_r, pe, pl = make_scene(n_ents, lights, first, total_ents, step=r)
        we_r["model"] = mids[r]
        ents_r, pairs_r = ctx.build_records(we_r, wl_r, pe, pl)
        res = ctx.frame(ents_r, pairs_r, want=bq2.WANT_SHADOW)
        assert res.overflow_pairs == 0 and int(res.total_indices) > 0
        kept.append(ctx.keep(own_results=own_results))  # its records (and result arrays) stay resident in HBM
        ring_indices.append(int(res.total_indices))
        if r == 0:
            scene0 = (we_r, wl_r, pe, pl, len(pairs_r))
    return {"kept": kept, "indices": ring_indices, "scene0": scene0, "mids": mids, "model_bytes": model_bytes,
            "setup_s": time.time() - t0, "V": V, "T": T}


def timed_regions(ctx, kept, steps, repeats, barrier, dev, dist, streams=1):
    """`repeats` device-timed regions of EXACTLY `steps` back-to-back launches each (bq2_resident_run[_streams]:
    launched from C between two CUDA events), a barrier + synchronize in front of each; every region's time is the MAX
    over ranks.  streams > 1: the launches are dealt round-robin onto that many streams, as independent frames in
    flight run (the events are on the first stream, the others are joined into it before the second event).
    Returns (median region ms, all region ms)."""
    from berserkerquake2_b200 import shard
    if len(kept) < 2:
        streams = 1                                     # one resident frame: its launches share result arrays, one stream
    ms = []
    for r in range(repeats):
        barrier()
        ms.append(ctx.run_ring(kept, (r * steps) % len(kept), steps, timed=True, streams=streams))
    barrier()
    ms = shard.max_over_ranks(ms, dev, dist)
    return float(np.median(ms)), [float(x) for x in ms]


def roofline_block(n_ents, n_pairs, V, T, indices_per_launch, kernel_ms, kernel="bq2_frame_warp_kernel"):
    peak, peak_src = peaks()
    uniq = min_traffic_bytes(n_ents, n_pairs, V, T, indices_per_launch)
    alg = algorithmic_bytes(n_ents, n_pairs, V, T, indices_per_launch)
    sec = kernel_ms * 1e-3
    return {"bound": "hbm", "achieved": uniq / sec / 1e9, "peak": peak, "unit": "GB/s", "frac": uniq / sec / 1e9 / peak,
            "traffic": None, "peak_source": peak_src, "kernel": kernel, "kernel_ms": kernel_ms,
            "bytes_per_launch": uniq,
            "bytes_definition": "UNIQUE bytes of one launch = SURVEY 8(d) in its unique-input form: per entity read two "
                                "4V-byte key-frame rows + the 128-byte record + the 12T-byte index/neighbour table "
                                "(once per entity, as the kernel stages it), write 12V; per pair read the 48-byte "
                                "record, write 12V + 4*ceil(T/32) + 4 + 4*I (I = measured index count)",
            "frac_of_spec_8000": uniq / sec / 1e9 / 8000.0,
            "survey_8d_per_pair_form": {"bytes_per_launch": alg, "achieved": alg / sec / 1e9, "frac": alg / sec / 1e9 / peak,
                                        "note": "8(d) as tabulated charges 18T table bytes per PAIR; the kernel reads the "
                                                "table once per entity, so this figure can exceed 1 and is not an HBM "
                                                "fraction -- kept for continuity with round 1"}}


# ---- DRAM traffic of the timed region, measured in this run ---------------------------------------------
def traffic_probe(args):
    """Child of measure_traffic(), run UNDER ncu: the same ring, the same K back-to-back launches, bracketed by
    cudaProfilerStart/Stop.  Prints nothing the parent parses; ncu writes the counters."""
    import torch
    import berserkerquake2_b200 as bq2
    from berserkerquake2_b200 import synth
    n_ents, lights, _ = WORKLOADS[args.workload]
    ctx = bq2.Context(0)
    V, T = synth.sphere_dims(SLICES, RINGS)
    l2 = int(getattr(torch.cuda.get_device_properties(0), "L2_cache_size", 126 << 20))
    ring, _ = ring_size(l2, n_ents, lights, V, T)
    R = build_ring(ctx, n_ents, lights, 0, n_ents, ring, own_results=ring > 1, threads=args.threads)
    ctx.run_ring(R["kept"], 0, max(args.warmup, 3) + ring)
    ctx.profiler_range(True)                            # range replay measures the range as one unit; application replay
    ctx.run_ring(R["kept"], 0, args.steps, timed=True, streams=args.streams if ring > 1 else 1)   # profiles the launches one by one
    ctx.profiler_range(False)
    ctx.close()


def parse_ncu_dram_csv(path):
    """(dram bytes read, dram bytes written, rows) summed over an `ncu --csv --print-units base` log: one row pair for a
    range replay, one per launch for an application replay."""
    import csv
    rd = wr = 0.0
    rows = 0
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    hdr = None
    for row in csv.reader(lines):
        if hdr is None:
            if "Metric Name" in row and "Metric Value" in row:
                hdr = {k: i for i, k in enumerate(row)}
            continue
        try:
            name, val = row[hdr["Metric Name"]], float(row[hdr["Metric Value"]].replace(",", ""))
        except (IndexError, ValueError):
            continue
        if name == "dram__bytes_read.sum":
            rd += val; rows += 1
        elif name == "dram__bytes_write.sum":
            wr += val
    return rd, wr, rows


def measure_traffic(args, log_dir):
    """dram__bytes_read.sum + dram__bytes_write.sum over the K back-to-back steps of the timed region, per step.
    A child process repeats the region under ncu (a number TIMED under a profiler is never reported; byte counters are
    not timings): first as one range (range replay over cudaProfilerStart/Stop: the launches overlap as they do in
    the timed region), else per kernel (application replay: launches serialised, cache state natural)."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    os.makedirs(log_dir, exist_ok=True)
    base = [sys.executable, os.path.abspath(__file__), "--_traffic_probe", "--workload", args.workload,
            "--steps", str(args.steps), "--warmup", str(args.warmup), "--threads", str(args.threads), "--streams", str(args.streams)]
    metrics = "dram__bytes_read.sum,dram__bytes_write.sum"
    tries = [
        ("range", "ncu --replay-mode app-range over cudaProfilerStart/Stop around the K back-to-back launches",
         [ncu, "--replay-mode", "app-range", "--metrics", metrics, "--print-units", "base",
          "--csv", "--clock-control", "none", "--cache-control", "none"]),
        ("kernel", "ncu --replay-mode application, one pass, the K launches of the region profiled one by one",
         [ncu, "--replay-mode", "application", "--profile-from-start", "off", "--metrics", metrics, "--print-units", "base",
          "--csv", "--clock-control", "none", "--cache-control", "none", "-k", "regex:bq2_frame"]),
    ]
    why = []
    for mode, how, cmd in tries:
        log = os.path.join(log_dir, f"traffic_{args.workload}_{mode}.csv")
        try:
            r = subprocess.run(cmd + ["--log-file", log] + base + ["--probe-mode", mode], capture_output=True, text=True,
                               timeout=args.traffic_timeout)
        except (subprocess.TimeoutExpired, OSError) as e:
            why.append(f"{mode}: {type(e).__name__}")
            continue
        with open(log + ".stderr", "w") as f:
            f.write(r.stdout[-4000:] + "\n" + r.stderr[-4000:])
        if r.returncode != 0 or not os.path.exists(log):
            why.append(f"{mode}: rc={r.returncode} {r.stderr.strip()[-200:]}")
            continue
        rd, wr, rows = parse_ncu_dram_csv(log)
        if rows == 0:
            why.append(f"{mode}: no counter rows in {os.path.basename(log)}")
            continue
        return {"bytes_per_step": (rd + wr) / args.steps, "read_per_step": rd / args.steps, "write_per_step": wr / args.steps,
                "steps": args.steps, "rows": rows, "how": how + " (--cache-control none: ncu neither flushes nor invalidates "
                "the caches around a launch, the L2 holds what the steps in front left in it)",
                "not_available": why}, None
    return None, "; ".join(why)


def run_gpu(args):
    import torch
    import berserkerquake2_b200 as bq2
    from berserkerquake2_b200 import synth, host, shard, _capi as c

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if hasattr(os, "sched_setaffinity"):               # every rank gets its own PHYSICAL cores (both hyper-threads of each):
        mine = rank_cpus(local, world)                 # the frame loop's thread and the ctx's submit worker then do not
        try:                                           # share a core with each other or with another rank
            os.sched_setaffinity(0, mine)
        except OSError:
            pass
    threads = args.threads or max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 4)
    args.threads = threads

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    n_ents, lights, desc = WORKLOADS[args.workload]
    first, _ = host.shard_entities(n_ents * world, rank, world)           # weak scaling: n_ents per GPU
    ctx = bq2.Context(local)
    V, T = synth.sphere_dims(SLICES, RINGS)
    # Inputs AND outputs larger than L2 instead of a flush between steps: a ring of resident frames, each on its own
    # device copy of the models and with its own result arrays.
    l2_bytes = int(getattr(torch.cuda.get_device_properties(local), "L2_cache_size", 126 << 20))
    ring, step_input_bytes = ring_size(l2_bytes, n_ents, lights, V, T)
    R = build_ring(ctx, n_ents, lights, first, n_ents * world, ring, own_results=ring > 1, threads=threads)
    kept_frames, ring_indices = R["kept"], R["indices"]
    we, wl, pe, pl, n_pairs = R["scene0"]
    mids, model_bytes, setup_s = R["mids"], R["model_bytes"], R["setup_s"]
    result_bytes_per_frame = int(c.lib.bq2_resident_result_bytes(kept_frames[0]))
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))

    def steps_resident(k, start=0):
        ctx.run_ring(kept_frames, start % ring, k)

    clocks = ClockSampler(local) if rank == 0 else None
    steps_resident(max(args.warmup, 3))
    t_spin = time.perf_counter()                       # idle GPUs sit at ~120 MHz: keep stepping until the
    while time.perf_counter() - t_spin < args.spinup:  # clocks have ramped, then start the timed region
        steps_resident(ring)
        torch.cuda.synchronize()
    if clocks:
        clocks.mark()
    launches0 = ctx.kernel_launches
    ms_total_max, region_ms = timed_regions(ctx, kept_frames, args.steps, args.repeats, barrier, dev, dist, streams=args.streams)
    launches = (ctx.kernel_launches - launches0) // args.repeats
    ms_one_stream, region_ms_one = timed_regions(ctx, kept_frames, args.steps, args.repeats, barrier, dev, dist, streams=1)
    # A region of K steps pays its start and its drain once (the last launches have no successor to overlap with: ~15 us
    # at C2), which a short region (the driver's K = 20) spreads over few steps.  Reported BESIDE the line's value: the same
    # launches in regions of 200 steps -- the steady state an engine's frame loop runs in.
    k_long = max(200, args.steps)
    ms_long, _ = timed_regions(ctx, kept_frames, k_long, 5, barrier, dev, dist, streams=args.streams) if k_long != args.steps else (ms_total_max, None)
    med_r = int(np.argsort(region_ms)[len(region_ms) // 2])               # the repeat that is reported
    total_indices = sum(ring_indices[(med_r * args.steps + i) % ring] for i in range(args.steps))   # over its K steps
    if clocks:                                         # the timed regions are shorter than nvidia-smi's period:
        t_more = time.perf_counter()                   # keep the SAME steps running until it has seen a few samples
        while clocks.since_mark() < 4 and time.perf_counter() - t_more < 3.0:
            steps_resident(ring)
            torch.cuda.synchronize()
        clocks.freeze()
    clock_info = clocks.stop() if clocks else None     # (the sampler polls the driver: not during the wall-clock leg)
    # informational: ONE launch alone between two events, L2 flushed before it by a 256 MiB memset (dirty lines), or by
    # the memset followed by a read of a second buffer (clean lines) -- the latency of an isolated launch
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    flush2 = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
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
        """k frames through the public C ABI from HOST data, driven from Python (two ctypes calls per frame)."""
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

    e2e_run(max(args.warmup, 40))      # every frame state of the ring has captured its graph before the timed frames
    barrier()
    _, e2e_py_s = e2e_run(args.steps, timed=True)                  # the same loop driven from Python (ctypes per call)
    # the frame loop as an engine runs it: C on top of the public ABI (csrc/bq2_frame_loop.c), same calls, same depth
    loop = C.CDLL(os.path.join(ROOT, "berserkerquake2_b200", "libbq2loop.so"))
    loop.bq2_frame_loop.restype = C.c_int
    loop.bq2_frame_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    sec, tot64 = C.c_double(0), C.c_uint64(0)

    def e2e_loop(k, desc=None, depth=E2E_IN_FLIGHT):
        if loop.bq2_frame_loop(ctx._h, C.byref(desc or wd), k, depth, C.byref(sec), C.byref(tot64)) != 0:
            raise RuntimeError(c.lib.bq2_last_error(ctx._h))
        return int(tot64.value), float(sec.value)

    e2e_loop(max(args.warmup, 40))
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
    e2e_loop(max(args.warmup, 40), wd_pageable)                    # the same from pageable arrays (staged by the library)
    _, e2e_pageable_s = e2e_loop(args.steps, wd_pageable)
    e2e_by_depth = {}
    for d in (2, 4, 8):                                            # fewer frames in flight: less latency, lower rate
        e2e_loop(20, None, d)
        e2e_by_depth[d] = float(np.median([e2e_loop(args.steps, None, d)[1] for _ in range(9)]))
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

    # ---- the route that DELIVERS geometry to host memory: per frame bq2_world_submit + bq2_frame_wait, then the frame's
    # lerped rows, extruded rows and index buffers (every pair's first max-count words, one strided copy) to page-locked
    # host memory -- what a renderer that draws from host arrays (the reference ends in glDrawArrays over TrisVerts,
    # M:63778-63788) would consume.  Frames in flight as above; the copies of frame i run while frames i+1.. compute.
    cnt_max = int(max(ring_indices) / n_pairs * 1.35) + 64              # words per pair brought back (>= every count: checked)
    vrow = (V + 3) & ~3
    g_l = ctx.host_array(np.zeros(3 * n_ents * vrow, np.float32)); g_e = ctx.host_array(np.zeros(3 * n_pairs * vrow, np.float32))
    g_i = ctx.host_array(np.zeros(n_pairs * cnt_max, np.uint32))

    def geometry_frames(k):
        for _ in range(E2E_IN_FLIGHT - 1):
            e2e_submit()
        t0 = time.perf_counter()
        for _ in range(k):
            e2e_submit(); e2e_wait()
            for which, buf in ((c.A_LERPED, g_l), (c.A_EXTRUDED, g_e)):
                if c.lib.bq2_download(ctx._h, which, 0, buf.size, c.ptr(buf)) != 0:
                    raise RuntimeError(c.lib.bq2_last_error(ctx._h))
            if c.lib.bq2_download_indices(ctx._h, 0, n_pairs, cnt_max, c.ptr(g_i)) != 0:
                raise RuntimeError(c.lib.bq2_last_error(ctx._h))
        dt = time.perf_counter() - t0
        for _ in range(E2E_IN_FLIGHT - 1):
            e2e_wait()
        return dt
    geometry_frames(3)
    k_geo = max(5, min(args.steps, 40))
    e2e_geo_s = geometry_frames(k_geo) / k_geo
    counts_h = np.ctypeslib.as_array(c.lib.bq2_host_index_counts(ctx._h), (n_pairs,))
    assert int(counts_h.max()) <= cnt_max, "geometry leg: a pair has more indices than were brought back"
    geo_bytes = 4 * (g_l.size + g_e.size + g_i.size)

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
    del flush

    # ---- reduce: NCCL all_gather of the per-shard counters (the only collective of the path) ----------
    per_rank = shard.gather_counters([n_pairs * args.steps, n_ents * V * args.steps, total_indices // 3, total_indices,
                                      e2e_indices // 3], dev, dist)           # summed over the K timed steps
    g_pairs, g_verts, g_tris, g_indices, g_e2e_tris = [int(x) for x in per_rank.sum(0)]
    (e2e_ms_max,) = shard.max_over_ranks([e2e_s * 1e3], dev, dist)

    # ---- BASELINE configs[2] as written: 65,536 x 8 sharded entity-major over the N GPUs (strong scaling) ----------
    for h in kept_frames:
        ctx.free_kept(h)
    ctx.close()
    torch.cuda.empty_cache()
    c3 = None
    if not args.no_c3:
        c3 = c3_strong_leg(args, local, rank, world, barrier, dev, dist, l2_bytes)

    dropin = None
    if rank == 0 and world == 1 and not args.no_dropin:
        dropin = dropin_leg(args)

    if rank == 0:
        ms_per_step = ms_total_max / args.steps
        out = {
            "metric": METRIC,
            "value": g_tris / (ms_total_max * 1e-3), "unit": "tris/s",
            "lerped_verts_per_s": g_verts / (ms_total_max * 1e-3),
            "pairs_per_s": g_pairs / (ms_total_max * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload),
            "measurement": {
                "model_bytes_in_hbm_per_gpu": model_bytes,
                "mean_indices_per_pair": total_indices / args.steps / n_pairs,
                "l2": f"not flushed: inputs and outputs larger than L2 -- the K steps cycle through a ring of {ring} resident "
                      f"frames (scene steps 0..{ring - 1}), each on its own device copy of the {n_ents} models and, "
                      f"{'with' if result_bytes_per_frame else 'WITHOUT'} result arrays of its own: one cycle reads "
                      f"{ring * step_input_bytes / 1e6:.0f} MB of key frames, triangle tables and records and writes "
                      f"{ring * (min_traffic_bytes(n_ents, n_pairs, V, T, total_indices / args.steps) - step_input_bytes) / 1e6:.0f}"
                      f" MB of results to {ring} different sets of arrays (L2: {l2_bytes / 1e6:.0f} MB)",
                "resident_frames": ring, "input_bytes_per_step": step_input_bytes,
                "result_array_bytes_owned_per_frame": result_bytes_per_frame,
                "timed_regions_ms": region_ms, "repeats": args.repeats, "streams": args.streams,
                "ms_per_step_one_stream": ms_one_stream / args.steps, "timed_regions_ms_one_stream": region_ms_one,
                "timing": f"EXACTLY K = {args.steps} launches back to back from C (bq2_resident_run_streams) between one pair of "
                          f"CUDA events; the region is run {args.repeats} times with a barrier + synchronize in front of each, "
                          "each region's time is the max over ranks, the MEDIAN region is reported (all of them in "
                          f"timed_regions_ms).  The K steps are K INDEPENDENT frames (own inputs, own result arrays), dealt "
                          f"round-robin onto {args.streams} streams as the library runs frames in flight (bq2_world_submit gives "
                          "every frame in flight its own stream): the first event is recorded on stream 0, the other streams "
                          "wait for it, and they are joined into stream 0 before the second event, so every launch starts after "
                          "the first and ends before the second.  On each stream the launches also carry programmatic stream "
                          "serialization.  ms_per_step_one_stream = all K launches on ONE stream (round 1's way): there step N+1 "
                          "may only stage, lerp and build normals under the tail of step N and then waits for it "
                          "(griddepcontrol.wait); on two streams nothing orders it behind step N (profiles/r02_c2_pdl_timeline.md). "
                          "roofline.isolated_launch_ms = one launch alone between two events after an L2 flush"},
            "e2e": {"value": g_e2e_tris / args.steps / (e2e_ms_max * 1e-3 / args.steps), "unit": "tris/s",
                    "ms_per_step": e2e_ms_max / args.steps, "gpu_launches_per_step": launches_e2e / args.steps,
                    "repeats": e2e_repeats, "ms_per_step_min_max_rank0": e2e_spread,
                    "ms_per_step_python_driver": e2e_py_s * 1e3 / args.steps,
                    "ms_per_step_pageable_inputs": e2e_pageable_s * 1e3 / args.steps,
                    "ms_per_step_serial": e2e_serial_s * 1e3 / args.steps,
                    "frames_in_flight": E2E_IN_FLIGHT,
                    "ms_per_step_2_in_flight": e2e_by_depth[2] * 1e3 / args.steps,
                    "ms_per_step_4_in_flight": e2e_by_depth[4] * 1e3 / args.steps,
                    "ms_per_step_8_in_flight": e2e_by_depth[8] * 1e3 / args.steps,
                    "ms_per_step_host_built_records": e2e_host_s * 1e3 / args.steps,
                    "geometry_d2h": {"ms_per_step": e2e_geo_s * 1e3, "tris_per_s": e2e_indices / args.steps / 3 / e2e_geo_s,
                                     "d2h_bytes_per_step": geo_bytes, "d2h_gb_per_s": geo_bytes / e2e_geo_s / 1e9,
                                     "what": "the same frames, and every frame's lerped rows, extruded rows and index buffers "
                                             f"(the first {cnt_max} words of each pair's row, one strided copy) copied to "
                                             "page-locked host memory: the route for a renderer that draws from host arrays"},
                    "h2d_bytes_per_step": n_ents * 64 + 16 + len(wl_c) * 16 + n_pairs * 8 + h2d_basis_bytes, "d2h_bytes_per_step": 4 * ((n_pairs + 1) // 2 * 2 + 8),
                    "what": "per frame: bq2_world_submit (caller's thread: per-entity validation and numbering into one 64-byte row per "
                            "entity that also carries the caller's entity fields; the ctx's submit worker: AngleVectors of the rotated "
                            "entities (libm), staging of the caller's page-locked light and pair lists, the frame's graph launch; H2D of "
                            "the rows, bases, lights and pairs; the record-building kernel + the fused kernel; per-pair index counts + "
                            "totals stored to the page-locked result buffer by the frame's last kernel -- d2h_bytes_per_step, over "
                            "PCIe like a copy) + bq2_frame_wait, wall clock over K frames (median of %d such runs) with %d frames in "
                            "flight (each has its own records, result arrays and stream; the clock runs over K submit+wait iterations "
                            "of the full pipeline, K frames in and K frames out); the loop itself is C on the public ABI "
                            "(csrc/bq2_frame_loop.c), as an engine would run it; ms_per_step_python_driver = the same calls made from "
                            "Python; ms_per_step_N_in_flight = fewer frames in flight; ms_per_step_serial = one at a time; "
                            % (e2e_repeats, E2E_IN_FLIGHT) +
                            "vertex/index buffers stay in HBM for the draw"},
            "shared_model_variant": {
                "what": "SURVEY 8(d): the same frame with every entity instantiating ONE model (one launch alone after an "
                        "L2 flush); scored on its unique bytes (the model's frames and table counted once)",
                "ms_per_step": shared_ms, "tris_per_s": shared_indices / 3 / (shared_ms * 1e-3),
                "unique_bytes": min_traffic_bytes(n_ents, n_pairs, V, T, shared_indices) - (n_ents - 1) * (2 * 4 * V + 12 * T),
                "frac": (min_traffic_bytes(n_ents, n_pairs, V, T, shared_indices) - (n_ents - 1) * (2 * 4 * V + 12 * T))
                        / (shared_ms * 1e-3) / 1e9 / peaks()[0]},
            "gpu_launches": int(launches),
            "roofline": roofline_block(n_ents, n_pairs, V, T, total_indices / args.steps, ms_per_step),
            "clocks": clock_info,
            "setup_s": setup_s,
        }
        out["roofline"]["one_stream"] = {
            "kernel_ms": ms_one_stream / args.steps,
            "frac": out["roofline"]["bytes_per_launch"] / (ms_one_stream / args.steps * 1e-3) / 1e9 / out["roofline"]["peak"],
            "what": "the same K launches on ONE stream (programmatic dependent launch only): each step waits for the step in "
                    "front before its first global write"}
        out["roofline"]["long_region"] = {
            "steps": k_long, "kernel_ms": ms_long / k_long,
            "frac": out["roofline"]["bytes_per_launch"] / (ms_long / k_long * 1e-3) / 1e9 / out["roofline"]["peak"],
            "what": "the same back-to-back launches in regions of this many steps (median of 5): a region pays its start and "
                    "its drain once, the line's value (K = --steps) includes them"}
        out["roofline"]["kernel_ms_definition"] = ("timed region / K: the K launches are independent frames dealt onto "
                                                   f"{args.streams} streams, so consecutive launches overlap; one_stream and "
                                                   "isolated_launch_ms are the same kernel without that overlap")
        out["roofline"]["isolated_launch_ms"] = {
            "l2_flushed_by_memset": iso["dirty"], "l2_flushed_then_read": iso["clean"],
            "what": "ONE launch between two events on an otherwise idle stream: adds the launch and completion latency "
                    "that back-to-back steps hide"}
        if c3 is not None:
            out["c3_strong"] = c3
        if dropin is not None:
            out["e2e"]["dropin"] = dropin
        if world == 1 and not args.no_traffic:
            t, why = measure_traffic(args, os.path.join(ROOT, "gpurun_out"))
            if t is not None:
                out["roofline"]["traffic"] = t["bytes_per_step"]
                out["roofline"]["traffic_detail"] = t
                out["roofline"]["traffic_over_bytes"] = t["bytes_per_step"] / out["roofline"]["bytes_per_launch"]
            else:
                out["roofline"]["traffic_unmeasured"] = why
        if out["roofline"]["traffic"] is None:          # not measured in this run: the committed capture, marked as such
            prof = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(prof):
                try:
                    with open(prof) as f:
                        out["roofline"]["traffic"] = json.load(f).get(args.workload)
                        out["roofline"]["traffic_source"] = "profiles/traffic.json (a committed ncu capture, NOT this run)"
                except (OSError, ValueError):
                    pass
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_arm(WORKLOADS["c2"][0], WORKLOADS["c2"][1], steps=None, warmup=1,
                                          budget_s=args.cpu_seconds)
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def dropin_worker(args):
    """The reference's own R_DrawAliasShadowVolume (unmodified object code) with integration/bq2_dropin.c linked in front of
    it (oracle/_ref/libbq2ref_dropin.so): the three compute functions it calls are the CUDA-backed ones, a batch of one
    per call, results copied back into the engine's globals (TrisVerts filled)."""
    from berserkerquake2_b200 import synth, host
    so = os.path.join(ROOT, "oracle", "_ref", "libbq2ref_dropin.so")
    lib = C.CDLL(so); fn = lib.bq2ref_md2_world_bench; lib.bq2ref_init()
    fn.restype = C.c_longlong
    fn.argtypes = [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 8
    n_ents, lights = args.cpu_ents, args.cpu_lights
    we, wl, pe, pl = make_scene(n_ents, lights)
    n = len(pe)
    blobs = [synth.md2(SLICES, RINGS, FRAMES, e) for e in range(n_ents)]
    nb = host.md2_neighbours(blobs[0])
    bp = (C.c_void_p * n)(*[blobs[int(e)].ctypes.data for e in pe])
    np_ = (C.c_void_p * n)(*[nb.ctypes.data] * n)
    a = dict(frames=we["frame"][pe].astype(np.int32), oldframes=we["oldframe"][pe].astype(np.int32),
             backlerps=we["backlerp"][pe].astype(np.float32), origins=np.ascontiguousarray(we["origin"][pe]),
             oldorigins=np.ascontiguousarray(we["oldorigin"][pe]), angles=np.ascontiguousarray(we["angles"][pe]),
             lights=np.ascontiguousarray(wl["origin"][pl]), radii=np.ascontiguousarray(wl["radius"][pl]))
    P = lambda x: x.ctypes.data_as(C.c_void_p)
    call = lambda: fn(bp, np_, n, P(a["frames"]), P(a["oldframes"]), P(a["backlerps"]), P(a["origins"]),
                      P(a["oldorigins"]), P(a["angles"]), P(a["lights"]), P(a["radii"]))
    call(); call()                                       # models become resident, clocks ramp
    times = []
    for _ in range(5):
        t0 = time.perf_counter(); tot = call(); times.append(time.perf_counter() - t0)
    lib.bq2_dropin_kernel_launches.restype = C.c_int
    print(json.dumps({"pairs": n, "indices": int(tot), "s": float(np.median(times)), "s_min_max": [min(times), max(times)],
                      "launches": int(lib.bq2_dropin_kernel_launches())}), flush=True)


def dropin_leg(args):
    so = os.path.join(ROOT, "oracle", "_ref", "libbq2ref_dropin.so")
    if not os.path.exists(so):
        return {"unavailable": "oracle/_ref/libbq2ref_dropin.so not built (needs the reference tree at build time)"}
    n_ents, lights = 64, 4                               # a bounded sample of the C2 scene: its first 64 entities
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--_dropin_worker", "--cpu-ents", str(n_ents),
                            "--cpu-lights", str(lights)], capture_output=True, text=True, timeout=180)
        d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    except (subprocess.TimeoutExpired, OSError, IndexError, ValueError) as e:
        return {"unavailable": f"{type(e).__name__}"}
    return {"ms_per_pair": d["s"] * 1e3 / d["pairs"], "tris_per_s": d["indices"] / 3 / d["s"], "pairs_per_s": d["pairs"] / d["s"],
            "sample": f"the first {n_ents} entities x {lights} lights of the C2 scene, 5 passes, median; one process, one pair per call",
            "what": "the signature-keeping route: the reference's own, unmodified R_DrawAliasShadowVolume (M:63713) calls "
                    "R_CalcAliasFrameLerpShadow / R_CalcAliasFrameShadowVolume, which integration/bq2_dropin.c implements as a "
                    "batch of ONE on the C ABI (submit, wait, lerped / extruded / facing / TrisVerts copied back into the "
                    "engine's globals); the cost of one synchronous round trip per pair -- the batched routes above are the product"}


def c3_strong_leg(args, local, rank, world, barrier, dev, dist, l2_bytes):
    """BASELINE configs[2]: 65,536 entities x 8 lights, the global entity list sharded entity-major over the N ranks
    (rank r owns a contiguous block of entities with all their lights; nothing crosses devices)."""
    import berserkerquake2_b200 as bq2
    from berserkerquake2_b200 import host, shard
    E, Lt = args.c3_entities, 8
    first, count = host.shard_entities(E, rank, world)
    ctx = bq2.Context(local)
    V, T = 258, 512
    ring, step_input_bytes = ring_size(l2_bytes, count, Lt, V, T, cap=4)
    R = build_ring(ctx, count, Lt, first, E, ring, own_results=ring > 1, threads=args.threads)
    kept = R["kept"]
    n_pairs = R["scene0"][4]
    steps = args.steps
    ctx.run_ring(kept, 0, max(3, min(args.warmup, 5)))
    ms_max, region_ms = timed_regions(ctx, kept, steps, max(3, min(args.repeats, 5)), barrier, dev, dist, streams=args.streams)
    med_r = int(np.argsort(region_ms)[len(region_ms) // 2])
    idx = sum(R["indices"][(med_r * steps + i) % ring] for i in range(steps))
    per_rank = shard.gather_counters([n_pairs * steps, count * V * steps, idx // 3, idx, count], dev, dist)
    g_pairs, g_verts, g_tris, g_idx, g_ents = [int(x) for x in per_rank.sum(0)]
    for h in kept:
        ctx.free_kept(h)
    ctx.close()
    if rank != 0:
        return None
    ms_per_step = ms_max / steps
    # roofline of the slowest rank's launch ~ the mean rank's bytes (shards are equal to within one entity)
    rf = roofline_block(g_ents / world, g_pairs / steps / world, V, T, g_idx / steps / world, ms_per_step)
    return {"workload": f"BASELINE configs[2]: {E} synthetic MD2 entities (V=258,T=512,198 frames) x {Lt} lights, one distinct model "
                        f"per entity, sharded entity-major over {world} GPU(s): {g_ents // world} entities x {Lt} lights per GPU",
            "scaling": "strong", "n_gpus": world, "entities_total": g_ents, "lights_per_entity": Lt, "pairs_total": g_pairs // steps,
            "value": g_tris / (ms_max * 1e-3), "unit": "tris/s", "pairs_per_s": g_pairs / (ms_max * 1e-3),
            "lerped_verts_per_s": g_verts / (ms_max * 1e-3), "ms_per_step": ms_per_step, "steps": steps,
            "timed_regions_ms": region_ms, "resident_frames": ring, "streams": args.streams if ring > 1 else 1, "setup_s": R["setup_s"],
            "roofline": rf}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own code (oracle/_ref/libbq2ref.so) or, failing that, the oracle port.
# The reference is single-threaded and keeps its state in globals, so "all host cores" = one worker
# PROCESS per core, each owning a contiguous slice of the pair list.
def cpu_worker(args):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from berserkerquake2_b200 import synth, host
    n_ents, lights = args.cpu_ents, args.cpu_lights
    lo, hi = args.cpu_lo, args.cpu_hi                                      # pair range of this worker
    if args.cpu_pin >= 0 and hasattr(os, "sched_setaffinity"):
        try:
            os.sched_setaffinity(0, {args.cpu_pin})
        except OSError:
            pass
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
    """Each step = the whole batch (n_ents x lights pairs) once, split over one process per core, every process
    pinned to its own core.  A measurement = one batch of `passes` back-to-back passes on all workers at once; at
    least five batches are run (more until the budget is spent), the MEDIAN batch is reported with the min / max."""
    cpus = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
    cores = cores or len(cpus)
    n_pairs = n_ents * lights
    cores = max(1, min(cores, n_pairs))
    kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libbq2ref.so")) else "port"
    procs = []
    for w in range(cores):
        lo, hi = n_pairs * w // cores, n_pairs * (w + 1) // cores
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--_cpu_worker", "--cpu-ents", str(n_ents),
                                       "--cpu-lights", str(lights), "--cpu-lo", str(lo), "--cpu-hi", str(hi),
                                       "--cpu-pin", str(cpus[w % len(cpus)])],
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
    passes = steps if steps is not None else 20         # a FIXED number of passes per batch
    times, t_begin = [], time.perf_counter()
    while len(times) < 5 or (time.perf_counter() - t_begin < budget_s and len(times) < 400):
        dt, tot = run(passes)
        times.append(dt)
    for p in procs:
        p.stdin.write("quit\n"); p.stdin.flush(); p.wait(timeout=10)
    s_per_step = float(np.median(times))
    V, T = 258, 512
    return {"value": tot / 3 / s_per_step, "unit": "tris/s", "cores": cores, "kind": kind,
            "lerped_verts_per_s": n_pairs * V / s_per_step, "pairs_per_s": n_pairs / s_per_step,
            "ms_per_step": s_per_step * 1e3, "steps": passes, "batches": len(times),
            "ms_per_step_min_max": [min(times) * 1e3, max(times) * 1e3],
            "sample": f"{len(times)} batches of {passes} passes over the whole {n_ents} x {lights} batch ({n_pairs} pairs, "
                      f"{tot // 3} tris per pass), one process per core, each pinned to its core and looping the reference's "
                      f"R_DrawAliasShadowVolume over its slice, no synchronisation between passes (time per pass = the "
                      f"slowest process's total / passes); the median batch is reported, min / max beside it"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_ents, lights, desc = WORKLOADS["c2"]
    r = cpu_arm(n_ents, lights, steps=args.steps, warmup=max(args.warmup, 1), budget_s=min(args.cpu_seconds, 6.0))
    out = {"impl": "reference", "metric": METRIC,
           "value": r["value"], "unit": "tris/s", "lerped_verts_per_s": r["lerped_verts_per_s"],
           "pairs_per_s": r["pairs_per_s"], "n_gpus": args.gpus, "steps": r["steps"], "warmup": max(args.warmup, 1),
           "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": workload_config("c2"),
           "measurement": {"what": "the reference's CPU path (the whole R_DrawAliasShadowVolume per pair) on the host cores; the "
                                   "reference lerps once per PAIR, so lerped verts/s counts V per pair; the same batch whatever "
                                   "--gpus says (the reference has one process image per core, not per GPU)",
                           "batches": r["batches"], "ms_per_step_min_max": r["ms_per_step_min_max"]},
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
    ap.add_argument("--cpu-pin", type=int, default=-1, help=argparse.SUPPRESS)
    ap.add_argument("--repeats", type=int, default=11, help="device-timed regions of K steps each; the median is reported")
    ap.add_argument("--threads", type=int, default=0, help="host threads generating the synthetic models (0 = this rank's cores)")
    ap.add_argument("--streams", type=int, default=2, help="streams the K independent resident frames of the timed region are dealt onto")
    ap.add_argument("--no-c3", action="store_true", help="skip the strong-scaling leg (BASELINE configs[2])")
    ap.add_argument("--c3-entities", type=int, default=65536)
    ap.add_argument("--no-traffic", action="store_true", help="skip the in-run DRAM traffic measurement (ncu child)")
    ap.add_argument("--traffic-timeout", type=float, default=240.0)
    ap.add_argument("--_traffic_probe", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--_dropin_worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-dropin", action="store_true", help="skip the drop-in (signature-keeping) e2e leg")
    ap.add_argument("--probe-mode", default="range", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args._cpu_worker:
        return cpu_worker(args)
    if args._traffic_probe:
        return traffic_probe(args)
    if args._dropin_worker:
        return dropin_worker(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    main()
