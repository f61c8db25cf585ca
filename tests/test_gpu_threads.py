"""The engine-style multi-GPU contract: ONE process, one context + one host thread per device -- no
torch.distributed, no torchrun.  N threads each create their own context (on the box's GPUs round-robin; a 1-GPU box puts
them all on device 0), upload their shard's models and run world frames concurrently; every thread's results must equal
the reference hashes of its own shard."""
import threading

import numpy as np
import pytest

import berserkerquake2_b200 as bq2
import oracle_lib
from berserkerquake2_b200 import host
from test_gpu_exhaustive import bench_scene

pytestmark = pytest.mark.gpu


def test_contexts_on_threads_in_one_process(oracle):
    import torch
    n_dev = max(1, torch.cuda.device_count())
    n_threads = max(2, min(4, n_dev))
    E, Lt = 96, 3
    blobs, nb, we, wl, pe, pl = bench_scene(E, Lt)
    want, _ = oracle_lib.reference_pair_hashes(blobs, [nb] * E, we, wl, pe, pl)
    errors, done = [], []

    def worker(rank):
        try:
            first, count = host.shard_entities(E, rank, n_threads)          # entity-major shard of this thread
            with bq2.Context(rank % n_dev) as ctx:
                mids = np.array([ctx.upload_md2(blobs[first + e], nb) for e in range(count)], np.int32)
                we_s = we[first:first + count].copy(); we_s["model"] = mids
                keep = (pe >= first) & (pe < first + count)
                pe_s = (pe[keep] - first).astype(np.int32); pl_s = pl[keep]
                for it in range(25):                                       # concurrently with the other threads
                    if it % 2:
                        ctx.world_frame(we_s, wl, pe_s, pl_s)
                    else:
                        ents, pairs = ctx.build_records(we_s, wl, pe_s, pl_s)
                        ctx.frame(ents, pairs)
                    got = ctx.hash_results(len(pe_s))
                    if not np.array_equal(got, want[keep]):
                        raise AssertionError(f"thread {rank} iteration {it}: {int((got != want[keep]).sum())} pairs differ")
                assert ctx.kernel_launches > 0
            done.append(rank)
        except Exception as e:                                              # noqa: BLE001 (reported by the main thread)
            errors.append((rank, repr(e)))

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(n_threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errors, errors
    assert sorted(done) == list(range(n_threads))


def test_frame_loop_in_c_through_the_submit_worker(oracle):
    """The engine-style frame loop (csrc/bq2_frame_loop.c on the public ABI): eight frames in flight from page-locked caller
    arrays, so that the ctx's submit worker stages the arrays and launches the frames while the caller's thread runs the
    per-entity loop of the next ones.  Every frame's total must be the reference's, the last frame's pairs all match, and
    mixing in calls that must first meet the worker (a download, a synchronous frame, an upload) keeps everything exact."""
    import ctypes as C
    import os
    from berserkerquake2_b200 import _capi as c
    E, Lt = 200, 4
    blobs, nb, we, wl, pe, pl = bench_scene(E, Lt)
    want, total = oracle_lib.reference_pair_hashes(blobs, [nb] * E, we, wl, pe, pl)
    loop = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(bq2.__file__)), "libbq2loop.so"))
    loop.bq2_frame_loop.restype = C.c_int
    loop.bq2_frame_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    sec, tot = C.c_double(0), C.c_uint64(0)
    with bq2.Context(0) as ctx:
        we["model"] = np.array([ctx.upload_md2(b, nb) for b in blobs], np.int32)
        pin = [ctx.host_array(x) for x in (we, wl, pe, pl)]
        wd = ctx.world_desc(*pin, None, bq2.WANT_SHADOW, 0)
        depth = c.lib.bq2_max_frames_in_flight()
        for k, d in ((300, depth), (50, 3), (40, 1), (300, depth)):
            assert loop.bq2_frame_loop(ctx._h, C.byref(wd), k, d, C.byref(sec), C.byref(tot)) == 0, c.lib.bq2_last_error(ctx._h)
            assert tot.value == k * total, (k, d)
            assert np.array_equal(ctx.hash_results(len(pe)), want)          # meets the worker, then hashes the last frame
            cnt = ctx.download(c.A_INDEX_COUNT, 0, len(pe), np.uint32)
            assert int(cnt.sum()) == total
            ctx.upload_md2(blobs[0], nb)                                    # an upload between bursts
        # pageable caller arrays: staged by the caller's thread, launched by the worker
        wd2 = ctx.world_desc(we, wl, pe, pl, None, bq2.WANT_SHADOW, 0)
        assert loop.bq2_frame_loop(ctx._h, C.byref(wd2), 200, depth, C.byref(sec), C.byref(tot)) == 0
        assert tot.value == 200 * total and np.array_equal(ctx.hash_results(len(pe)), want)
        # an error inside a pipelined frame surfaces at its wait, the frames around it are unharmed
        bad_pe = pe.copy(); bad_pe[5] = E + 7
        wd_bad = ctx.world_desc(we, wl, bad_pe, pl, None, bq2.WANT_SHADOW, 0); keep_bad = ctx._keep
        fo = c.FrameOut()
        seq = [wd2, wd2, wd_bad, wd2, wd2]
        for d in seq:
            assert c.lib.bq2_world_submit(ctx._h, C.byref(d)) == 0
        rcs = [c.lib.bq2_frame_wait(ctx._h, C.byref(fo)) for _ in seq]
        assert rcs == [0, 0, c.E_INVALID, 0, 0]
        assert fo.total_indices == total
