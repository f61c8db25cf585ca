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
