"""Full-size parity, EVERY pair: BASELINE configs 2 (1,024 x 4), 3 (one GPU's shard: 8,192 x 8) and 4 (MD3, 4 surfaces of
2,048 triangles) against the unmodified reference run over all host cores.  The device reduces everything it produced
for a pair slot -- lerped row, extruded row, facing words, index count, the expanded vertex stream -- to one 64-bit hash
(bq2_hash_results); oracle/ref/ref_driver.c computes the same hash from the engine's globals after the whole
R_DrawAliasShadowVolume ran for that pair (tests/test_hash_definition.py pins the two definitions to each other).
Both routes (records built on the host, records built on the device) are checked."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import berserkerquake2_b200 as bq2
import oracle_lib
from berserkerquake2_b200 import synth, host, _capi as c_api
from scenes import Md3Model, world_scene, bits

pytestmark = pytest.mark.gpu
SLICES, RINGS, FRAMES = 16, 17, 198                     # bench.py's model: V = 258, T = 512, 198 frames


def bench_scene(n_ents, lights, step=0):
    """exactly bench.py's scene: one distinct model per entity, SURVEY 8(d) entity / light recipe"""
    with ThreadPoolExecutor(max_workers=max(1, len(os.sched_getaffinity(0)))) as pool:
        blobs = list(pool.map(lambda e: synth.md2(SLICES, RINGS, FRAMES, e), range(n_ents)))
    nb = host.md2_neighbours(blobs[0])
    we, wl, pe, pl = world_scene(n_ents, lights, FRAMES, n_models=n_ents, seed=12345, step=step)
    return blobs, nb, we, wl, pe, pl


def every_pair(ctx, blobs, nb, we, wl, pe, pl):
    mids = np.array([ctx.upload_md2(b, nb) for b in blobs], np.int32)
    we_d = we.copy(); we_d["model"] = mids[we["model"]]
    want, total = oracle_lib.reference_pair_hashes(blobs, [nb] * len(blobs), we, wl, pe, pl)
    assert len(set(want.tolist())) == len(pe)                     # (no two pairs alike: a slot mix-up cannot hide)
    res_w = ctx.world_frame(we_d, wl, pe, pl)                      # records built on the device
    got_w = ctx.hash_results(len(pe))
    assert int(res_w.total_indices) == total and res_w.overflow_pairs == 0
    bad = np.nonzero(got_w != want)[0]
    assert len(bad) == 0, f"world route: {len(bad)} of {len(pe)} pairs differ from the reference, first {bad[:5]}"
    ents, pairs = ctx.build_records(we_d, wl, pe, pl)             # records built on the host
    res_h = ctx.frame(ents, pairs)
    got_h = ctx.hash_results(len(pe))
    assert int(res_h.total_indices) == total
    bad = np.nonzero(got_h != want)[0]
    assert len(bad) == 0, f"host-built route: {len(bad)} of {len(pe)} pairs differ from the reference, first {bad[:5]}"
    return want, ents, pairs


def test_c2_every_pair_matches_the_reference(oracle):
    with bq2.Context(0) as ctx:
        want, ents, pairs = every_pair(ctx, *bench_scene(1024, 4))
        assert len(want) == 4096
        # kept frames with their OWN result arrays (bench.py's ring): replays launched from C reproduce every pair,
        # each frame into its own arrays
        ring = []
        for step in range(3):
            ctx.frame(ents, pairs)
            ring.append(ctx.keep(own_results=True))
            assert c_api.lib.bq2_resident_result_bytes(ring[-1]) > 0
        ms = ctx.run_ring(ring, 1, 7, timed=True)
        assert ms > 0
        for h in ring:
            assert np.array_equal(ctx.hash_results(len(want), h), want)
        # the same frames dealt onto two and three streams (independent frames in flight): still every pair
        for streams in (2, 3):
            assert ctx.run_ring(ring, 0, 11, timed=True, streams=streams) > 0
            for h in ring:
                assert np.array_equal(ctx.hash_results(len(want), h), want), streams
        # one strided copy brings the index rows of a frame back: equal to the per-pair downloads
        res = ctx.frame(ents, pairs)
        words = int(res.index_counts.max())
        packed = np.zeros((len(want), words), np.uint32)
        assert c_api.lib.bq2_download_indices(ctx._h, 0, len(want), words, c_api.ptr(packed)) == 0
        for p in (0, 1, 777, len(want) - 1):
            n = int(res.index_counts[p])
            assert np.array_equal(packed[p, :n], res.indices(p))
        assert c_api.lib.bq2_download_indices(ctx._h, 0, len(want) + 1, words, c_api.ptr(packed)) < 0      # outside the frame
        # a frame kept WITHOUT own arrays shares the state's: still the same results after the ring ran
        ctx.frame(ents, pairs)
        shared = ctx.keep()
        assert c_api.lib.bq2_resident_result_bytes(shared) == 0
        ctx.run_ring([shared], 0, 3)
        assert np.array_equal(ctx.hash_results(len(want), shared), want)
        arr = (__import__("ctypes").c_void_p * 2)(shared.value, ring[0].value)                           # several streams need frames
        assert c_api.lib.bq2_resident_run_streams(ctx._h, arr, 2, 0, 4, 2, None) < 0                      # that own their result arrays
        cnt = ctx.download_kept(ring[0], c_api.A_INDEX_COUNT, 0, len(want), np.uint32)
        assert int(cnt.sum()) > 0 and (cnt % 6 == 0).all()
        for h in ring + [shared]:
            ctx.free_kept(h)


def test_c3_shard_every_pair_matches_the_reference(oracle):
    with bq2.Context(0) as ctx:
        want, _, _ = every_pair(ctx, *bench_scene(8192, 8))
        assert len(want) == 65536


def test_c4_md3_every_slot_matches_the_reference(oracle, ref):
    """BASELINE config 4 at bench_c4's size: 256 entities x 4 lights on a 4-surface MD3 (V = 1,026, T = 2,048 each)."""
    md3 = Md3Model(oracle, 32, 33, frames=8, num_meshes=4, seed=3)
    with bq2.Context(0) as ctx:
        mid = ctx.upload_md3(md3.frame_translate, md3.meshes)
        we, wl, pe, pl = world_scene(256, 4, 8, seed=4242)
        we["model"] = mid
        ents, pairs = ctx.build_records(we, wl, pe, pl)
        res = ctx.frame(ents, pairs)
        assert res.n_pair_slots == 4 * len(pe)
        got = ctx.hash_results(res.n_pair_slots)
        # team-kernel frames kept with their own result arrays AND scratch rows, dealt onto two streams: same hashes
        ring = []
        for _ in range(2):
            ctx.frame(ents, pairs)
            ring.append(ctx.keep(own_results=True))
        assert ctx.run_ring(ring, 0, 6, timed=True, streams=2) > 0
        for h in ring:
            assert np.array_equal(ctx.hash_results(res.n_pair_slots, h), got)
            ctx.free_kept(h)
        for p in range(len(pe)):
            W = we[int(pe[p])]
            drawn, out = ref.md3_shadow(md3.frame_translate, md3.meshes, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]),
                                        W["origin"], W["oldorigin"], W["angles"], wl["origin"][pl[p]], float(wl["radius"][pl[p]]))
            for k, o in enumerate(out):
                assert got[4 * p + k] == oracle_lib.pair_hash(o["lerped"], o["extruded"], o["viss"], o["tris_verts"]), (p, k)


def test_hash_sees_a_single_wrong_word(oracle):
    """the device hash is sensitive: a frame with a different light differs in every pair"""
    blobs, nb, we, wl, pe, pl = bench_scene(8, 2)
    with bq2.Context(0) as ctx:
        mids = np.array([ctx.upload_md2(b, nb) for b in blobs], np.int32)
        we["model"] = mids
        ctx.world_frame(we, wl, pe, pl)
        a = ctx.hash_results(len(pe))
        wl2 = wl.copy(); wl2["origin"][:, 0] += np.float32(2 ** -10)
        ctx.world_frame(we, wl2, pe, pl)
        b = ctx.hash_results(len(pe))
        assert (a != b).all() and (a != 0).all()
