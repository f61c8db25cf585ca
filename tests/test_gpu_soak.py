"""Soak test of the frame-state ring: a few hundred frames of several shapes submitted at random pipeline depths, mixed
with synchronous frames, host-built frames, replays of kept frames and model uploads in between -- every waited frame
must return exactly what the same scene returns when computed alone (counts, totals, and a sampled pair's arrays)."""
import numpy as np
import pytest

import berserkerquake2_b200 as bq2
from scenes import Md2Model, world_scene, bits

pytestmark = pytest.mark.gpu


def test_ring_soak(oracle):
    rng = np.random.RandomState(2024)
    with bq2.Context(0) as ctx:
        models = [Md2Model(oracle, s, r, frames=6, seed=500 + i) for i, (s, r) in enumerate([(8, 5), (16, 17), (12, 9), (62, 34)])]
        mids = [m.upload(ctx) for m in models]
        want = bq2.WANT_SHADOW | bq2.WANT_TABLES
        scenes = []
        for i, (n, L) in enumerate([(5, 1), (40, 3), (130, 2), (17, 9), (64, 4), (1, 12)]):
            we, wl, pe, pl = world_scene(n, L, 6, n_models=4, seed=3000 + i)
            we["model"] = np.array(mids)[we["model"]]
            if i == 3:
                we["angles"][::2] = rng.uniform(-180, 180, (len(we[::2]), 3))
            res = ctx.world_frame(we, wl, pe, pl)
            p = int(rng.randint(0, len(pe)))
            m = models[mids.index(int(we["model"][pe[p]]))]
            ref = dict(counts=res.index_counts.copy(), total=int(res.total_indices), p=p, V=m.V, T=m.T, e=int(pe[p]),
                       idx=res.indices(p).copy(), ext=bits(res.extruded(p, m.V)).copy(), lerp=bits(res.lerped(int(pe[p]), m.V)).copy(),
                       face=res.facing(p, m.T).copy())
            d = ctx.world_desc(we, wl, pe, pl, None, want, 0); d._keep = ctx._keep
            ents, pairs = ctx.build_records(we, wl, pe, pl)
            scenes.append((d, ref, ents, pairs))

        def check(res, k):
            ref = scenes[k][1]
            assert np.array_equal(res.index_counts, ref["counts"]) and int(res.total_indices) == ref["total"], k
            if rng.rand() < 0.3:                                     # the device arrays of the sampled pair
                assert np.array_equal(res.indices(ref["p"]), ref["idx"]), k
                assert np.array_equal(bits(res.extruded(ref["p"], ref["V"])), ref["ext"]), k
                assert np.array_equal(bits(res.lerped(ref["e"], ref["V"])), ref["lerp"]), k
                assert np.array_equal(res.facing(ref["p"], ref["T"]), ref["face"]), k

        in_flight = []                                               # scene indices, oldest first
        kept = None
        for step in range(700):
            action = rng.rand()
            if action < 0.55 and len(in_flight) < 16:                 # submit (world path)
                k = int(rng.randint(0, len(scenes)))
                ctx.world_submit(scenes[k][0]); in_flight.append(k)
            elif action < 0.62 and len(in_flight) < 16:               # submit (host-built records)
                k = int(rng.randint(0, len(scenes)))
                ctx.submit(scenes[k][2], scenes[k][3]); in_flight.append(k)
            elif action < 0.90 and in_flight:                        # wait for the oldest
                check(ctx.wait(), in_flight.pop(0))
            elif action < 0.93 and not in_flight:                    # a synchronous frame in between
                k = int(rng.randint(0, len(scenes)))
                fo = bq2._capi.FrameOut()
                import ctypes as C
                assert bq2._capi.lib.bq2_world_frame(ctx._h, C.byref(scenes[k][0]), C.byref(fo)) == 0
                assert int(fo.total_indices) == scenes[k][1]["total"]
            elif action < 0.95 and not in_flight:                    # keep a host-built frame, replay it later
                k = int(rng.randint(0, len(scenes)))
                ctx.frame(scenes[k][2], scenes[k][3])
                if kept is not None:
                    ctx.free_kept(kept[0])
                kept = (ctx.keep(), k)
            elif action < 0.97 and not in_flight and kept is not None:
                ctx.replay_kept(kept[0]); ctx.wait()
                got = ctx.download(bq2._capi.A_INDEX_COUNT, 0, len(scenes[kept[1]][1]["counts"]), np.uint32)
                assert np.array_equal(got, scenes[kept[1]][1]["counts"])
            elif action < 0.985:                                     # a model arrives while frames are in flight
                Md2Model(oracle, 6, 4, frames=2, seed=900 + step).upload(ctx)
                if kept is not None:                                 # (the mesh table moved: a kept frame is stale now)
                    ctx.free_kept(kept[0]); kept = None
        while in_flight:
            check(ctx.wait(), in_flight.pop(0))
        if kept is not None:
            ctx.free_kept(kept[0])
