"""GPU parity: the CUDA path, called through the C ABI (libbq2gpu.so), against the CPU oracle on the
same seeded inputs.  Bit-exact everywhere (the kernels use the reference's operation order with
round-to-nearest, no FMA), which is stronger than north_star's 1e-5 relative for positions."""
import numpy as np
import pytest

import berserkerquake2_b200 as bq2
from berserkerquake2_b200 import synth as synth_mod
from scenes import Md2Model, Md3Model, world_scene, bits

pytestmark = pytest.mark.gpu


def check_md2_pairs(ctx, o, models, mids, we, wl, pe, pl, which=None):
    ents, pairs = ctx.build_records(we, wl, pe, pl)
    assert len(pairs) == len(pe)
    res = ctx.frame(ents, pairs)
    assert res.overflow_pairs == 0
    sel = range(len(pe)) if which is None else which
    total = 0
    for p in sel:
        e = int(pe[p]); W = we[e]; m = models[mids.index(int(W["model"]))]
        w = o.md2_pair(m.blob, m.nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                       W["oldorigin"], W["angles"], wl["origin"][pl[p]], float(wl["radius"][pl[p]]), idx=m.idx)
        assert np.array_equal(bits(res.lerped(e, m.V)), bits(w["lerped"])), f"lerped pair {p}"
        used = np.unique(m.idx)               # the reference only writes extrudedVerts of referenced verts
        assert np.array_equal(bits(res.extruded(p, m.V))[used], bits(w["extruded"])[used]), f"extruded pair {p}"
        assert np.array_equal(res.facing(p, m.T), w["viss"]), f"facing pair {p}"
        assert int(res.index_counts[p]) == len(w["indices"]), f"TrisCounter pair {p}"
        assert np.array_equal(res.indices(p), w["indices"]), f"indices pair {p}"
        total += len(w["indices"])
    return res, total


def test_c1_single_md2_one_light(gpu_ctx, oracle):
    """BASELINE config 1: one MD2 (V=258, T=512, 198 frames) x 1 point light."""
    m = Md2Model(oracle)
    mid = m.upload(gpu_ctx)
    for step in range(4):
        we, wl, pe, pl = world_scene(1, 1, m.F, step=step * 37)
        we["model"] = mid
        res, _ = check_md2_pairs(gpu_ctx, oracle, [m], [mid], we, wl, pe, pl)
        w = oracle.md2_pair(m.blob, m.nb, int(we["frame"][0]), int(we["oldframe"][0]), float(we["backlerp"][0]),
                            we["origin"][0], we["oldorigin"][0], we["angles"][0], wl["origin"][0], 300.0)
        assert np.array_equal(bits(res.trisverts(0)), bits(w["tris_verts"]))      # TrisVerts stream itself


def test_md2_batch_many_models(gpu_ctx, oracle):
    """64 entities x 4 lights over 8 distinct models, rotated entities included."""
    models = [Md2Model(oracle, frames=12, seed=s) for s in range(8)]
    mids = [m.upload(gpu_ctx) for m in models]
    we, wl, pe, pl = world_scene(64, 4, 12, n_models=8, seed=777)
    we["model"] = np.array(mids)[we["model"]]
    assert (we["angles"][:, 1] != 0).any()
    res, total = check_md2_pairs(gpu_ctx, oracle, models, mids, we, wl, pe, pl)
    assert int(res.total_indices) == total
    assert int(res.total_lerped_verts) == 64 * models[0].V


def test_md2_mixed_sizes_and_pair_order(gpu_ctx, oracle):
    """Different V/T in one launch, pairs given in light-major (the renderer's) order, entities
    with zero lights, and one light shared by many entities."""
    shapes = [(8, 5), (16, 17), (24, 13), (62, 34), (3, 2)]
    models = [Md2Model(oracle, s, r, frames=6, seed=10 + i) for i, (s, r) in enumerate(shapes)]
    assert models[3].V == 2048 and models[3].T == 4092 and models[4].T == 6    # MAX_VERTS; a double tetrahedron
    mids = [m.upload(gpu_ctx) for m in models]
    n = 20
    we, wl, _, _ = world_scene(n, 3, 6, n_models=len(models), seed=4242)
    we["model"] = np.array(mids)[we["model"]]
    pe, pl = [], []
    for l in range(3):                          # light-major, entity 7 and 13 unlit, light 0 hits everyone
        for e in range(n):
            if e in (7, 13):
                continue
            pe.append(e); pl.append(0 if l == 0 else e * 3 + l)
    pe, pl = np.array(pe, np.int32), np.array(pl, np.int32)
    res, _ = check_md2_pairs(gpu_ctx, oracle, models, mids, we, wl, pe, pl)
    # unlit entities are still lerped
    for e in (7, 13):
        m = models[mids.index(int(we["model"][e]))]
        w = oracle.md2_lerp(m.blob, int(we["frame"][e]), int(we["oldframe"][e]), float(we["backlerp"][e]),
                            we["origin"][e], we["oldorigin"][e], we["angles"][e])
        assert np.array_equal(bits(res.lerped(e, m.V)), bits(w))


def test_md2_open_mesh_and_degenerate(gpu_ctx, oracle):
    """Open edges (neighbour -1), an ambiguous 3-triangle edge (dup rule) and a zero-area triangle whose
    NaN normal must count as facing (M:63620-63626)."""
    m = Md2Model(oracle, 8, 5, frames=3, seed=5)
    from berserkerquake2_b200 import synth
    tris = synth.md2_tris(m.blob)
    tris[3, 0:3] = tris[3, [0, 0, 2]]           # degenerate: repeated vertex -> zero normal -> NaN
    tris[10, 0:3] = tris[11, 0:3]               # duplicate triangle -> edges shared by 3 tris
    tris[20, 0:3] = (0, 1, 1)
    m.idx = __import__("oracle_lib").md2_idx(m.blob)
    m.nb = oracle.md2_neighbours(m.blob)
    assert (m.nb == -1).any()
    mid = m.upload(gpu_ctx)
    we, wl, pe, pl = world_scene(6, 2, 3, seed=99)
    we["model"] = mid
    res, _ = check_md2_pairs(gpu_ctx, oracle, [m], [mid], we, wl, pe, pl)
    assert res.facing(0, m.T)[3] == 1


def test_md2_shell_lerp(gpu_ctx, oracle):
    """RF_SHELL_* push-out along the current frame's anorm (M:63565-63572); shells cast no volume
    (M:63725), so only the lerp is compared and the pair is filtered on the host."""
    m = Md2Model(oracle, frames=5, seed=2)
    mid = m.upload(gpu_ctx)
    we, wl, pe, pl = world_scene(4, 1, 5, seed=5)
    we["model"] = mid
    we["rf_flags"][1] = 0x400                        # RF_SHELL_RED
    we["rf_flags"][2] = 0x800 | 0x4                  # RF_SHELL_GREEN | RF_WEAPONMODEL -> half scale
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl)
    assert len(pairs) == 2 and list(pairs["entity"]) == [0, 3]
    res = gpu_ctx.frame(ents, pairs)
    for e in range(4):
        shell = bool(we["rf_flags"][e] & 0x1c00)
        scale = 0.5 if we["rf_flags"][e] & 4 else 1.0
        move, fv, bv, _ = oracle.md2_lerp_consts(m.blob, int(we["frame"][e]), int(we["oldframe"][e]),
                                                 float(we["backlerp"][e]), we["origin"][e], we["oldorigin"][e],
                                                 we["angles"][e])
        w = np.zeros((m.V, 3), np.float32)
        import oracle_lib as ol
        oracle.lib.bq2o_md2_lerp(ol.P(m.blob), int(we["frame"][e]), int(we["oldframe"][e]), ol.P(move), ol.P(fv),
                                 ol.P(bv), int(shell), scale if shell else 0.0, ol.P(w))
        assert np.array_equal(bits(res.lerped(e, m.V)), bits(w)), f"entity {e}"


def test_md2_lerp_only(gpu_ctx, oracle):
    """R_CalcAliasFrameLerp alone (ambient pass, M:50303): no pairs."""
    m = Md2Model(oracle, frames=7, seed=8)
    mid = m.upload(gpu_ctx)
    we, wl, pe, pl = world_scene(9, 1, 7, seed=31)
    we["model"] = mid
    ents, _ = gpu_ctx.build_records(we, wl, pe[:0], pl[:0])
    res = gpu_ctx.frame(ents, None)
    assert res.n_pair_slots == 0
    for e in range(9):
        w = oracle.md2_lerp(m.blob, int(we["frame"][e]), int(we["oldframe"][e]), float(we["backlerp"][e]),
                            we["origin"][e], we["oldorigin"][e], we["angles"][e])
        assert np.array_equal(bits(res.lerped(e, m.V)), bits(w))


def test_index_stride_overflow_reports_exact_count(gpu_ctx, oracle):
    m = Md2Model(oracle, frames=3, seed=1)
    mid = m.upload(gpu_ctx)
    we, wl, pe, pl = world_scene(2, 1, 3, seed=3)
    we["model"] = mid
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl)
    full = gpu_ctx.frame(ents, pairs)
    small = gpu_ctx.frame(ents, pairs, index_stride=600, allow_overflow=True)
    assert small.status == -5 and small.overflow_pairs == 2
    assert np.array_equal(small.index_counts, full.index_counts)          # counts stay exact
    for p in range(2):
        assert np.array_equal(small.indices(p), full.indices(p)[:600])   # prefix written, nothing beyond


def test_empty_frame_and_bad_arguments(gpu_ctx, oracle):
    res = gpu_ctx.frame(np.zeros(0, bq2.ENTITY_DTYPE), None)
    assert res.n_ent_slots == 0 and res.total_indices == 0
    e = np.zeros(1, bq2.ENTITY_DTYPE); e["model"] = 10 ** 6
    with pytest.raises(bq2.Bq2Error):
        gpu_ctx.frame(e, None)
    with pytest.raises(bq2.Bq2Error):
        gpu_ctx.upload_md2(np.zeros(100, np.uint8), None)


def test_c4_md3_multisurface(gpu_ctx, oracle):
    """BASELINE config 4: MD3, 4 surfaces x (V=1026, T=2048) = 8192 tris, one surface's skin casts no
    shadow (M:63862-63880)."""
    md3 = Md3Model(oracle, 32, 33, frames=6, num_meshes=4, seed=1, casts=[1, 1, 0, 1])
    mid = md3.upload(gpu_ctx)
    we, wl, pe, pl = world_scene(3, 2, 6, seed=2024)
    we["model"] = mid
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl)
    res = gpu_ctx.frame(ents, pairs)
    assert res.n_ent_slots == 12 and res.n_pair_slots == 24
    for p in range(len(pe)):
        e = int(pe[p]); W = we[e]
        for k, mesh in enumerate(md3.meshes):
            es, ps = 4 * e + k, 4 * p + k
            V, T = mesh["verts"].shape[1], len(mesh["indexes"])
            w = oracle.md3_mesh_pair(mesh["verts"], mesh["indexes"], mesh["neighbours"], md3.frame_translate,
                                     int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                                     W["oldorigin"], W["angles"], wl["origin"][pl[p]], 300.0)
            assert np.array_equal(bits(res.lerped(es, V)), bits(w["lerped"])), (p, k)
            if not mesh["casts_shadow"]:
                assert int(res.index_counts[ps]) == 0
                continue
            assert np.array_equal(bits(res.extruded(ps, V)), bits(w["extruded"])), (p, k)
            assert np.array_equal(res.facing(ps, T), w["viss"]), (p, k)
            assert np.array_equal(res.indices(ps), w["indices"]), (p, k)
            assert np.array_equal(bits(res.trisverts(ps)), bits(w["tris_verts"])), (p, k)


@pytest.mark.parametrize("smooth", [True, False])
def test_md2_tangent_space_basis(gpu_ctx, oracle, smooth):
    """a10/a11: lerped N/T/B (per vertex when RF_SMOOTHVECS, else per triangle) and LightP / tsH."""
    m = Md2Model(oracle, 12, 9, frames=5, seed=4, tangent_space=True, smooth=smooth)
    mid = m.upload(gpu_ctx)
    we, wl, pe, pl = world_scene(5, 2, 5, seed=11)
    we["model"] = mid
    view = np.array([100.0, -50.0, 80.0], np.float32)
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl, view_world=view)
    want = bq2.WANT_SHADOW | bq2.WANT_NTB | bq2.WANT_TSLIGHT
    res = gpu_ctx.frame(ents, pairs, want=want)
    from berserkerquake2_b200 import synth
    for e in range(5):
        f, of = int(we["frame"][e]), int(we["oldframe"][e])
        bl = float(we["backlerp"][e]); fl = float(ents["frontlerp"][e])
        if smooth:
            nc, no = synth.md2_frame_verts(m.blob, f)[:, 3], synth.md2_frame_verts(m.blob, of)[:, 3]
            rows = m.V
        else:
            nc, no = m.normals[f], m.normals[of]
            rows = m.T
        N, T_, B = oracle.ntb_rows(nc, no, m.tangents[f], m.tangents[of], m.binormals[f], m.binormals[of], bl, fl)
        gN, gT, gB = res.ntb(e, rows)
        assert np.array_equal(bits(gN), bits(N)) and np.array_equal(bits(gT), bits(T_)) and np.array_equal(bits(gB), bits(B))
        P = oracle.md2_lerp(m.blob, f, of, bl, we["origin"][e], we["oldorigin"][e], we["angles"][e])
        corner = m.idx.reshape(-1)
        for l in range(2):
            p = 2 * e + l
            if smooth:
                lp, ts = oracle.tslight(P, N, T_, B, pairs["light"][p], pairs["view"][p])
                gl, gt = res.tslight(p, m.V)
            else:                           # per corner, the triangle's basis (M:64972-64986)
                row = np.repeat(np.arange(m.T), 3)
                lp, ts = oracle.tslight(P[corner], N[row], T_[row], B[row], pairs["light"][p], pairs["view"][p])
                gl, gt = res.tslight(p, 3 * m.T)
            assert np.array_equal(bits(gl), bits(lp)) and np.array_equal(bits(gt), bits(ts))


def test_md3_tangent_space_basis(gpu_ctx, oracle):
    md3 = Md3Model(oracle, 16, 9, frames=4, num_meshes=2, seed=3)
    mid = md3.upload(gpu_ctx)
    we, wl, pe, pl = world_scene(2, 1, 4, seed=8)
    we["model"] = mid
    view = np.array([10.0, 20.0, 30.0], np.float32)
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl, view_world=view)
    res = gpu_ctx.frame(ents, pairs, want=bq2.WANT_SHADOW | bq2.WANT_NTB | bq2.WANT_TSLIGHT)
    for e in range(2):
        f, of = int(we["frame"][e]), int(we["oldframe"][e])
        bl = float(we["backlerp"][e]); fl = float(ents["frontlerp"][e])
        for k, mesh in enumerate(md3.meshes):
            v = mesh["verts"]; V = v.shape[1]
            N, T_, B = oracle.ntb_rows(v["normal"][f], v["normal"][of], v["tangent"][f], v["tangent"][of],
                                       v["binormal"][f], v["binormal"][of], bl, fl)
            gN, gT, gB = res.ntb(2 * e + k, V)
            assert np.array_equal(bits(gN), bits(N)) and np.array_equal(bits(gT), bits(T_)) and np.array_equal(bits(gB), bits(B))
            P = res.lerped(2 * e + k, V)
            lp, ts = oracle.tslight(P, N, T_, B, pairs["light"][e], pairs["view"][e])
            gl, gt = res.tslight(2 * e + k, V)
            assert np.array_equal(bits(gl), bits(lp)) and np.array_equal(bits(gt), bits(ts))


def test_c2_full_size_properties(gpu_ctx, oracle):
    """BASELINE config 2 at full size (1,024 entities x 4 lights, one distinct model per 16 entities):
    size-independent properties + oracle spot checks on a sample of pairs."""
    n_models = 64
    models = [Md2Model(oracle, frames=16, seed=100 + s) for s in range(n_models)]
    mids = [m.upload(gpu_ctx) for m in models]
    we, wl, pe, pl = world_scene(1024, 4, 16, n_models=n_models, seed=12345)
    we["model"] = np.array(mids)[we["model"]]
    sample = list(range(0, 4096, 97))
    res, _ = check_md2_pairs(gpu_ctx, oracle, models, mids, we, wl, pe, pl, which=sample)
    cnt = res.index_counts
    assert (cnt % 6 == 0).all() and (cnt > 0).all()
    assert int(res.total_indices) == int(cnt.sum())
    # closed 2-manifold: every silhouette edge borders exactly one facing tri, so the silhouette is a set of
    # closed loops; caps = 6 * facing, sides = 6 * silhouette edges
    T, V = models[0].T, models[0].V
    for p in sample[:8]:
        f = res.facing(p, T)
        ind = res.indices(p)
        nf = int(f.sum())
        sides = (len(ind) - 6 * nf) // 6
        quads = ind[:6 * sides].reshape(-1, 6)
        assert (quads[:, 1] == quads[:, 0] + V).all() and (quads[:, 5] == quads[:, 0]).all()
        assert (quads[:, 2] == quads[:, 3]).all() and (quads[:, 3] == quads[:, 4] + V).all()
        starts, ends = np.sort(quads[:, 0]), np.sort(quads[:, 4])
        assert np.array_equal(starts, ends)                       # loops: each vertex entered once, left once
        caps = ind[6 * sides:].reshape(-1, 6)
        assert (caps[:, 3:] == caps[:, 2::-1] + V).all()
    # replay (inputs resident) is idempotent
    before = res.indices(sample[3]).copy()
    gpu_ctx.replay()
    gpu_ctx.wait()
    from berserkerquake2_b200 import _capi
    cnt2 = gpu_ctx.download(_capi.A_INDEX_COUNT, 0, len(cnt), np.uint32)
    assert np.array_equal(cnt2, cnt) and np.array_equal(res.indices(sample[3]), before)


def test_c3_shard_full_size_properties(oracle):
    """One GPU's share of BASELINE config 3 at full size (8,192 entities x 8 lights = 65,536 pairs, 1.1e8 index words):
    the world path and the host-built path agree on every count, the totals are the sum of the counts, a checksum of
    the whole index / extruded arrays is the same for both routes and for a second run, every silhouette of the sampled
    pairs is a set of closed loops, and a sample of pairs matches the oracle."""
    n_models = 48
    with bq2.Context(0) as ctx:
        models = [Md2Model(oracle, frames=8, seed=700 + s) for s in range(n_models)]
        mids = [m.upload(ctx) for m in models]
        n, L = 8192, 8
        we, wl, pe, pl = world_scene(n, L, 8, n_models=n_models, seed=777)
        we["model"] = np.array(mids)[we["model"]]
        T, V = models[0].T, models[0].V
        stride = 12 * T

        def checksums(res):
            cnt = res.index_counts.copy()
            idx = ctx.download(c_api.A_INDICES, 0, len(pe) * stride, np.uint32).reshape(len(pe), stride)
            wgt = np.arange(stride, dtype=np.uint64) % 251 + 1
            h_idx = 0
            for lo in range(0, len(pe), 4096):                           # position-weighted sum of the valid words
                blk = idx[lo:lo + 4096]
                mask = np.arange(stride)[None, :] < cnt[lo:lo + 4096, None]
                h_idx += int(((blk * mask).astype(np.uint64) @ wgt).sum())
            return cnt, h_idx, idx

        from berserkerquake2_b200 import _capi as c_api
        res_w = ctx.world_frame(we, wl, pe, pl)
        cnt_w, h_w, idx_w = checksums(res_w)
        assert res_w.overflow_pairs == 0 and int(res_w.total_indices) == int(cnt_w.sum())
        assert (cnt_w % 6 == 0).all() and (cnt_w > 0).all()
        ext_w = bits(ctx.download(c_api.A_EXTRUDED, 0, 3 * len(pe) * ((V + 3) & ~3), np.float32))
        res_w2 = ctx.world_frame(we, wl, pe, pl)                         # again (graph path): idempotent
        cnt_w2, h_w2, _ = checksums(res_w2)
        assert np.array_equal(cnt_w, cnt_w2) and h_w == h_w2
        ents, pairs = ctx.build_records(we, wl, pe, pl)
        res_h = ctx.frame(ents, pairs)
        cnt_h, h_h, _ = checksums(res_h)
        assert np.array_equal(cnt_w, cnt_h) and h_w == h_h               # the two routes write the same index words
        ext_h = bits(ctx.download(c_api.A_EXTRUDED, 0, 3 * len(pe) * ((V + 3) & ~3), np.float32))
        rows_w = ext_w.reshape(len(pe), -1)[:, :3 * V]; rows_h = ext_h.reshape(len(pe), -1)[:, :3 * V]
        assert np.array_equal(rows_w, rows_h)
        sample = list(range(0, len(pe), 4099))
        for p in sample:
            e = int(pe[p]); W = we[e]; m = models[mids.index(int(W["model"]))]
            w = oracle.md2_pair(m.blob, m.nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                                W["oldorigin"], W["angles"], wl["origin"][pl[p]], float(wl["radius"][pl[p]]), idx=m.idx)
            ind = idx_w[p, :cnt_w[p]]
            assert np.array_equal(ind, w["indices"]), p
            nf = int(w["viss"].sum()); sides = (len(ind) - 6 * nf) // 6
            quads = ind[:6 * sides].reshape(-1, 6)
            assert np.array_equal(np.sort(quads[:, 0]), np.sort(quads[:, 4]))      # closed loops


# ------------------------------------------------------------------------------------------------------
# bq2_world_frame: per-pair records built on the device
def check_world(ctx, o, models, mids, we, wl, pe, pl, view=None):
    res = ctx.world_frame(we, wl, pe, pl, view_world=view)
    assert res.n_pair_slots == len(pe) and res.overflow_pairs == 0
    total = 0
    for p in range(len(pe)):
        e = int(pe[p]); W = we[e]; m = models[mids.index(int(W["model"]))]
        import berserkerquake2_b200.host as host
        if not host.casts_shadow(int(W["rf_flags"])) or int(W["rf_flags"]) & (0x80000 | 0x10000000):
            assert int(res.index_counts[p]) == 0, p
            continue
        w = o.md2_pair(m.blob, m.nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                       W["oldorigin"], W["angles"], wl["origin"][pl[p]], float(wl["radius"][pl[p]]), idx=m.idx)
        used = np.unique(m.idx)
        assert np.array_equal(bits(res.lerped(e, m.V)), bits(w["lerped"])), p
        assert np.array_equal(bits(res.extruded(p, m.V))[used], bits(w["extruded"])[used]), p
        assert np.array_equal(res.facing(p, m.T), w["viss"]), p
        assert np.array_equal(res.indices(p), w["indices"]), p
        total += len(w["indices"])
    assert int(res.total_indices) == total
    return res


def test_world_frame_matches_oracle(gpu_ctx, oracle):
    """Mixed mesh sizes (both kernels), rotated entities, light-major pair order, a shared light, filtered
    entities (shell / translucent / nocastshadow) that keep their slot with count 0."""
    shapes = [(8, 5), (16, 17), (24, 13), (62, 34), (3, 2)]
    models = [Md2Model(oracle, s, r, frames=6, seed=40 + i) for i, (s, r) in enumerate(shapes)]
    mids = [m.upload(gpu_ctx) for m in models]
    n = 24
    we, wl, _, _ = world_scene(n, 3, 6, n_models=len(models), seed=999)
    we["model"] = np.array(mids)[we["model"]]
    we["angles"][5] = (10, 20, 30); we["angles"][6] = (0, 0, 90)
    we["rf_flags"][2] = 0x400; we["rf_flags"][9] = 0x20; we["rf_flags"][11] = 0x80000
    pe, pl = [], []
    for l in range(3):
        for e in range(n):
            if e in (7, 13):
                continue
            pe.append(e); pl.append(0 if l == 0 else e * 3 + l)
    pe, pl = np.array(pe, np.int32), np.array(pl, np.int32)
    res = check_world(gpu_ctx, oracle, models, mids, we, wl, pe, pl)
    for e in (2, 7, 13):                                     # shell entity / unlit entities are still lerped
        m = models[mids.index(int(we["model"][e]))]
        shell = bool(we["rf_flags"][e] & 0x1c00)
        w = oracle.md2_lerp(m.blob, int(we["frame"][e]), int(we["oldframe"][e]), float(we["backlerp"][e]),
                            we["origin"][e], we["oldorigin"][e], we["angles"][e], shell=shell, shell_scale=1.0 if shell else 0.0)
        assert np.array_equal(bits(res.lerped(e, m.V)), bits(w)), e


def test_world_frame_equals_host_built_frame_at_c2_size(gpu_ctx, oracle):
    """Full C2 size: device-built records give exactly the arrays of bq2_build_records + bq2_frame."""
    n_models = 32
    models = [Md2Model(oracle, frames=8, seed=300 + s) for s in range(n_models)]
    mids = [m.upload(gpu_ctx) for m in models]
    we, wl, pe, pl = world_scene(1024, 4, 8, n_models=n_models, seed=2468)
    we["model"] = np.array(mids)[we["model"]]
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl)
    a = gpu_ctx.frame(ents, pairs)
    V, T = models[0].V, models[0].T
    sample = list(range(0, 4096, 211))
    keep = [(a.extruded(p, V).copy(), a.facing(p, T).copy(), a.indices(p).copy()) for p in sample]
    cnt_a, tot_a = a.index_counts.copy(), int(a.total_indices)
    b = gpu_ctx.world_frame(we, wl, pe, pl)
    assert np.array_equal(b.index_counts, cnt_a) and int(b.total_indices) == tot_a
    assert np.array_equal(b.pair_vert_offset[:4096], a.pair_vert_offset[:4096])
    for p, (ext, f, ind) in zip(sample, keep):
        assert np.array_equal(bits(b.extruded(p, V)), bits(ext)) and np.array_equal(b.facing(p, T), f)
        assert np.array_equal(b.indices(p), ind)
    # replay keeps the totals exact (they are accumulated by the kernels)
    gpu_ctx.replay(); r = gpu_ctx.wait()
    assert int(r.total_indices) == tot_a


def test_world_frame_bad_pair_is_reported(gpu_ctx, oracle):
    m = Md2Model(oracle, 8, 5, frames=2, seed=1)
    mid = m.upload(gpu_ctx)
    we, wl, pe, pl = world_scene(2, 1, 2, seed=1)
    we["model"] = mid
    pe = np.array([0, 5, 1], np.int32); pl = np.array([0, 0, 9], np.int32)      # entity 5 and light 9 do not exist
    with pytest.raises(bq2.Bq2Error) as e:
        gpu_ctx.world_frame(we, wl, pe, pl)
    assert e.value.status == -1
    ok = gpu_ctx.world_frame(we, wl, pe[:1], pl[:1])
    assert int(ok.index_counts[0]) > 0


def test_world_frame_md3_and_tangent_space_take_the_host_route(gpu_ctx, oracle):
    md3 = Md3Model(oracle, 12, 9, frames=4, num_meshes=2, seed=6)
    mid = md3.upload(gpu_ctx)
    we, wl, pe, pl = world_scene(3, 2, 4, seed=17)
    we["model"] = mid
    we["rf_flags"][1] = 0x20                                  # translucent: its two pairs keep their slots, count 0
    res = gpu_ctx.world_frame(we, wl, pe, pl, want=bq2.WANT_SHADOW | bq2.WANT_TABLES)
    assert res.n_pair_slots == 12
    for p in range(6):
        e = int(pe[p]); W = we[e]
        for k, mesh in enumerate(md3.meshes):
            ps = 2 * p + k
            if e == 1:
                assert int(res.index_counts[ps]) == 0
                continue
            w = oracle.md3_mesh_pair(mesh["verts"], mesh["indexes"], mesh["neighbours"], md3.frame_translate,
                                     int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                                     W["oldorigin"], W["angles"], wl["origin"][pl[p]], 300.0)
            assert np.array_equal(res.indices(ps), w["indices"]), (p, k)
            assert np.array_equal(bits(res.extruded(ps, mesh["verts"].shape[1])), bits(w["extruded"]))


def test_two_frames_in_flight(gpu_ctx, oracle):
    """bq2_world_submit for frame i+1 before bq2_frame_wait for frame i: waits complete the OLDEST frame and
    every frame's counts/totals are its own; one submit more than the library keeps in flight is refused."""
    models = [Md2Model(oracle, frames=6, seed=70 + s) for s in range(4)]
    mids = [m.upload(gpu_ctx) for m in models]
    import ctypes as C
    from berserkerquake2_b200 import _capi as c
    scenes, serial = [], []
    depth = c.lib.bq2_max_frames_in_flight()
    assert depth == 16
    for i in range(depth + 1):
        we, wl, pe, pl = world_scene(40 + 7 * i, 3, 6, n_models=4, seed=500 + i)
        we["model"] = np.array(mids)[we["model"]]
        r = gpu_ctx.world_frame(we, wl, pe, pl)
        serial.append((int(r.total_indices), r.index_counts.copy()))
        scenes.append(gpu_ctx.world_desc(we, wl, pe, pl, None, bq2.WANT_SHADOW, 0))
        scenes[-1]._keep = gpu_ctx._keep
    fo = c.FrameOut()
    got = []

    def waited():
        assert c.lib.bq2_frame_wait(gpu_ctx._h, C.byref(fo)) == 0
        got.append((int(fo.total_indices), np.ctypeslib.as_array(c.lib.bq2_host_index_counts(gpu_ctx._h), (fo.n_pair_slots,)).copy()))
    for i in range(depth):
        assert c.lib.bq2_world_submit(gpu_ctx._h, C.byref(scenes[i])) == 0
    assert c.lib.bq2_world_submit(gpu_ctx._h, C.byref(scenes[depth])) == -1      # only `depth` frames in flight
    waited()                                                                      # the OLDEST frame
    assert c.lib.bq2_world_submit(gpu_ctx._h, C.byref(scenes[depth])) == 0
    for _ in range(depth):
        waited()
    assert len(got) == depth + 1
    for (ta, ca), (tb, cb) in zip(serial, got):
        assert ta == tb and np.array_equal(ca, cb)


def test_host_built_frames_in_flight_and_host_alloc_edges(gpu_ctx, oracle):
    """bq2_frame_submit (records built by the caller) through the same ring: four frames of different shapes in flight,
    each wait returns the oldest frame's own counts and device arrays.  bq2_host_alloc / bq2_host_free edge cases."""
    import ctypes as C
    from berserkerquake2_b200 import _capi as c
    models = [Md2Model(oracle, s, r, frames=5, seed=140 + i) for i, (s, r) in enumerate([(8, 5), (12, 9)])]
    mids = [m.upload(gpu_ctx) for m in models]
    frames = []
    for i in range(6):
        we, wl, pe, pl = world_scene(11 + 5 * i, 2, 5, n_models=2, seed=900 + i)
        we["model"] = np.array(mids)[we["model"]]
        ents, pairs = gpu_ctx.build_records(we, wl, pe, pl)
        r = gpu_ctx.frame(ents, pairs)
        frames.append((ents, pairs, r.index_counts.copy(), [r.indices(p).copy() for p in range(r.n_pair_slots)]))
    for i in range(3):
        gpu_ctx.submit(frames[i][0], frames[i][1])
    for i in range(3, 6 + 3):
        if i < 6:
            gpu_ctx.submit(frames[i][0], frames[i][1])
        r = gpu_ctx.wait()
        _, _, cnt, idx = frames[i - 3]
        assert np.array_equal(r.index_counts, cnt), i
        for p in range(r.n_pair_slots):
            assert np.array_equal(r.indices(p), idx[p]), (i, p)
    assert not c.lib.bq2_host_alloc(gpu_ctx._h, 0)                       # nothing to allocate
    assert not c.lib.bq2_host_alloc(None, 64)
    c.lib.bq2_host_free(gpu_ctx._h, None)                                # no-ops
    c.lib.bq2_host_free(gpu_ctx._h, C.c_void_p(0x1000))
    p = c.lib.bq2_host_alloc(gpu_ctx._h, 4096)
    assert p
    c.lib.bq2_host_free(gpu_ctx._h, p)
    c.lib.bq2_host_free(gpu_ctx._h, p)                                   # already gone: ignored


def test_kept_frames_replay_bit_identically(gpu_ctx, oracle):
    """bq2_frame_keep / bq2_resident_replay (bench.py's ring of resident frames): re-running a kept frame reproduces
    its results exactly, in any order, with other frames computed in between; world-path frames cannot be kept."""
    import ctypes as C
    from berserkerquake2_b200 import _capi as c
    models = [Md2Model(oracle, s, r, frames=5, seed=160 + i) for i, (s, r) in enumerate([(8, 5), (16, 17)])]
    mids = [m.upload(gpu_ctx) for m in models]
    kept = []
    for i in range(3):
        we, wl, pe, pl = world_scene(40, 2, 5, n_models=2, seed=1000 + i)       # same shape: the result arrays are reused
        we["model"] = np.array(mids)[we["model"]]
        ents, pairs = gpu_ctx.build_records(we, wl, pe, pl)
        r = gpu_ctx.frame(ents, pairs)
        snap = (r.index_counts.copy(), [r.indices(p).copy() for p in range(r.n_pair_slots)],
                [bits(r.extruded(p, models[mids.index(int(we["model"][pe[p]]))].V)).copy() for p in range(len(pe))],
                int(r.index_stride))
        kept.append((gpu_ctx.keep(), snap))
    for order in ((2, 0, 1), (1, 1, 0, 2)):
        for k in order:
            h, (cnt, idx, ext, stride) = kept[k]
            gpu_ctx.replay_kept(h)
            gpu_ctx.wait()                                                       # nothing pending: a stream sync
            got = gpu_ctx.download(c.A_INDEX_COUNT, 0, len(cnt), np.uint32)
            assert np.array_equal(got, cnt), k
            for p in range(len(cnt)):
                assert np.array_equal(gpu_ctx.download(c.A_INDICES, p * stride, cnt[p], np.uint32), idx[p]), (k, p)
    for h, _ in kept:
        gpu_ctx.free_kept(h)
    we, wl, pe, pl = world_scene(10, 2, 5, n_models=2, seed=7)
    we["model"] = np.array(mids)[we["model"]]
    gpu_ctx.world_frame(we, wl, pe, pl)
    h = C.c_void_p()
    assert c.lib.bq2_frame_keep(gpu_ctx._h, C.byref(h)) == c.E_INVALID and not h


def test_overlapped_frames_keep_their_own_device_results(gpu_ctx, oracle):
    """Frame i+1 is submitted (and may run) before frame i is waited for: the vertex / facing / index arrays that
    bq2_frame_wait returns for frame i are still frame i's (each of the two frame states owns its result arrays),
    for alternating shapes, over several rounds (graph replays included)."""
    models = [Md2Model(oracle, s, r, frames=6, seed=90 + i) for i, (s, r) in enumerate([(8, 5), (16, 17), (62, 34)])]
    mids = [m.upload(gpu_ctx) for m in models]
    want = bq2.WANT_SHADOW | bq2.WANT_TABLES
    scenes = []
    for i in range(2):
        we, wl, pe, pl = world_scene(30 + 11 * i, 2 + i, 6, n_models=3, seed=800 + i)
        we["model"] = np.array(mids)[we["model"]]
        scenes.append((we, wl, pe, pl))

    def snapshot(res, we, pe):
        out = []
        for p in range(len(pe)):
            m = models[mids.index(int(we["model"][pe[p]]))]
            out.append((bits(res.lerped(int(pe[p]), m.V)).copy(), bits(res.extruded(p, m.V)).copy(),
                        res.facing(p, m.T).copy(), res.indices(p).copy()))
        return out

    serial = []
    for we, wl, pe, pl in scenes:
        res = check_world(gpu_ctx, oracle, models, mids, we, wl, pe, pl)          # oracle-checked, one at a time
        serial.append(snapshot(res, we, pe))
    descs = []
    for we, wl, pe, pl in scenes:
        d = gpu_ctx.world_desc(we, wl, pe, pl, None, want, 0); d._keep = gpu_ctx._keep
        descs.append(d)
    gpu_ctx.world_submit(descs[0]); gpu_ctx.world_submit(descs[1])
    n_rounds = 11
    for rnd in range(2, n_rounds + 2):
        if rnd < n_rounds:
            gpu_ctx.world_submit(descs[rnd & 1])                                 # frame rnd goes in (third in flight) ...
        res = gpu_ctx.wait()                                                     # ... frame rnd-2 comes out
        k = (rnd - 2) & 1
        got = snapshot(res, scenes[k][0], scenes[k][2])
        for p, (a, b) in enumerate(zip(serial[k], got)):
            for x, y in zip(a, b):
                assert np.array_equal(x, y), (rnd, p)


def test_world_frame_from_page_locked_caller_arrays(gpu_ctx, oracle):
    """Arrays from bq2_host_alloc are copied to the device from where they are (no staging copy, five copy nodes in
    the captured graph): same results as pageable arrays, new contents are seen on every replay, and a descriptor that
    mixes pinned and pageable arrays falls back to staging.  (The library takes this route from 1 MiB of input per
    frame; BQ2_DIRECT_MIN=0, read at bq2_create, makes the small scene take it.)"""
    import os
    os.environ["BQ2_DIRECT_MIN"] = "0"
    try:
        own_ctx = bq2.Context(0)
    finally:
        del os.environ["BQ2_DIRECT_MIN"]
    with own_ctx as gpu_ctx:
        _pinned_arrays_body(gpu_ctx, oracle)


def _pinned_arrays_body(gpu_ctx, oracle):
    models = [Md2Model(oracle, s, r, frames=6, seed=120 + i) for i, (s, r) in enumerate([(8, 5), (16, 17)])]
    mids = [m.upload(gpu_ctx) for m in models]
    we, wl, pe, pl = world_scene(37, 3, 6, n_models=2, seed=4242)
    we["model"] = np.array(mids)[we["model"]]
    we["angles"][3] = (5, 50, -20)
    pwe, pwl, ppe, ppl = (gpu_ctx.host_array(a) for a in (we, wl, pe, pl))
    for rnd in range(4):                                         # rounds 2+ replay the captured graph
        we["frame"] = (we["frame"] + rnd) % 6; we["backlerp"] = np.float32(0.125 * (rnd + 1))
        wl["origin"][:, 2] += np.float32(3.0 * rnd)
        pwe[...] = we; pwl[...] = wl
        ref = check_world(gpu_ctx, oracle, models, mids, we, wl, pe, pl)         # pageable, oracle-checked
        ref_counts = ref.index_counts.copy()
        ref_idx = [ref.indices(p).copy() for p in range(len(pe))]
        ref_ext = [bits(ref.extruded(p, models[mids.index(int(we["model"][pe[p]]))].V)).copy() for p in range(len(pe))]
        for arrays in ((pwe, pwl, ppe, ppl), (pwe, wl, ppe, ppl)):               # all pinned / mixed
            res = gpu_ctx.world_frame(*arrays)
            assert np.array_equal(res.index_counts, ref_counts) and int(res.total_indices) == int(ref.total_indices)
            for p in range(len(pe)):
                m = models[mids.index(int(we["model"][pe[p]]))]
                assert np.array_equal(res.indices(p), ref_idx[p]), (rnd, p)
                assert np.array_equal(bits(res.extruded(p, m.V)), ref_ext[p]), (rnd, p)


def test_world_frame_single_mesh_md3_and_many_lights(gpu_ctx, oracle):
    """Device-built records with a one-surface MD3 (float vertices in the warp kernel) next to MD2s, and an entity lit
    by more lights than the per-entity bucket of the fast record-building path holds (the library must fall back)."""
    md3 = Md3Model(oracle, 12, 9, frames=4, num_meshes=1, seed=12)
    m2 = Md2Model(oracle, 10, 7, frames=4, seed=12)
    id3, id2 = md3.upload(gpu_ctx), m2.upload(gpu_ctx)
    we, wl, _, _ = world_scene(6, 24, 4, seed=61)
    we["model"] = [id3, id2, id3, id2, id2, id3]
    pe, pl = [], []
    for l in range(24):                       # entity 1 and 2 get all 24 lights, the others 3
        for e in range(6):
            if e in (1, 2) or l < 3:
                pe.append(e); pl.append(e * 24 + l)
    pe, pl = np.array(pe, np.int32), np.array(pl, np.int32)
    res = gpu_ctx.world_frame(we, wl, pe, pl)
    assert res.n_pair_slots == len(pe)
    mesh = md3.meshes[0]
    for p in range(len(pe)):
        e = int(pe[p]); W = we[e]
        if int(W["model"]) == id3:
            w = oracle.md3_mesh_pair(mesh["verts"], mesh["indexes"], mesh["neighbours"], md3.frame_translate, int(W["frame"]),
                                     int(W["oldframe"]), float(W["backlerp"]), W["origin"], W["oldorigin"], W["angles"],
                                     wl["origin"][pl[p]], 300.0)
            V, T = mesh["verts"].shape[1], len(mesh["indexes"])
        else:
            w = oracle.md2_pair(m2.blob, m2.nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                                W["oldorigin"], W["angles"], wl["origin"][pl[p]], 300.0, idx=m2.idx)
            V, T = m2.V, m2.T
        assert np.array_equal(bits(res.lerped(e, V)), bits(w["lerped"])), p
        assert np.array_equal(bits(res.extruded(p, V)), bits(w["extruded"])), p
        assert np.array_equal(res.facing(p, T), w["viss"]) and np.array_equal(res.indices(p), w["indices"]), p


def test_world_frame_graph_replay_sees_new_data(gpu_ctx, oracle):
    """Same-shape frames replay a captured CUDA graph from the third frame on: the graph must pick up each
    frame's NEW records (different frames, lights, view) from the pinned buffers it was captured with."""
    m = Md2Model(oracle, 10, 7, frames=6, seed=77)
    mid = m.upload(gpu_ctx)
    for it in range(7):
        we, wl, pe, pl = world_scene(12, 3, 6, seed=900 + it)        # same shape, different content
        we["model"] = mid
        res = gpu_ctx.world_frame(we, wl, pe, pl, want=bq2.WANT_SHADOW)      # no tables: the graph-eligible route
        tot = 0
        for p in range(len(pe)):
            W = we[pe[p]]
            w = oracle.md2_pair(m.blob, m.nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                                W["oldorigin"], W["angles"], wl["origin"][pl[p]], 300.0, idx=m.idx)
            assert int(res.index_counts[p]) == len(w["indices"]), (it, p)
            got = gpu_ctx.download(_capi_module().A_INDICES, p * int(res.index_stride), len(w["indices"]), np.uint32)
            assert np.array_equal(got, w["indices"]), (it, p)
            tot += len(w["indices"])
        assert int(res.total_indices) == tot, it


def _capi_module():
    from berserkerquake2_b200 import _capi
    return _capi


def test_world_frame_degenerate_shapes(gpu_ctx, oracle):
    """no pairs, no entities, an entity nobody lights, one pair."""
    m = Md2Model(oracle, 8, 5, frames=3, seed=2)
    mid = m.upload(gpu_ctx)
    we, wl, pe, pl = world_scene(3, 1, 3, seed=5)
    we["model"] = mid
    r = gpu_ctx.world_frame(we, wl, pe[:0], pl[:0])                       # lerp only
    assert r.n_pair_slots == 0 and int(r.total_indices) == 0 and int(r.total_lerped_verts) == 3 * m.V
    for e in range(3):
        w = oracle.md2_lerp(m.blob, int(we["frame"][e]), int(we["oldframe"][e]), float(we["backlerp"][e]),
                            we["origin"][e], we["oldorigin"][e], we["angles"][e])
        assert np.array_equal(bits(r.lerped(e, m.V)), bits(w))
    r = gpu_ctx.world_frame(we[:0], wl, pe[:0], pl[:0])                   # nothing at all
    assert r.n_ent_slots == 0 and r.n_pair_slots == 0
    r = gpu_ctx.world_frame(we, wl, pe[1:2], pl[1:2])                     # one pair; entities 0 and 2 unlit
    w = oracle.md2_pair(m.blob, m.nb, int(we["frame"][1]), int(we["oldframe"][1]), float(we["backlerp"][1]),
                        we["origin"][1], we["oldorigin"][1], we["angles"][1], wl["origin"][1], 300.0, idx=m.idx)
    assert np.array_equal(r.indices(0), w["indices"])
    from berserkerquake2_b200 import _capi as c
    assert gpu_ctx.brush_volumes(np.zeros(0, c.BRUSH_PAIR_DTYPE), np.zeros(0, np.uint8)) == []


def test_kernel_class_boundary(gpu_ctx, oracle):
    """T = 1,024 is the last mesh the warp kernel takes (32 ballot words per warp), T = 1,088 the first for the team
    kernel; V = 1,022 / 1,026 straddle the vertex limit."""
    shapes = [(16, 33), (17, 33), (30, 35), (32, 33)]
    models = [Md2Model(oracle, s, r, frames=3, seed=20 + i) for i, (s, r) in enumerate(shapes)]
    assert [m.T for m in models] == [1024, 1088, 2040, 2048] and [m.V for m in models] == [514, 546, 1022, 1026]
    mids = [m.upload(gpu_ctx) for m in models]
    we, wl, pe, pl = world_scene(8, 3, 3, n_models=4, seed=31)
    we["model"] = np.array(mids)[we["model"]]
    check_md2_pairs(gpu_ctx, oracle, models, mids, we, wl, pe, pl)
    check_world(gpu_ctx, oracle, models, mids, we, wl, pe, pl)


def test_every_edge_a_silhouette(gpu_ctx, oracle):
    """Neighbour tables with no neighbours at all (r_simple-style / fully open meshes): every edge of every facing
    triangle is a silhouette edge, S = 3F -- far more than either kernel stages in shared memory, so the
    write-through paths carry the frame.  Both kernel classes, host-built and device-built records, and a
    truncating index_stride on the team kernel."""
    shapes = [(16, 17), (16, 33), (32, 33), (62, 34)]
    models = [Md2Model(oracle, s, r, frames=3, seed=50 + i) for i, (s, r) in enumerate(shapes)]
    for m in models:
        m.nb = np.full_like(m.nb, -1)
    mids = [m.upload(gpu_ctx) for m in models]
    we, wl, pe, pl = world_scene(8, 2, 3, n_models=4, seed=77)
    we["model"] = np.array(mids)[we["model"]]
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl)
    stride = 24 * 4092
    res = gpu_ctx.frame(ents, pairs, index_stride=stride)
    res_w = gpu_ctx.world_frame(we, wl, pe, pl, index_stride=stride)
    want = []
    for p in range(len(pe)):
        W = we[pe[p]]; m = models[mids.index(int(W["model"]))]
        w = oracle.md2_pair(m.blob, m.nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                            W["oldorigin"], W["angles"], wl["origin"][pl[p]], 300.0, idx=m.idx)
        assert len(w["indices"]) == 24 * int(w["viss"].sum())
        want.append(w["indices"])
    # FrameResult objects share the device arrays: compare the later frame (world) first, then redo the host-built one
    for p in range(len(pe)):
        assert np.array_equal(res_w.indices(p), want[p]), ("world", p)
    res = gpu_ctx.frame(ents, pairs, index_stride=stride)
    for p in range(len(pe)):
        assert np.array_equal(res.indices(p), want[p]), ("host-built", p)
    small = gpu_ctx.frame(ents, pairs, index_stride=6000, allow_overflow=True)       # truncation: prefix only, exact counts
    assert small.status == -5
    for p in range(len(pe)):
        assert int(small.index_counts[p]) == len(want[p])
        n = min(len(want[p]), 6000)
        assert np.array_equal(small.indices(p), want[p][:n]), ("truncated", p)


def test_want_flag_combinations_and_many_lights(gpu_ctx, oracle):
    """9 lights on one entity through host-built records (each warp takes several), BQ2_WANT_NOLERPOUT, and
    tangent space without shadow volumes."""
    m = Md2Model(oracle, 10, 7, frames=4, seed=5, tangent_space=True, smooth=True)
    mid = m.upload(gpu_ctx)
    we, wl, pe, pl = world_scene(2, 9, 4, seed=9)
    we["model"] = mid
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl, view_world=np.array([1, 2, 3], np.float32))
    res = gpu_ctx.frame(ents, pairs)
    want = []
    for p in range(18):
        W = we[pe[p]]
        want.append(oracle.md2_pair(m.blob, m.nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                                    W["oldorigin"], W["angles"], wl["origin"][pl[p]], 300.0, idx=m.idx))
        assert np.array_equal(res.indices(p), want[p]["indices"]), p
    # shadow-only caller: lerped[] is not written (poison it first through a frame with other content)
    gpu_ctx.frame(ents[::-1].copy(), None)
    before = res.lerped(0, m.V).copy()
    res2 = gpu_ctx.frame(ents, pairs, want=bq2.WANT_SHADOW | bq2.WANT_NOLERPOUT)
    assert np.array_equal(bits(res2.lerped(0, m.V)), bits(before))                 # untouched
    assert not np.array_equal(bits(before), bits(want[0]["lerped"]))
    for p in range(18):
        assert np.array_equal(res2.indices(p), want[p]["indices"]) and np.array_equal(bits(res2.extruded(p, m.V)), bits(want[p]["extruded"]))
    # tangent space only
    res3 = gpu_ctx.frame(ents, pairs, want=bq2.WANT_NTB | bq2.WANT_TSLIGHT)
    assert int(res3.total_indices) == 0 and not res3.index_counts.any()         # no volumes asked for: no stale counts
    from berserkerquake2_b200 import synth
    for e in range(2):
        f, of = int(we["frame"][e]), int(we["oldframe"][e]); bl = float(we["backlerp"][e]); fl = float(ents["frontlerp"][e])
        N, T_, B = oracle.ntb_rows(synth.md2_frame_verts(m.blob, f)[:, 3], synth.md2_frame_verts(m.blob, of)[:, 3],
                                   m.tangents[f], m.tangents[of], m.binormals[f], m.binormals[of], bl, fl)
        gN, gT, gB = res3.ntb(e, m.V)
        assert np.array_equal(bits(gN), bits(N)) and np.array_equal(bits(gT), bits(T_)) and np.array_equal(bits(gB), bits(B))
        P = want[9 * e]["lerped"]
        for l in range(9):
            p = 9 * e + l
            lp, ts = oracle.tslight(P, N, T_, B, pairs["light"][p], pairs["view"][p])
            gl, gt = res3.tslight(p, m.V)
            assert np.array_equal(bits(gl), bits(lp)) and np.array_equal(bits(gt), bits(ts))


def test_counts_are_cleared_without_shadow_and_unknown_want_bits_are_refused(oracle):
    """A frame without BQ2_WANT_SHADOW reports zero indices on a FRESH context (nothing stale to read back), on both
    routes; bits outside the public mask (the team kernel's internal one among them) are refused."""
    m = Md2Model(oracle, 10, 7, frames=4, seed=6, tangent_space=True, smooth=True)
    we, wl, pe, pl = world_scene(3, 2, 4, seed=10)
    with bq2.Context(0) as ctx:
        we["model"] = m.upload(ctx)
        ents, pairs = ctx.build_records(we, wl, pe, pl)
        r = ctx.frame(ents, pairs, want=bq2.WANT_NTB)
        assert int(r.total_indices) == 0 and not r.index_counts.any()
        r = ctx.world_frame(we, wl, pe, pl, want=bq2.WANT_TABLES)          # lerp only through the world route
        assert int(r.total_indices) == 0 and not r.index_counts.any()
        for bad in (0x40000000, 0x20, 0x80000000):
            with pytest.raises(bq2.Bq2Error):
                ctx.frame(ents, pairs, want=bq2.WANT_SHADOW | bad)
            with pytest.raises(bq2.Bq2Error):
                ctx.world_frame(we, wl, pe, pl, want=bq2.WANT_SHADOW | bad)
        r = ctx.frame(ents, pairs)                                           # the context is still usable
        assert int(r.total_indices) > 0


def test_anorm_bytes_beyond_the_table_decode_to_zero(oracle):
    """Bytes >= 162 (never produced by Normal2Index; the reference reads past r_avertexnormals) are DEFINED here: the
    device table has 256 rows, the extra ones zero -- a shell push-out along such a "normal" moves nothing, and a basis
    row decoded from it is the zero vector."""
    m = Md2Model(oracle, 10, 7, frames=3, seed=8, tangent_space=True, smooth=True)
    blob = m.blob.copy()
    V = m.V
    for f in range(3):
        synth_rows = synth_mod.md2_frame_verts(blob, f)
        synth_rows[::2, 3] = np.arange(len(synth_rows[::2]), dtype=np.uint8) % 94 + 162      # 162..255 on even vertices
    tang = m.tangents.copy(); tang[:, 1::3] = 255
    with bq2.Context(0) as ctx:
        mid = ctx.upload_md2(blob, m.nb, tang, m.binormals, None, smooth=True)
        mid0 = ctx.upload_md2(m.blob, m.nb, m.tangents, m.binormals, None, smooth=True)
        we, wl, pe, pl = world_scene(1, 1, 3, seed=3)
        we["rf_flags"] = 0x400                                              # RF_SHELL_RED: the push-out path
        ents = []
        for mm in (mid, mid0):
            we["model"] = mm
            e, p = ctx.build_records(we, wl, pe, pl)
            ents.append(e)
        bad = ctx.frame(ents[0], None, want=bq2.WANT_NTB)
        lerped_bad = bad.lerped(0, V); N, T_, _ = bad.ntb(0, V)
        ok = ctx.frame(ents[1], None, want=bq2.WANT_NTB)
        lerped_ok = ok.lerped(0, V); N0, T0, _ = ok.ntb(0, V)
        noshell = ents[1].copy(); noshell["flags"] = 0
        plain = ctx.frame(noshell, None).lerped(0, V)
        assert np.array_equal(bits(lerped_bad[::2]), bits(plain[::2]))       # zero normal: no push-out on those vertices
        assert np.array_equal(bits(lerped_bad[1::2]), bits(lerped_ok[1::2]))  # the others as before
        assert not N[::2].any() and np.array_equal(bits(N[1::2]), bits(N0[1::2]))
        assert not T_[1::3].any() and np.array_equal(bits(T_[0::3]), bits(T0[0::3]))


def test_ratio_sequence_is_exact(gpu_ctx):
    """The straight-line `num / sqrtf(len2)` sequence of the kernels (scale / VectorLength M:63607, 1 / VectorLength
    M:63620) against div.rn(sqrt.rn) on the device: EVERY bit pattern of len2 for the numerators the path uses most
    (1.0 for the triangle normals, 32 * radius for the extrusion), then 2^31 random pairs across the guarded exponent
    ranges and their edges, then raw random bit patterns (denormals, infinities, NaNs, negatives -> fallback)."""
    for num in (1.0, 9600.0, 32.0 * 150.0, 32.0 * 87.5, -9600.0, 3.0e-12, 7.0e11):
        assert gpu_ctx.selftest_ratio(0, num) == 0, num
    assert gpu_ctx.selftest_ratio(1, count=1 << 31, seed=1) == 0
    assert gpu_ctx.selftest_ratio(2, count=1 << 30, seed=2) == 0


@pytest.mark.parametrize("want", [bq2.WANT_SHADOW, bq2.WANT_SHADOW | bq2.WANT_NTB | bq2.WANT_TSLIGHT])
def test_back_to_back_replays_keep_stream_order(gpu_ctx, oracle, want):
    """Replays are launched with programmatic dependent launch: the next frame's kernel starts under the tail of the
    current one.  Kept frames of one shape SHARE their result arrays, so a kernel that wrote anything before the kernel
    in front of it had finished would leave a mixture behind.  Many replays back to back without a sync in between,
    then every result array must be exactly the last frame's."""
    from berserkerquake2_b200 import _capi as c
    m = Md2Model(oracle, 16, 17, frames=6, seed=77, tangent_space=True, smooth=True)
    mid = m.upload(gpu_ctx)
    n_ents, lights = 1024, 4
    kept = []
    view = np.array([40.0, -30.0, 25.0], np.float32)
    for i in range(3):
        we, wl, pe, pl = world_scene(n_ents, lights, 6, seed=3000 + i)
        we["model"] = mid
        ents, pairs = gpu_ctx.build_records(we, wl, pe, pl, view_world=view)
        r = gpu_ctx.frame(ents, pairs, want=want)
        arrays = {c.A_INDEX_COUNT: n_ents * lights, c.A_INDICES: n_ents * lights * int(r.index_stride),
                  c.A_LERPED: 3 * n_ents * m.V, c.A_FACING_BITS: n_ents * lights * ((m.T + 31) // 32)}
        if want & bq2.WANT_TSLIGHT:
            arrays[c.A_NORMALS] = 3 * n_ents * m.V
            arrays[c.A_LIGHTP] = 3 * n_ents * lights * m.V
        snap = {a: gpu_ctx.download(a, 0, n, np.uint32) for a, n in arrays.items()}
        ext = [bits(r.extruded(p, m.V)).copy() for p in (0, 1, 2047, 4095)]
        kept.append((gpu_ctx.keep(), snap, ext, r))
    assert not np.array_equal(kept[0][1][c.A_INDICES], kept[1][1][c.A_INDICES])
    for last in (0, 1, 2):
        for rep in range(40):
            for k in (0, 1, 2):
                gpu_ctx.replay_kept(kept[k][0])
        gpu_ctx.replay_kept(kept[last][0])
        gpu_ctx.wait()
        h, snap, ext, r = kept[last]
        cnt = snap[c.A_INDEX_COUNT]
        live = np.arange(int(r.index_stride))[None, :] < cnt[:, None]           # beyond a pair's count the row holds leftovers
        for a, want_arr in snap.items():
            got = gpu_ctx.download(a, 0, len(want_arr), np.uint32)
            if a == c.A_INDICES:
                got, want_arr = got.reshape(live.shape)[live], want_arr.reshape(live.shape)[live]
            assert np.array_equal(got, want_arr), (last, a)
        for p, e in zip((0, 1, 2047, 4095), ext):
            assert np.array_equal(bits(r.extruded(p, m.V)), e), (last, p)
    for h, *_ in kept:
        gpu_ctx.free_kept(h)


def test_world_frame_counters_survive_whatever_ran_on_the_state_before(gpu_ctx, oracle):
    """A world frame's last kernel hands the counts to the host and clears what the NEXT frame of the same shape
    accumulates into, so a steady-state frame has no clear in front.  Whatever else used the frame state in between --
    a frame of another shape, a host-built frame, a replay, a frame without shadow volumes, frames in flight -- the
    next world frame must still count from zero: counts, totals and the per-pair hash of every frame are compared with
    the first one's (graph route: no BQ2_WANT_TABLES)."""
    models = [Md2Model(oracle, s, r, frames=6, seed=300 + i) for i, (s, r) in enumerate([(8, 5), (16, 9)])]
    mids = [m.upload(gpu_ctx) for m in models]
    want = bq2.WANT_SHADOW

    def scene(n_ents, lights, seed):
        we, wl, pe, pl = world_scene(n_ents, lights, 6, n_models=2, seed=seed)
        we["model"] = np.array(mids)[we["model"]]
        return we, wl, pe, pl

    A, B = scene(90, 3, 41), scene(57, 5, 42)

    def run(s, w=want):
        r = gpu_ctx.world_frame(*s, want=w)
        return r.index_counts.copy(), int(r.total_indices), (gpu_ctx.hash_results(len(s[2])) if w & bq2.WANT_SHADOW else None)

    cnt_a, tot_a, h_a = run(A)
    cnt_b, tot_b, h_b = run(B)
    assert tot_a > 0 and tot_b > 0 and tot_a == int(cnt_a.sum()) and tot_b == int(cnt_b.sum())

    def same(got, ref):
        return np.array_equal(got[0], ref[0]) and got[1] == ref[1] and np.array_equal(got[2], ref[2])

    ref_a, ref_b = (cnt_a, tot_a, h_a), (cnt_b, tot_b, h_b)
    for _ in range(4):                                            # same shape again and again: the graph, no clear
        assert same(run(A), ref_a)
    assert same(run(B), ref_b) and same(run(A), ref_a) and same(run(A), ref_a)       # another shape in between
    ents, pairs = gpu_ctx.build_records(*A)                        # a host-built frame on the same state
    hb = gpu_ctx.frame(ents, pairs)
    assert int(hb.total_indices) == tot_a
    assert same(run(A), ref_a) and same(run(A), ref_a)
    gpu_ctx.replay(); r = gpu_ctx.wait()                           # a replay adds to the totals it cleared itself ...
    assert int(r.total_indices) == tot_a
    assert np.array_equal(gpu_ctx.hash_results(len(A[2])), h_a)    # ... and still finds the frame's pair lists
    assert same(run(A), ref_a) and same(run(A), ref_a)
    c0, t0, _ = run(A, 0)                                          # no shadow volumes: every count reads zero
    assert t0 == 0 and not c0.any()
    assert same(run(A), ref_a) and same(run(A), ref_a)
    d = gpu_ctx.world_desc(*A, None, want, 0); d._keep = gpu_ctx._keep
    for _ in range(3):
        gpu_ctx.world_submit(d)                                    # three in flight on three states, then the same again
    for _ in range(9):
        r = gpu_ctx.wait()
        assert int(r.total_indices) == tot_a and np.array_equal(r.index_counts, cnt_a)
        gpu_ctx.world_submit(d)
    for _ in range(3):
        r = gpu_ctx.wait()
        assert int(r.total_indices) == tot_a and np.array_equal(r.index_counts, cnt_a)
    assert same(run(B), ref_b) and same(run(A), ref_a)


def test_world_route_tangent_space_outputs_match_the_host_built_route(gpu_ctx, oracle):
    """N/T/B rows and LightP / tsH through the device-built route (bq2_world_frame with BQ2_WANT_NTB | BQ2_WANT_TSLIGHT:
    every model single-mesh with per-vertex rows) are bit-identical to the host-built route, which the oracle and the
    golden vectors check -- smooth MD2 of the warp-kernel and the team-kernel class and a one-surface MD3, rotated entities,
    a view origin; a frame that mixes in a NON-smooth MD2 takes the host-built route and returns its tables."""
    from scenes import Md3Model
    md2 = [Md2Model(oracle, 8, 5, frames=4, seed=21, tangent_space=True), Md2Model(oracle, 40, 30, frames=3, seed=22, tangent_space=True)]
    md3 = Md3Model(oracle, slices=12, rings=9, frames=4, num_meshes=1, seed=23)
    assert md2[1].T > 1024                                           # team-kernel class
    mids = [m.upload(gpu_ctx) for m in md2] + [md3.upload(gpu_ctx)]
    we, wl, pe, pl = world_scene(24, 3, 3, n_models=3, seed=77)
    we["model"] = np.array(mids)[we["model"]]
    view = np.array([31.0, -17.0, 55.0], np.float32)
    want = bq2.WANT_SHADOW | bq2.WANT_NTB | bq2.WANT_TSLIGHT
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl, view_world=view)
    rh = gpu_ctx.frame(ents, pairs, want=want)
    # rows are padded to four: the last three of a padded row may be padding (never written), they are left out
    Vs = [int(rh.ent_vert_offset[e + 1] - rh.ent_vert_offset[e]) for e in range(len(we))]
    host_ntb = [[x.copy() for x in rh.ntb(e, Vs[e])] for e in range(len(we))]
    host_ts = [[x.copy() for x in rh.tslight(p, Vs[pe[p]])] for p in range(len(pe))]
    host_ext = [rh.extruded(p, Vs[pe[p]]).copy() for p in range(len(pe))]
    host_idx = [rh.indices(p).copy() for p in range(len(pe))]
    for rnd in range(3):                                          # the third frame replays the captured graph
        rw = gpu_ctx.world_frame(we, wl, pe, pl, view_world=view, want=want | (bq2.WANT_TABLES if rnd == 0 else 0))
        if rnd:                                                   # without tables: rows of pair p start at p * stride
            stride_v = int(max(Vs))
            rw.pair_vert_offset = rw.pair_ts_offset = np.arange(len(pe) + 1, dtype=np.uint32) * stride_v
        assert np.array_equal(rw.index_counts, rh.index_counts)
        for e in range(len(we)):
            for x, y in zip(rw.ntb(e, Vs[e]), host_ntb[e]):
                assert np.array_equal(bits(x).reshape(-1)[:3 * Vs[e] - 9], bits(y).reshape(-1)[:3 * Vs[e] - 9]), (rnd, e)
        for p in range(len(pe)):
            v = Vs[pe[p]]
            for x, y in zip(rw.tslight(p, v), host_ts[p]):
                assert np.array_equal(bits(x).reshape(-1)[:3 * v - 9], bits(y).reshape(-1)[:3 * v - 9]), (rnd, p)
            assert np.array_equal(bits(rw.extruded(p, v)).reshape(-1)[:3 * v - 9], bits(host_ext[p]).reshape(-1)[:3 * v - 9])
            assert np.array_equal(rw.indices(p), host_idx[p])
    # a non-smooth MD2 in the frame: per-triangle rows, the host-built route takes over (its tables come back)
    flat = Md2Model(oracle, 6, 5, frames=3, seed=24, tangent_space=True, smooth=False)
    fid = flat.upload(gpu_ctx)
    we2 = we.copy(); we2["model"][0] = fid; we2["frame"][0] %= 3; we2["oldframe"][0] %= 3
    r2 = gpu_ctx.world_frame(we2, wl, pe, pl, view_world=view, want=want | bq2.WANT_TABLES)
    e2, p2 = gpu_ctx.build_records(we2, wl, pe, pl, view_world=view)
    r2h = gpu_ctx.frame(e2, p2, want=want)
    assert np.array_equal(r2.index_counts, r2h.index_counts) and np.array_equal(r2.ent_tb_offset, r2h.ent_tb_offset)
    assert int(r2.ent_tb_offset[1] - r2.ent_tb_offset[0]) == ((flat.T + 3) & ~3)


def test_world_route_keeps_a_model_whose_one_mesh_casts_no_shadow(gpu_ctx, oracle):
    """A one-surface MD3 whose skin says "no shadow" (M:63811-63816) is lerped but casts no volume.  It does not push the
    frame onto the host-built route (device-built frames return no pair_slot_first table): its pairs keep their slots
    with index_count 0, everything else equals the host-built route."""
    quiet = Md3Model(oracle, 10, 7, frames=4, num_meshes=1, seed=31, casts=[0])
    m2 = Md2Model(oracle, 10, 7, frames=4, seed=32)
    idq, id2 = quiet.upload(gpu_ctx), m2.upload(gpu_ctx)
    we, wl, pe, pl = world_scene(8, 3, 4, seed=63)
    we["model"] = [idq, id2, id2, idq, id2, idq, id2, id2]
    rw = gpu_ctx.world_frame(we, wl, pe, pl)
    assert not rw.out.pair_slot_first                                # the device-built route
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl)
    rh = gpu_ctx.frame(ents, pairs)
    assert np.array_equal(rw.index_counts, rh.index_counts) and int(rw.total_indices) == int(rh.total_indices) > 0
    Vq = quiet.meshes[0]["verts"].shape[1]
    for p in range(len(pe)):
        e = int(pe[p])
        if int(we["model"][e]) == idq:
            assert rw.index_counts[p] == 0
            assert np.array_equal(bits(rw.lerped(e, Vq)), bits(rh.lerped(e, Vq)))
        else:
            assert rw.index_counts[p] > 0 and np.array_equal(rw.indices(p), rh.indices(p))
            assert np.array_equal(bits(rw.extruded(p, m2.V)), bits(rh.extruded(p, m2.V)))
