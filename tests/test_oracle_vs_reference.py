"""Pins the CPU oracle (oracle/bq2_oracle.c) to the reference's OWN object code: the unmodified main.c
compiled headless into oracle/_ref/libbq2ref.so (oracle/ref/build_ref.sh).  Everything is compared bit
for bit.  The reference ships no tests or fixtures; executing it is the pin.
Skipped when the reference build is absent (it needs /root/reference)."""
import numpy as np
import pytest

import oracle_lib
from berserkerquake2_b200 import synth, host
from scenes import Md2Model, Md3Model, world_scene, bits, SMOOTH_COSINE


def same(a, b):
    return np.array_equal(bits(np.asarray(a, np.float32)), bits(np.asarray(b, np.float32)))


def test_constants(oracle, ref):
    assert same(oracle.anorms(), ref.anorms())                    # anorms.h, 162 rows
    assert np.float32(ref.smooth_cosine()) == np.float32(SMOOTH_COSINE)
    a = np.ctypeslib.as_array(
        __import__("ctypes").cast(__import__("berserkerquake2_b200")._capi.lib.bq2_host_anorms(),
                                  __import__("ctypes").POINTER(__import__("ctypes").c_float)), (162, 4))
    assert same(a[:, :3], ref.anorms())                            # the table the CUDA library carries


def test_angle_vectors(oracle, ref):
    rng = np.random.RandomState(1)
    for ang in [(0, 0, 0), (0, 90, 0), (30, 0, 0), (0, 0, 45), (-12.5, 333.25, 7), (-0.0, 45, 0), (0, -0.0, 10),
                (-0.0, -0.0, 33), (1e-40, 20, 0), (0, 1e-44, -0.0), (360, 0, 0), (0, 180, 0)] + [tuple(rng.uniform(-360, 360, 3)) for _ in range(50)]:
        for x, y, z in zip(oracle.angle_vectors(ang), ref.angle_vectors(ang), host.angle_vectors(ang)):
            assert same(x, y) and same(y, z), ang


def test_normal2index(oracle, ref):
    rng = np.random.RandomState(2)
    vs = list(rng.randn(300, 3).astype(np.float32)) + [np.zeros(3, np.float32), -np.ones(3, np.float32)]
    vs += list(ref.anorms()[:20]) + [np.array([np.nan, 0, 0], np.float32)]
    for v in vs:
        assert oracle.normal2index(v) == ref.normal2index(v) == synth.normal2index(v)


def test_vecs_for_tris(oracle, ref):
    rng = np.random.RandomState(3)
    for _ in range(200):
        v = rng.randint(0, 256, (3, 3)).astype(np.float32)
        st = rng.rand(3, 2).astype(np.float32)
        a, b = oracle.vecs_for_tris(*v, *st), ref.vecs_for_tris(*v, *st)
        assert same(a[0], b[0]) and same(a[1], b[1])


@pytest.mark.parametrize("shape", [(16, 17), (8, 5), (5, 3), (24, 13)])
def test_md2_neighbours(oracle, ref, shape):
    rng = np.random.RandomState(shape[0])
    for trial in range(6):
        blob = synth.md2(shape[0], shape[1], 1, trial)
        tris = synth.md2_tris(blob)
        T, V = len(tris), int(tris[:, :3].max()) + 1
        for _ in range(trial * 2):                                 # open edges, 3-triangle edges, degenerates
            k, a = rng.randint(4), rng.randint(T)
            if k == 0: tris[a, :3] = tris[rng.randint(T), :3]
            elif k == 1: tris[a, rng.randint(3)] = rng.randint(V)
            elif k == 2: tris[a, :3] = tris[a, [0, 1, 0]]
            else: tris[a, :3] = tris[rng.randint(T), [2, 1, 0]]
        r = ref.md2_neighbours(blob)
        assert np.array_equal(oracle.md2_neighbours(blob), r)
        assert np.array_equal(host.md2_neighbours(blob), r)        # the product's sorted-edge builder
        ix = tris[:, :3].astype(np.uint16)
        r3 = ref.md3_neighbours(ix)
        assert np.array_equal(oracle.md3_neighbours(ix), r3)
        assert np.array_equal(host.md3_neighbours(ix), r3)


def test_md2_lerp_and_shell(oracle, ref):
    """R_CalcAliasFrameLerp M:63510 alone, with and without the shell push-out."""
    m = Md2Model(oracle, frames=20, seed=7)
    rng = np.random.RandomState(4)
    for _ in range(20):
        f, of = rng.randint(20), rng.randint(20)
        bl = float(np.float32(rng.rand()))
        org, oorg = rng.randn(3) * 100, rng.randn(3) * 100
        ang = (0, 0, 0) if rng.rand() < 0.5 else rng.uniform(-180, 180, 3)
        assert same(oracle.md2_lerp(m.blob, f, of, bl, org, oorg, ang), ref.md2_lerp(m.blob, f, of, bl, org, oorg, ang))
        for flags, scale in ((0x400, 1.0), (0x800 | 0x4, 0.5)):    # RF_SHELL_RED; RF_SHELL_GREEN|RF_WEAPONMODEL
            assert same(oracle.md2_lerp(m.blob, f, of, bl, org, oorg, ang, shell=True, shell_scale=scale),
                        ref.md2_lerp(m.blob, f, of, bl, org, oorg, ang, flags=flags))


def check_pair(o, r, V, idx):
    used = np.unique(idx)                      # the reference writes extrudedVerts only for referenced verts
    assert same(o["lerped"], r["lerped"])
    assert same(o["extruded"][used], r["extruded"][used])
    assert np.array_equal(o["viss"], r["viss"])
    assert same(o["tris_verts"], r["tris_verts"])
    if "trinormals" in r and "trinormals" in o:
        assert np.array_equal(bits(o["trinormals"]), bits(r["trinormals"]))
    # canonical index form dereferences to the reference's vertex stream
    pool = np.concatenate([o["lerped"], o["extruded"]])
    assert same(pool[o["indices"]], r["tris_verts"])
    assert o["indices"].max(initial=0) < 2 * V


def test_md2_shadow_volume_whole(oracle, ref):
    """R_DrawAliasShadowVolume M:63713 end to end: light -> model space, lerp, facing, sides + caps."""
    models = [Md2Model(oracle, frames=12, seed=s) for s in range(3)] + [Md2Model(oracle, 8, 5, 4, 9), Md2Model(oracle, 62, 34, 2, 3)]
    n = 0
    for m in models:
        we, wl, pe, pl = world_scene(8, 2, m.F, seed=100 + m.T)
        for p in range(len(pe)):
            W = we[pe[p]]
            a = (m.blob, m.nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"], W["oldorigin"],
                 W["angles"], wl["origin"][pl[p]], 300.0)
            o, r = oracle.md2_pair(*a, idx=m.idx), ref.md2_pair(*a)
            assert same(o["light_local"], r["light_local"])
            assert same(o["light_local"], host.point_to_model(wl["origin"][pl[p]], W["origin"], W["angles"]))
            check_pair(o, r, m.V, m.idx)
            n += 1
    assert n == 80


def test_md2_open_and_degenerate(oracle, ref):
    m = Md2Model(oracle, 8, 5, frames=3, seed=5)
    tris = synth.md2_tris(m.blob)
    tris[3, 0:3] = tris[3, [0, 0, 2]]; tris[10, 0:3] = tris[11, 0:3]; tris[20, 0:3] = (0, 1, 1)
    m.idx = oracle_lib.md2_idx(m.blob); m.nb = ref.md2_neighbours(m.blob)
    assert (m.nb == -1).any()
    for lx in (50.0, -200.0):
        o = oracle.md2_pair_local(m.blob, m.nb, 1, 0, 0.25, (lx, 30, 80), 300.0, idx=m.idx)
        r = ref.md2_pair_local(m.blob, m.nb, 1, 0, 0.25, (lx, 30, 80), 300.0)
        check_pair(o, r, m.V, m.idx)
        assert r["viss"][3] == 1 and np.isnan(r["trinormals"][3]).all()       # NaN normal -> "facing"


def test_md2_flag_filter(oracle, ref):
    """M:63722-63727: which entities cast a volume at all."""
    m = Md2Model(oracle, 8, 5, frames=2, seed=1)
    z = np.zeros(3, np.float32)
    for flags in (0, 0x2, 0x4, 0x8, 0x20, 0x80, 0x400, 0x800, 0x1000, 0x10000000, 0x80000):
        r = ref.md2_pair(m.blob, m.nb, 1, 0, 0.5, z, z, z, (100, 50, 60), 300.0, flags=flags)
        assert (r is not None) == host.casts_shadow(flags), hex(flags)


def test_md2_tangent_precompute(oracle, ref):
    """Loader loops M:51661-51804 restated in the oracle, every inner call checked against the reference's
    VecsForTris / Normal2Index above; here the assembled byte arrays feed the reference's own light-lerp."""
    for smooth in (True, False):
        m = Md2Model(oracle, 10, 7, frames=4, seed=6, tangent_space=True, smooth=smooth)
        light, view = np.array([120, -40, 90], np.float32), np.array([-30, 60, 20], np.float32)
        for f, of, bl in ((1, 0, 0.25), (3, 2, 0.8125), (2, 2, 0.0)):
            fl = float(np.float32(1) - np.float32(bl))
            P = ref.md2_lerp(m.blob, f, of, bl)
            if smooth:
                nc, no = synth.md2_frame_verts(m.blob, f)[:, 3], synth.md2_frame_verts(m.blob, of)[:, 3]
            else:
                nc, no = m.normals[f], m.normals[of]
            N, T_, B = oracle.ntb_rows(nc, no, m.tangents[f], m.tangents[of], m.binormals[f], m.binormals[of], bl, fl)
            corner = m.idx.reshape(-1)
            row = corner if smooth else np.repeat(np.arange(m.T), 3)
            # GL_DrawAliasFrameLerpLightARB_vp M:65140: TMU1 binormal, TMU2 normal, TMU3 tangent, per corner
            n, tmu, vert = ref.md2_light(m.blob, m.tangents, m.binormals, m.normals, smooth, f, of, bl, light, view, 1)
            assert n == 3 * m.T and same(vert, P[corner])
            assert same(tmu[1], B[row]) and same(tmu[2], N[row]) and same(tmu[3], T_[row])
            # GL_DrawAliasFrameLerpLightARB M:64889: TMU1 LightP, TMU3 tsH, TMU2 expanded vertices
            n, tmu, vert = ref.md2_light(m.blob, m.tangents, m.binormals, m.normals, smooth, f, of, bl, light, view, 0)
            lp, ts = oracle.tslight(P[corner], N[row], T_[row], B[row], light, view)
            assert same(tmu[1], lp) and same(tmu[3], ts) and same(tmu[2], P[corner])


def test_md3_shadow_volume_whole(oracle, ref):
    """R_DrawMD3AliasShadowVolume M:63800 whole, per-mesh results captured inside its glDrawArrays."""
    md3 = Md3Model(oracle, 12, 9, frames=5, num_meshes=3, seed=2, casts=[1, 0, 1])
    for mesh in md3.meshes:
        assert np.array_equal(mesh["neighbours"], ref.md3_neighbours(mesh["indexes"]))
    we, wl, pe, pl = world_scene(6, 2, 5, seed=77)
    for p in range(len(pe)):
        W = we[pe[p]]
        drawn, res = ref.md3_shadow(md3.frame_translate, md3.meshes, int(W["frame"]), int(W["oldframe"]),
                                    float(W["backlerp"]), W["origin"], W["oldorigin"], W["angles"], wl["origin"][pl[p]], 300.0)
        assert drawn == 2 and res[1] is None
        for k in (0, 2):
            mesh = md3.meshes[k]
            o = oracle.md3_mesh_pair(mesh["verts"], mesh["indexes"], mesh["neighbours"], md3.frame_translate,
                                     int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                                     W["oldorigin"], W["angles"], wl["origin"][pl[p]], 300.0)
            check_pair(o, res[k], mesh["verts"].shape[1], np.asarray(mesh["indexes"], np.int32))


def test_md3_light_family(oracle, ref):
    """GL_DrawMD3AliasFrameLerpLightARB(_vp) M:65671-65898 on the oracle's tangent/binormal bytes."""
    md3 = Md3Model(oracle, 10, 7, frames=4, num_meshes=1, seed=4)
    v = md3.meshes[0]["verts"]; V = v.shape[1]
    light, view = np.array([50, 80, -20], np.float32), np.array([5, -60, 40], np.float32)
    for f, of, bl in ((1, 0, 0.25), (3, 1, 0.59375)):
        fl = float(np.float32(1) - np.float32(bl))
        move, _ = oracle.md3_lerp_consts(md3.frame_translate[f], md3.frame_translate[of], bl, (0, 0, 0), (0, 0, 0), (0, 0, 0))
        P = np.zeros((V, 3), np.float32)
        oracle.lib.bq2o_md3_lerp_extrude(oracle_lib.P(np.ascontiguousarray(v[f])), oracle_lib.P(np.ascontiguousarray(v[of])), V,
                                         oracle_lib.P(move), bl, fl, oracle_lib.P(light), 9600.0, oracle_lib.P(P), None)
        N, T_, B = oracle.ntb_rows(v["normal"][f], v["normal"][of], v["tangent"][f], v["tangent"][of],
                                   v["binormal"][f], v["binormal"][of], bl, fl)
        tmu = ref.md3_light(v, f, of, bl, P, light, view, 1)
        assert same(tmu[1], B) and same(tmu[2], N) and same(tmu[3], T_)
        tmu = ref.md3_light(v, f, of, bl, P, light, view, 0)
        lp, ts = oracle.tslight(P, N, T_, B, light, view)
        assert same(tmu[1], lp) and same(tmu[3], ts)


def test_md3_tangent_precompute(oracle, ref):
    """TangentForTrimd3 M:68755 per triangle, then the accumulate/normalise/quantise loop M:50993-51014."""
    verts, idx, st = synth.md3_mesh(8, 5, 2, 3)
    import ctypes as C
    for t in range(len(idx)):
        tan_r, bin_r = np.zeros(3, np.float32), np.zeros(3, np.float32)
        ix = np.ascontiguousarray(idx[t]); vv = np.ascontiguousarray(verts[0]); s2 = np.ascontiguousarray(st)
        ref.lib.bq2ref_tangent_for_tri_md3(oracle_lib.P(ix), oracle_lib.P(vv), oracle_lib.P(s2), oracle_lib.P(tan_r), oracle_lib.P(bin_r))
        # the same through the oracle's public entry: one-triangle mesh accumulates exactly this tangent
        one = verts[:1].copy()
        oracle.md3_tangents(one, st, idx[t:t + 1])
        tn = tan_r.copy(); bn = bin_r.copy()
        for k in idx[t]:
            assert one[0]["tangent"][k] == ref.normal2index(tn) and one[0]["binormal"][k] == ref.normal2index(bn)


def test_brush_model_volumes(oracle, ref):
    """R_DrawBrushModelVolumes M:63326-63499 whole: vcache + icache, rotated entities,
    fence surfaces, open edges, everything lit / nothing lit."""
    from scenes import BrushModel
    rng = np.random.RandomState(12)
    n = 0
    for seed, (S, B, oe) in enumerate([(8, 3, 0), (5, 2, 3), (12, 4, 0), (3, 1, 1)]):
        bm = BrushModel(S, B, seed, oe)
        for trial in range(6):
            origin = rng.randn(3) * 50
            angles = np.zeros(3) if trial % 2 == 0 else rng.uniform(-180, 180, 3)
            light_world = origin + rng.randn(3) * 120
            lm = oracle.point_to_model(light_world, origin, angles)
            lit = bm.lit(lm)
            if trial == 4: lit[:] = 1
            if trial == 5: lit[:] = 0
            vr, ir = ref.brush_volume(bm, lit, origin, angles, light_world, 300.0)
            vo, io = oracle.brush_volume(bm, lit, lm, 300.0)
            assert np.array_equal(bits(vo), bits(vr)) and np.array_equal(io, ir), (seed, trial)
            assert (len(io) > 0) == bool((lit & (1 - bm.fence)).any())
            n += len(io)
    assert n > 1000
