"""Random garbage-in parity: arbitrary (in-range) triangle indices and neighbour tables, random vertex bytes,
degenerate triangles everywhere, lights inside the mesh.  The kernels must reproduce the oracle exactly on any
table the upload accepts.  NaNs compare equal to NaNs (a 0*inf NaN has a different payload on x86 and on the GPU)."""
import numpy as np
import pytest

import berserkerquake2_b200 as bq2
from berserkerquake2_b200 import synth
import oracle_lib

pytestmark = pytest.mark.gpu


def same_f(a, b):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint32), b[~nb].view(np.uint32))


def random_md2(rng, V, T, F):
    """an MD2 image with the given sizes and random contents (the synthetic generator only fixes the layout)"""
    S, R = 3, 2                                   # smallest sphere, then resize by hand
    hdr = np.zeros(17, "<i4")
    framesize = 40 + 4 * V
    ofs_st = 68; ofs_tris = ofs_st + 4 * V; ofs_frames = ofs_tris + 12 * T; end = ofs_frames + F * framesize
    hdr[:] = [844121161, 8, 256, 256, framesize, 0, V, V, T, 0, F, 68, ofs_st, ofs_tris, ofs_frames, end, end]
    blob = np.zeros(end, np.uint8)
    blob[:68] = hdr.view(np.uint8)
    blob[ofs_st:ofs_tris] = rng.randint(0, 256, 4 * V).astype(np.uint8)
    tris = rng.randint(0, V, (T, 6)).astype("<i2")
    blob[ofs_tris:ofs_frames] = tris.view(np.uint8).reshape(-1)
    for f in range(F):
        o = ofs_frames + f * framesize
        blob[o:o + 24] = np.concatenate([rng.uniform(0.1, 0.5, 3), rng.uniform(-40, 0, 3)]).astype("<f4").view(np.uint8)
        vb = rng.randint(0, 256, (V, 4)).astype(np.uint8); vb[:, 3] = rng.randint(0, 162, V)
        blob[o + 40:o + 40 + 4 * V] = vb.reshape(-1)
    return blob


@pytest.mark.parametrize("seed", range(6))
def test_random_tables(gpu_ctx, oracle, seed):
    rng = np.random.RandomState(1000 + seed)
    models, mids = [], []
    for k in range(5):
        V = int(rng.randint(3, 70)); T = int(rng.randint(1, 140)) if k else 1
        if k == 4:
            V, T = 1500, 2100                     # team-kernel class
        blob = random_md2(rng, V, T, 3)
        nb = rng.randint(-1, T, (T, 3)).astype(np.int32)
        models.append((blob, nb, oracle_lib.md2_idx(blob), V, T)); mids.append(gpu_ctx.upload_md2(blob, nb))
    n = 12
    we = np.zeros(n, bq2.WORLD_ENTITY_DTYPE)
    we["model"] = np.array(mids)[rng.randint(0, 5, n)]
    we["frame"] = rng.randint(0, 3, n); we["oldframe"] = rng.randint(0, 3, n)
    we["backlerp"] = rng.randint(0, 1025, n) / np.float32(1024)
    we["origin"] = rng.randn(n, 3) * 30; we["oldorigin"] = we["origin"] + rng.randn(n, 3)
    we["angles"][::3] = rng.uniform(-180, 180, (len(we[::3]), 3))
    wl = np.zeros(20, bq2.WORLD_LIGHT_DTYPE)
    wl["origin"] = rng.randn(20, 3) * 60; wl["radius"] = rng.uniform(50, 400, 20)
    wl["origin"][:4] = we["origin"][:4] + rng.randn(4, 3) * 5            # lights inside the meshes
    pe = rng.randint(0, n, 40).astype(np.int32); pl = rng.randint(0, 20, 40).astype(np.int32)
    res = gpu_ctx.world_frame(we, wl, pe, pl, index_stride=24 * 2100)
    for p in range(40):
        e = int(pe[p]); W = we[e]
        blob, nb, idx, V, T = models[mids.index(int(W["model"]))]
        w = oracle.md2_pair(blob, nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"], W["oldorigin"],
                            W["angles"], wl["origin"][pl[p]], float(wl["radius"][pl[p]]), idx=idx)
        used = np.unique(idx)
        assert same_f(res.lerped(e, V), w["lerped"]), (seed, p)
        assert same_f(res.extruded(p, V)[used], w["extruded"][used]), (seed, p)
        assert np.array_equal(res.facing(p, T), w["viss"]), (seed, p)
        assert np.array_equal(res.indices(p), w["indices"]), (seed, p)


@pytest.mark.parametrize("seed", range(4))
def test_random_entity_records(gpu_ctx, oracle, seed):
    """The entity records the world path builds on the device (lerp constants M:63526-63553 / M:63825-63856, shell
    offset, frame rows): random angles (single-axis, signed zeros, full), backlerp incl. 0 and 1, origins from tiny to
    1e6, random RF_ flag words, MD2 and one-surface MD3 -- lerped vertices against the oracle for every entity, and
    the volume of every pair that casts one."""
    from scenes import Md2Model, Md3Model, bits
    import berserkerquake2_b200.host as host
    rng = np.random.RandomState(7000 + seed)
    m2 = [Md2Model(oracle, 8 + 2 * k, 5 + k, frames=5, seed=300 + 10 * seed + k) for k in range(2)]
    m3 = Md3Model(oracle, 10, 7, frames=5, num_meshes=1, seed=40 + seed)
    ids = [m.upload(gpu_ctx) for m in m2] + [m3.upload(gpu_ctx)]
    n = 60
    we = np.zeros(n, bq2.WORLD_ENTITY_DTYPE)
    which = rng.randint(0, 3, n)
    we["model"] = np.array(ids)[which]
    we["frame"] = rng.randint(0, 5, n); we["oldframe"] = rng.randint(0, 5, n)
    we["backlerp"] = rng.choice(np.array([0.0, 1.0, 0.5, 0.3330078125, 1e-6, 0.999], np.float32), n)
    scale = rng.choice(np.array([1e-3, 1.0, 100.0, 1e4, 1e6], np.float32), n)[:, None]
    we["origin"] = (rng.randn(n, 3) * scale).astype(np.float32)
    we["oldorigin"] = we["origin"] + (rng.randn(n, 3) * np.minimum(scale, 10)).astype(np.float32)
    kind = rng.randint(0, 5, n)
    ang = rng.uniform(-720, 720, (n, 3)).astype(np.float32)
    ang[kind == 0] = 0.0
    ang[kind == 1, 0] = 0.0; ang[kind == 1, 2] = 0.0                   # yaw only
    ang[kind == 2, 1] = -0.0; ang[kind == 2, 2] = 0.0                  # pitch only, a negative zero beside it
    ang[kind == 3, 0] = -0.0; ang[kind == 3, 1] = -0.0                 # roll only
    we["angles"] = ang
    flag_bits = np.array([0x4, 0x20, 0x400, 0x800, 0x1000, 0x8, 0x2, 0x80000, 0x10000000, 0x40], np.uint32)
    fl = np.zeros(n, np.uint32)
    for e in range(n):
        if rng.rand() < 0.5:
            fl[e] = np.bitwise_or.reduce(rng.choice(flag_bits, rng.randint(1, 3)))
    we["rf_flags"] = fl
    L = 2
    wl = np.zeros(n * L, bq2.WORLD_LIGHT_DTYPE)
    wl["origin"] = np.repeat(we["origin"], L, axis=0) + (rng.randn(n * L, 3) * 120).astype(np.float32)
    wl["radius"] = rng.uniform(50, 400, n * L)
    pe = np.repeat(np.arange(n, dtype=np.int32), L); pl = np.arange(n * L, dtype=np.int32)
    res = gpu_ctx.world_frame(we, wl, pe, pl)
    mesh = m3.meshes[0]
    for e in range(n):
        W = we[e]; f = int(W["rf_flags"])
        if which[e] == 2:
            V = mesh["verts"].shape[1]
            w = oracle.md3_mesh_pair(mesh["verts"], mesh["indexes"], mesh["neighbours"], m3.frame_translate, int(W["frame"]),
                                     int(W["oldframe"]), float(W["backlerp"]), W["origin"], W["oldorigin"], W["angles"],
                                     wl["origin"][e * L], float(wl["radius"][e * L]))["lerped"]
        else:
            m = m2[which[e]]; V = m.V
            shell = bool(f & 0x1c00)
            w = oracle.md2_lerp(m.blob, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"], W["oldorigin"],
                                W["angles"], shell=shell, shell_scale=(0.5 if f & 0x4 else 1.0) if shell else 0.0)
        assert same_f(res.lerped(e, V), w), (seed, e, hex(f))
    for p in range(n * L):
        e = int(pe[p]); W = we[e]; f = int(W["rf_flags"])
        casts = host.casts_shadow(f) and not (f & (0x80000 | 0x10000000))
        if not casts:
            assert int(res.index_counts[p]) == 0, (seed, p, hex(f))
            continue
        if which[e] == 2:
            w = oracle.md3_mesh_pair(mesh["verts"], mesh["indexes"], mesh["neighbours"], m3.frame_translate, int(W["frame"]),
                                     int(W["oldframe"]), float(W["backlerp"]), W["origin"], W["oldorigin"], W["angles"],
                                     wl["origin"][pl[p]], float(wl["radius"][pl[p]]))
            V, T = mesh["verts"].shape[1], len(mesh["indexes"])
        else:
            m = m2[which[e]]
            w = oracle.md2_pair(m.blob, m.nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                                W["oldorigin"], W["angles"], wl["origin"][pl[p]], float(wl["radius"][pl[p]]), idx=m.idx)
            V, T = m.V, m.T
        assert same_f(res.extruded(p, V), w["extruded"]), (seed, p)
        assert np.array_equal(res.facing(p, T), w["viss"]) and np.array_equal(res.indices(p), w["indices"]), (seed, p)
