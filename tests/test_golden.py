"""Golden vectors produced by the reference's own code (tools/make_golden.py ran oracle/_ref/libbq2ref.so
in the build container) replayed against the CPU oracle (-m "not gpu") and the CUDA library (-m gpu).
The reference tree is NOT needed at run time: everything comes from tests/golden/reference_vectors.npz."""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib
import berserkerquake2_b200 as bq2
from berserkerquake2_b200 import synth, host
from scenes import bits

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


def same(a, b):
    return np.array_equal(bits(np.asarray(a, np.float32)), bits(np.asarray(b, np.float32)))


def md2_case(name):
    S, R, F, seed, n_ent, n_light = [int(x) for x in G[f"{name}/params"]]
    blob = synth.md2(S, R, F, seed)
    # the generator must still produce the bytes the reference was run on
    assert np.array_equal(np.frombuffer(hashlib.sha256(blob.tobytes()).digest(), np.uint8), G[f"{name}/blob_sha256"])
    if f"{name}/blob" in G:
        assert np.array_equal(blob, G[f"{name}/blob"])
    we = np.ascontiguousarray(G[f"{name}/world_ents"]).view(bq2.WORLD_ENTITY_DTYPE).reshape(-1)
    wl = np.ascontiguousarray(G[f"{name}/world_lights"]).view(bq2.WORLD_LIGHT_DTYPE).reshape(-1)
    pe = np.repeat(np.arange(n_ent, dtype=np.int32), n_light)
    pl = np.arange(n_ent * n_light, dtype=np.int32)
    return blob, G[f"{name}/neighbours"], we, wl, pe, pl


def md3_case():
    meshes = []
    for k in range(3):
        v = np.ascontiguousarray(G[f"md3/mesh{k}/verts"]).view(synth.MD3_VERTEX).reshape(G[f"md3/mesh{k}/verts"].shape[:2])
        meshes.append(dict(verts=v, indexes=G[f"md3/mesh{k}/indexes"], neighbours=G[f"md3/mesh{k}/neighbours"],
                           casts_shadow=bool(G["md3/casts"][k])))
    we = np.ascontiguousarray(G["md3/world_ents"]).view(bq2.WORLD_ENTITY_DTYPE).reshape(-1)
    wl = np.ascontiguousarray(G["md3/world_lights"]).view(bq2.WORLD_LIGHT_DTYPE).reshape(-1)
    return G["md3/frame_translate"], meshes, we, wl


# ---------------------------------------------------------------- CPU: oracle vs the reference's vectors
@pytest.mark.parametrize("name", ["md2_small", "md2_c1"])
def test_oracle_md2_pairs(oracle, name):
    blob, nb, we, wl, pe, pl = md2_case(name)
    assert np.array_equal(oracle.md2_neighbours(blob), nb)
    assert np.array_equal(host.md2_neighbours(blob), nb)
    V = oracle_lib.md2_dims(blob)[0]
    for p in range(len(pe)):
        W = we[pe[p]]
        o = oracle.md2_pair(blob, nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                            W["oldorigin"], W["angles"], wl["origin"][pl[p]], float(wl["radius"][pl[p]]))
        for k in ("lerped", "extruded", "tris_verts", "light_local", "trinormals"):
            assert same(o[k], G[f"{name}/pair{p}/{k}"]), (p, k)
        assert np.array_equal(o["viss"], G[f"{name}/pair{p}/viss"])
        pool = np.concatenate([o["lerped"], o["extruded"]])
        assert same(pool[o["indices"]], G[f"{name}/pair{p}/tris_verts"])        # canonical index form
        assert same(host.point_to_model(wl["origin"][pl[p]], W["origin"], W["angles"]), G[f"{name}/pair{p}/light_local"])


def test_oracle_md2_broken_mesh(oracle):
    blob, nb = G["md2_broken/blob"], G["md2_broken/neighbours"]
    assert np.array_equal(oracle.md2_neighbours(blob), nb) and np.array_equal(host.md2_neighbours(blob), nb)
    ix = synth.md2_tris(blob)[:, :3].astype(np.uint16)
    assert np.array_equal(oracle.md3_neighbours(ix), G["md2_broken/md3style_neighbours"])
    assert np.array_equal(host.md3_neighbours(ix), G["md2_broken/md3style_neighbours"])
    used = np.unique(ix)
    for i in range(2):
        o = oracle.md2_pair_local(blob, nb, 1, 0, 0.25, G[f"md2_broken/pair{i}/light"], 300.0)
        assert same(o["lerped"], G[f"md2_broken/pair{i}/lerped"])
        assert same(o["extruded"][used], G[f"md2_broken/pair{i}/extruded"][used])
        assert np.array_equal(o["viss"], G[f"md2_broken/pair{i}/viss"]) and o["viss"][3] == 1
        assert same(o["tris_verts"], G[f"md2_broken/pair{i}/tris_verts"])


def test_oracle_shell(oracle):
    blob = synth.md2(8, 5, 4, 21)
    for i, scale in enumerate((1.0, 0.5)):
        o = oracle.md2_lerp(blob, 2, 1, 0.375, (1, 2, 3), (2, 2, 2.5), (0, 25, 0), shell=True, shell_scale=scale)
        assert same(o, G[f"md2_shell/{i}/lerped"])
        e = host.entity_md2(blob, 2, 1, 0.375, (1, 2, 3), (2, 2, 2.5), (0, 25, 0), rf_flags=int(G[f"md2_shell/{i}/flags"][0]))
        assert e["flags"] == 1 and e["shell_scale"] == scale


@pytest.mark.parametrize("kind", ["smooth", "flat"])
def test_oracle_md2_light_family(oracle, kind):
    tag = f"light/md2_{kind}"
    blob, tan, bi = G[f"{tag}/blob"], G[f"{tag}/tangents"], G[f"{tag}/binormals"]
    idx = oracle_lib.md2_idx(blob); T = len(idx)
    for i in range(2):
        f, of = [int(x) for x in G[f"{tag}/{i}/frames"]]; bl = float(G[f"{tag}/{i}/backlerp"][0])
        fl = float(np.float32(1) - np.float32(bl))
        if kind == "smooth":
            nc, no = synth.md2_frame_verts(blob, f)[:, 3], synth.md2_frame_verts(blob, of)[:, 3]
            row = idx.reshape(-1)
        else:
            nc, no = G[f"{tag}/normals"][f], G[f"{tag}/normals"][of]
            row = np.repeat(np.arange(T), 3)
        N, T_, B = oracle.ntb_rows(nc, no, tan[f], tan[of], bi[f], bi[of], bl, fl)
        assert same(N[row], G[f"{tag}/{i}/normal"]) and same(T_[row], G[f"{tag}/{i}/tangent"]) and same(B[row], G[f"{tag}/{i}/binormal"])
        P = oracle.md2_lerp(blob, f, of, bl)[idx.reshape(-1)]
        assert same(P, G[f"{tag}/{i}/verts"])
        lp, ts = oracle.tslight(P, N[row], T_[row], B[row], G["light/light"], G["light/view"])
        assert same(lp, G[f"{tag}/{i}/lightp"]) and same(ts, G[f"{tag}/{i}/tsh"])


def test_oracle_md3(oracle):
    ft, meshes, we, wl = md3_case()
    for k, m in enumerate(meshes):
        assert np.array_equal(oracle.md3_neighbours(m["indexes"]), m["neighbours"])
        assert np.array_equal(host.md3_neighbours(m["indexes"]), m["neighbours"])
    for p in range(6):
        W = we[p // 2]
        for k in (0, 2):
            m = meshes[k]
            o = oracle.md3_mesh_pair(m["verts"], m["indexes"], m["neighbours"], ft, int(W["frame"]), int(W["oldframe"]),
                                     float(W["backlerp"]), W["origin"], W["oldorigin"], W["angles"], wl["origin"][p], 300.0)
            for key in ("lerped", "extruded", "tris_verts"):
                assert same(o[key], G[f"md3/pair{p}/mesh{k}/{key}"]), (p, k, key)
            assert np.array_equal(o["viss"], G[f"md3/pair{p}/mesh{k}/viss"])
    v = meshes[0]["verts"]
    for i in range(2):
        f, of = [int(x) for x in G[f"md3/light{i}/frames"]]; bl = float(G[f"md3/light{i}/backlerp"][0])
        fl = float(np.float32(1) - np.float32(bl))
        N, T_, B = oracle.ntb_rows(v["normal"][f], v["normal"][of], v["tangent"][f], v["tangent"][of],
                                   v["binormal"][f], v["binormal"][of], bl, fl)
        assert same(N, G[f"md3/light{i}/normal"]) and same(T_, G[f"md3/light{i}/tangent"]) and same(B, G[f"md3/light{i}/binormal"])
        lp, ts = oracle.tslight(G["md3/pair0/mesh0/lerped"], N, T_, B, G["light/light"], G["light/view"])
        assert same(lp, G[f"md3/light{i}/lightp"]) and same(ts, G[f"md3/light{i}/tsh"])


def test_host_trig_golden(oracle):
    for a, fru in zip(G["trig/angles"], G["trig/fru"]):
        assert same(np.concatenate(host.angle_vectors(a)), fru)
        assert same(np.concatenate(oracle.angle_vectors(a)), fru)


# ---- the engine's loaders (vectors from Mod_LoadAliasModel / Mod_LoadAliasMD3Model run whole) -----------------------
def loader_md2(kind, invert):
    """file image -> in-memory image + st through the product's host C, and the engine's outputs for the same file"""
    from berserkerquake2_b200 import _capi as c
    file = np.ascontiguousarray(G["loader/md2/file"])
    h = synth.md2_header(file)
    blob = np.zeros(int(h["ofs_end"]), np.uint8); st = np.zeros((int(h["num_st"]), 2), np.float32)
    assert c.lib.bq2_host_md2_from_file(c.ptr(file), file.nbytes, 1.0, invert, c.ptr(blob), blob.nbytes, c.ptr(st)) == 0
    key = f"loader/md2/{kind}{invert}"
    want = {k: G[f"{key}/{k}"] for k in ("image", "neighbours", "tangents", "binormals", "cache_file")}
    if kind == "flat":
        want["normals"] = G[f"{key}/normals"]
    V, F, fs, of = int(h["num_xyz"]), int(h["num_frames"]), int(h["framesize"]), int(h["ofs_frames"])
    for f in range(F):      # what the engine's loader writes of a frame: scale + translate, vertex bytes
        o = of + f * fs
        assert np.array_equal(blob[o:o + 24], want["image"][o:o + 24])
        assert np.array_equal(blob[o + 40:o + 40 + 4 * V], want["image"][o + 40:o + 40 + 4 * V])
    assert np.array_equal(synth.md2_tris(blob), synth.md2_tris(want["image"]))
    return blob, st, want


def loader_md3(invert, k):
    from berserkerquake2_b200 import _capi as c
    file = np.ascontiguousarray(G["loader/md3/file"])
    key = f"loader/md3/inv{invert}/mesh{k}"
    want = np.ascontiguousarray(G[f"{key}/verts"]).view(synth.MD3_VERTEX).reshape(G[f"{key}/verts"].shape[:2])
    F, V = want.shape; T = len(G[f"{key}/indexes"])
    verts = np.zeros((F, V), synth.MD3_VERTEX); idx = np.zeros((T, 3), np.uint16); st = np.zeros((V, 2), np.float32)
    tr = np.zeros((F, 3), np.float32)
    assert c.lib.bq2_host_md3_mesh_from_file(c.ptr(file), file.nbytes, k, 1.0, invert, c.ptr(tr), c.ptr(verts), c.ptr(idx), c.ptr(st)) == 0
    assert same(verts["point"], want["point"]) and np.array_equal(verts["normal"], want["normal"])
    assert np.array_equal(idx, G[f"{key}/indexes"]) and same(st, G[f"{key}/st"])
    assert same(tr, G[f"loader/md3/inv{invert}/frame_translate"])
    return file, verts, idx, st, want, G[f"{key}/neighbours"], G[f"{key}/cache_file"]


@pytest.mark.parametrize("kind", ["smooth", "flat"])
@pytest.mark.parametrize("invert", [0, 1])
def test_oracle_md2_loader(oracle, kind, invert):
    from berserkerquake2_b200 import _capi as c
    blob, st, want = loader_md2(kind, invert)
    assert np.array_equal(oracle.md2_neighbours(blob), want["neighbours"])
    assert np.array_equal(host.md2_neighbours(blob), want["neighbours"])
    cosine = float(G["loader/smooth_cosine"][0])
    if kind == "smooth":
        tn, bn = oracle.md2_tangents_smooth(blob, st, cosine); nm = None
    else:
        nm, tn, bn = oracle.md2_tangents_flat(blob, st)
        assert np.array_equal(nm, want["normals"])
    assert np.array_equal(tn, want["tangents"]) and np.array_equal(bn, want["binormals"])
    F, rows = tn.shape; T = len(want["neighbours"])
    buf = np.zeros(want["cache_file"].size, np.uint8)
    assert c.lib.bq2_md2_cache_bytes(int(kind == "smooth"), T, rows, F) == buf.size
    assert c.lib.bq2_md2_cache_write(c.ptr(buf), buf.size, int(kind == "smooth"), cosine, c.ptr(want["neighbours"]), T, rows, F,
                                     c.ptr(bn), c.ptr(tn), c.ptr(nm)) == 0
    assert np.array_equal(buf, want["cache_file"])          # the engine's cache file, byte for byte


@pytest.mark.parametrize("invert", [0, 1])
def test_oracle_md3_loader(oracle, invert):
    from berserkerquake2_b200 import _capi as c
    for k in range(2):
        file, verts, idx, st, want, nb, cache = loader_md3(invert, k)
        assert np.array_equal(oracle.md3_neighbours(idx), nb) and np.array_equal(host.md3_neighbours(idx), nb)
        oracle.md3_tangents(verts, st, idx)
        assert np.array_equal(verts["tangent"], want["tangent"]) and np.array_equal(verts["binormal"], want["binormal"])
        F, V = verts.shape
        buf = np.zeros(cache.size, np.uint8)
        crc = int(c.lib.bq2_block_checksum(c.ptr(file), file.nbytes))
        assert c.lib.bq2_md3_cache_write(c.ptr(buf), buf.size, crc, c.ptr(verts), F, V) == 0
        assert np.array_equal(buf, cache)


# ---------------------------------------------------------------- GPU: the CUDA library vs the same vectors
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["md2_small", "md2_c1"])
def test_gpu_md2_pairs(gpu_ctx, name):
    blob, nb, we, wl, pe, pl = md2_case(name)
    V, T, F = oracle_lib.md2_dims(blob)
    mid = gpu_ctx.upload_md2(blob, host.md2_neighbours(blob))
    we = we.copy(); we["model"] = mid
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl)
    res = gpu_ctx.frame(ents, pairs)
    for p in range(len(pe)):
        assert same(pairs["light"][p], G[f"{name}/pair{p}/light_local"])
        assert same(res.lerped(int(pe[p]), V), G[f"{name}/pair{p}/lerped"])
        assert same(res.extruded(p, V), G[f"{name}/pair{p}/extruded"])
        assert np.array_equal(res.facing(p, T), G[f"{name}/pair{p}/viss"])
        tv = res.trisverts(p)                                   # indices dereferenced = the reference's TrisVerts
        assert same(tv, G[f"{name}/pair{p}/tris_verts"])
        assert int(res.index_counts[p]) == len(G[f"{name}/pair{p}/tris_verts"])


@pytest.mark.gpu
def test_gpu_md2_broken_mesh(gpu_ctx):
    blob, nb = G["md2_broken/blob"], G["md2_broken/neighbours"]
    V, T, F = oracle_lib.md2_dims(blob)
    mid = gpu_ctx.upload_md2(blob, nb)
    used = np.unique(oracle_lib.md2_idx(blob))
    z = np.zeros(3, np.float32)
    ents = np.array([host.entity_md2(blob, 1, 0, 0.25, z, z, z, model=mid)], bq2.ENTITY_DTYPE)
    pairs = np.zeros(2, bq2.PAIR_DTYPE)
    for i in range(2):
        pairs["entity"][i] = 0; pairs["light"][i] = G[f"md2_broken/pair{i}/light"]; pairs["scale"][i] = 32 * 300.0
    res = gpu_ctx.frame(ents, pairs)
    for i in range(2):
        assert same(res.lerped(0, V), G[f"md2_broken/pair{i}/lerped"])
        assert same(res.extruded(i, V)[used], G[f"md2_broken/pair{i}/extruded"][used])
        assert np.array_equal(res.facing(i, T), G[f"md2_broken/pair{i}/viss"])
        assert same(res.trisverts(i), G[f"md2_broken/pair{i}/tris_verts"])


@pytest.mark.gpu
def test_gpu_shell(gpu_ctx):
    blob = synth.md2(8, 5, 4, 21)
    mid = gpu_ctx.upload_md2(blob, host.md2_neighbours(blob))
    for i in range(2):
        e = host.entity_md2(blob, 2, 1, 0.375, (1, 2, 3), (2, 2, 2.5), (0, 25, 0), rf_flags=int(G[f"md2_shell/{i}/flags"][0]), model=mid)
        res = gpu_ctx.frame(np.array([e], bq2.ENTITY_DTYPE), None)
        assert same(res.lerped(0, 34), G[f"md2_shell/{i}/lerped"])


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["smooth", "flat"])
def test_gpu_md2_light_family(gpu_ctx, kind):
    tag = f"light/md2_{kind}"
    blob, tan, bi = G[f"{tag}/blob"], G[f"{tag}/tangents"], G[f"{tag}/binormals"]
    nm = G[f"{tag}/normals"] if kind == "flat" else None
    idx = oracle_lib.md2_idx(blob); V, T, F = oracle_lib.md2_dims(blob)
    mid = gpu_ctx.upload_md2(blob, host.md2_neighbours(blob), tan, bi, nm, smooth=(kind == "smooth"))
    z = np.zeros(3, np.float32)
    for i in range(2):
        f, of = [int(x) for x in G[f"{tag}/{i}/frames"]]; bl = float(G[f"{tag}/{i}/backlerp"][0])
        ents = np.array([host.entity_md2(blob, f, of, bl, z, z, z, model=mid)], bq2.ENTITY_DTYPE)
        pairs = np.zeros(1, bq2.PAIR_DTYPE)
        pairs["light"][0] = G["light/light"]; pairs["view"][0] = G["light/view"]; pairs["scale"][0] = 9600
        want = bq2.WANT_SHADOW | bq2.WANT_NTB | bq2.WANT_TSLIGHT
        res = gpu_ctx.frame(ents, pairs, want=want)
        rows = V if kind == "smooth" else T
        row = idx.reshape(-1) if kind == "smooth" else np.repeat(np.arange(T), 3)
        N, T_, B = res.ntb(0, rows)
        assert same(N[row], G[f"{tag}/{i}/normal"]) and same(T_[row], G[f"{tag}/{i}/tangent"]) and same(B[row], G[f"{tag}/{i}/binormal"])
        if kind == "smooth":        # the reference expands per corner; per-vertex rows gathered by index are identical
            lp, ts = res.tslight(0, V)
            assert same(lp[row], G[f"{tag}/{i}/lightp"]) and same(ts[row], G[f"{tag}/{i}/tsh"])
        else:                       # per-corner rows, exactly the reference's arrays
            lp, ts = res.tslight(0, 3 * T)
            assert same(lp, G[f"{tag}/{i}/lightp"]) and same(ts, G[f"{tag}/{i}/tsh"])


@pytest.mark.gpu
def test_gpu_md3(gpu_ctx):
    ft, meshes, we, wl = md3_case()
    mid = gpu_ctx.upload_md3(ft, meshes)
    we = we.copy(); we["model"] = mid
    pe = np.repeat(np.arange(3, dtype=np.int32), 2); pl = np.arange(6, dtype=np.int32)
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl)
    res = gpu_ctx.frame(ents, pairs)
    for p in range(6):
        for k in (0, 2):
            V, T = meshes[k]["verts"].shape[1], len(meshes[k]["indexes"])
            assert same(res.lerped(3 * int(pe[p]) + k, V), G[f"md3/pair{p}/mesh{k}/lerped"])
            assert same(res.extruded(3 * p + k, V), G[f"md3/pair{p}/mesh{k}/extruded"])
            assert np.array_equal(res.facing(3 * p + k, T), G[f"md3/pair{p}/mesh{k}/viss"])
            assert same(res.trisverts(3 * p + k), G[f"md3/pair{p}/mesh{k}/tris_verts"])
        assert int(res.index_counts[3 * p + 1]) == 0            # the non-casting surface


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["smooth", "flat"])
@pytest.mark.parametrize("invert", [0, 1])
def test_gpu_md2_loader(gpu_ctx, kind, invert):
    """the device's neighbour and tangent-space precompute against what the engine's loader produced"""
    blob, st, want = loader_md2(kind, invert)
    assert np.array_equal(gpu_ctx.gpu_md2_neighbours(blob), want["neighbours"])
    nm, tn, bn = gpu_ctx.gpu_md2_tangents(blob, st, float(G["loader/smooth_cosine"][0]), smooth=(kind == "smooth"))
    assert np.array_equal(tn, want["tangents"]) and np.array_equal(bn, want["binormals"])
    if kind == "flat":
        assert np.array_equal(nm, want["normals"])


@pytest.mark.gpu
@pytest.mark.parametrize("invert", [0, 1])
def test_gpu_md3_loader(gpu_ctx, invert):
    for k in range(2):
        file, verts, idx, st, want, nb, cache = loader_md3(invert, k)
        assert np.array_equal(gpu_ctx.gpu_md3_neighbours(idx), nb)
        gpu_ctx.gpu_md3_tangents(verts, st, idx)
        assert np.array_equal(verts["tangent"], want["tangent"]) and np.array_equal(verts["binormal"], want["binormal"])
