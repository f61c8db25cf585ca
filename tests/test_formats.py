"""the data formats either side of the path -- MD2 file -> in-memory image, the engine's
precompute cache files, Com_BlockChecksum.  CPU tests; the last one (gpu) runs a model from file bytes to shadow
volumes with every load-time table computed on the device and exchanged through a cache file."""
import ctypes as C

import numpy as np
import pytest

import berserkerquake2_b200 as bq2
from berserkerquake2_b200 import _capi as c, synth, host
from scenes import SMOOTH_COSINE, bits

L = c.lib


def checksum(buf):
    a = np.ascontiguousarray(buf).view(np.uint8)
    return int(L.bq2_block_checksum(c.ptr(a) if a.size else None, a.size))


def test_md4_published_vectors():
    """RFC 1320 appendix A.5 digests, folded the way Com_BlockChecksum (M:40589-40602) folds them."""
    kat = {b"": "31d6cfe0d16ae931b73c59d7e0c089c0", b"a": "bde52cb31de33e46245e05fbdbd6fb24",
           b"abc": "a448017aaf21d8525fc10ae87aa6729d", b"message digest": "d9130a8164549fe818874806e1c7014b",
           b"abcdefghijklmnopqrstuvwxyz": "d79e1c308aa5bbcdeea8ed63df412da9",
           b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789": "043f8582f241db351ce627e153e7f0e4",
           b"1234567890" * 8: "e33b4ddc9c38f2199c3e7b164fcc0536"}
    for msg, hexd in kat.items():
        w = np.frombuffer(bytes.fromhex(hexd), "<u4")
        assert checksum(np.frombuffer(msg, np.uint8)) == int(w[0] ^ w[1] ^ w[2] ^ w[3]), msg


def test_checksum_matches_reference(ref):
    ref.lib.bq2ref_block_checksum.restype = C.c_uint32
    ref.lib.bq2ref_block_checksum.argtypes = [C.c_void_p, C.c_int]
    rng = np.random.RandomState(0)
    for n in list(range(0, 130)) + [255, 256, 1000, 6144, 12288, 51084]:
        a = rng.randint(0, 256, n).astype(np.uint8)
        want = ref.lib.bq2ref_block_checksum(a.ctypes.data_as(C.c_void_p) if n else None, n)
        assert checksum(a) == want, n


def md2_from_file(file, scale=1.0, invert=False):
    h = synth.md2_header(file)
    blob = np.zeros(int(h["ofs_end"]), np.uint8)
    st = np.zeros((int(h["num_st"]), 2), np.float32)
    rc = L.bq2_host_md2_from_file(c.ptr(file), file.nbytes, scale, int(invert), c.ptr(blob), blob.nbytes, c.ptr(st))
    return rc, blob, st


def test_md2_file_to_memory():
    file = synth.md2(8, 5, 3, 4)                      # the generator emits the on-disk layout with scale 1
    rc, blob, st = md2_from_file(file)
    assert rc == 0 and np.array_equal(blob, file) and np.array_equal(st, synth.md2_st(file))
    rc, blob, _ = md2_from_file(file, scale=2.5)
    assert rc == 0
    h = synth.md2_header(file)
    for f in range(3):
        o = int(h["ofs_frames"]) + f * int(h["framesize"])
        assert np.array_equal(blob[o:o + 24].view("<f4"), file[o:o + 24].view("<f4") * np.float32(2.5))
        assert np.array_equal(blob[o + 24:o + int(h["framesize"])], file[o + 24:o + int(h["framesize"])])
    rc, inv, _ = md2_from_file(file, invert=True)      # winding reversed, y mirrored (M:51274-51281, 51303-51307)
    assert rc == 0
    assert np.array_equal(synth.md2_tris(inv)[:, :3], synth.md2_tris(file)[:, 2::-1])
    assert np.array_equal(synth.md2_tris(inv)[:, 3:], synth.md2_tris(file)[:, :2:-1])
    o = int(h["ofs_frames"])
    a, b = inv[o:o + 24].view("<f4"), file[o:o + 24].view("<f4")
    assert a[1] == -b[1] and a[4] == -b[4] and a[0] == b[0] and a[5] == b[5]
    # the loader's checks (Com_Error in the reference, status codes here)
    bad = file.copy(); bad[4:8].view("<i4")[0] = 7
    assert md2_from_file(bad)[0] == -1                                   # wrong version number
    bad = file.copy(); bad[24:28].view("<i4")[0] = 5000
    assert md2_from_file(bad)[0] == -4                                   # too many vertices
    bad = file.copy(); bad[32:36].view("<i4")[0] = 0
    assert md2_from_file(bad)[0] == -1                                   # no triangles
    assert md2_from_file(file, scale=0.0)[0] == -1                       # invalid scale
    assert md2_from_file(file[:100])[0] == -1                            # truncated


@pytest.mark.parametrize("smooth", [True, False])
def test_md2_cache_roundtrip_and_validation(oracle, smooth):
    blob = synth.md2(10, 7, 4, 6); st = synth.md2_st(blob)
    h = synth.md2_header(blob); V, T, F = int(h["num_xyz"]), int(h["num_tris"]), int(h["num_frames"])
    nb = host.md2_neighbours(blob)
    if smooth:
        nm = None; tn, bn = oracle.md2_tangents_smooth(blob, st, SMOOTH_COSINE); rows = V
    else:
        nm, tn, bn = oracle.md2_tangents_flat(blob, st); rows = T
    n = L.bq2_md2_cache_bytes(int(smooth), T, rows, F)
    assert n == 1 + 4 + 4 + 12 * T + (2 if smooth else 3) * (4 + rows * F)
    buf = np.zeros(n, np.uint8)
    assert L.bq2_md2_cache_write(c.ptr(buf), n, int(smooth), SMOOTH_COSINE, c.ptr(nb), T, rows, F, c.ptr(bn), c.ptr(tn), c.ptr(nm)) == 0
    # layout, field by field (M:51811-51838)
    assert buf[0] == int(smooth) and buf[1:5].view("<u4")[0] == int(np.float32(SMOOTH_COSINE) * np.float32(0x7fffffff))
    assert buf[5:9].view("<u4")[0] == checksum(nb) and np.array_equal(buf[9:9 + 12 * T].view("<i4").reshape(T, 3), nb)
    o = 9 + 12 * T
    assert buf[o:o + 4].view("<u4")[0] == checksum(bn) and np.array_equal(buf[o + 4:o + 4 + rows * F], bn.reshape(-1))

    def read(b, sm=smooth, cosine=SMOOTH_COSINE):
        nb2 = np.zeros((T, 3), np.int32); o_ = [np.zeros((F, rows), np.uint8) for _ in range(3)]
        why = C.c_int32(-1)
        rc = L.bq2_md2_cache_read(c.ptr(b), b.nbytes, int(sm), cosine, T, rows, F, c.ptr(nb2), c.ptr(o_[0]), c.ptr(o_[1]), c.ptr(o_[2]), C.byref(why))
        return rc, why.value, nb2, o_
    rc, why, nb2, (b2, t2, n2) = read(buf)
    assert rc == 0 and why == 0 and np.array_equal(nb2, nb) and np.array_equal(b2, bn) and np.array_equal(t2, tn)
    if not smooth:
        assert np.array_equal(n2, nm)
    assert read(buf, sm=not smooth)[:2] == (-1, 1)                       # smooth flag differs -> invalid cache
    assert read(buf, cosine=0.5)[:2] == (-1, 1)                          # r_smoothangle changed -> invalid cache
    assert read(buf[:-1].copy())[:2] == (-1, 1)                          # truncated
    bad = buf.copy(); bad[20] ^= 1
    assert read(bad)[:2] == (-1, 1)                                      # neighbour table corrupted
    bad = buf.copy(); bad[-1] ^= 1
    assert read(bad)[:2] == (-1, 2)                                      # "invalid checksum!"


def test_md3_cache_roundtrip(oracle):
    verts, idx, st = synth.md3_mesh(8, 5, 3, 2)
    oracle.md3_tangents(verts, st, idx)
    F, V = verts.shape
    n = L.bq2_md3_cache_bytes(F, V)
    assert n == 4 + 3 * F * V
    buf = np.zeros(n, np.uint8)
    assert L.bq2_md3_cache_write(c.ptr(buf), n, 0xdeadbeef, c.ptr(verts), F, V) == 0
    assert buf[:4].view("<u4")[0] == 0xdeadbeef
    assert np.array_equal(buf[4:].reshape(F * V, 3), np.stack([verts["normal"], verts["tangent"], verts["binormal"]], -1).reshape(-1, 3))
    back = verts.copy(); back["normal"] = 0; back["tangent"] = 0; back["binormal"] = 0
    assert L.bq2_md3_cache_read(c.ptr(buf), n, 0xdeadbeef, c.ptr(back), F, V) == 0
    assert np.array_equal(back, verts)
    assert L.bq2_md3_cache_read(c.ptr(buf), n, 0x12345678, c.ptr(back), F, V) == -1      # another .md3 file
    assert L.bq2_md3_cache_read(c.ptr(buf), n - 1, 0xdeadbeef, c.ptr(back), F, V) == -1  # truncated


@pytest.mark.gpu
def test_md2_blob_entry_points_refuse_malformed_headers(gpu_ctx):
    """bq2_upload_md2 and bq2_gpu_md2_tangents share one header check: negative section offsets (which wrap to small
    numbers in unsigned arithmetic and used to pass the bounds test of the tangent entry point), a wrong version, a
    frame size too small for the vertex count -- all refused, none read."""
    rc, blob, st = md2_from_file(synth.md2(8, 5, 2, 1), scale=1.0)
    assert rc == 0
    h = synth.md2_header(blob); V, T, F = int(h["num_xyz"]), int(h["num_tris"]), int(h["num_frames"])
    tn = np.zeros((F, V), np.uint8); bn = np.zeros((F, V), np.uint8); mid = C.c_int32(-1)
    for field, value in (("ofs_tris", -12), ("ofs_frames", -40), ("ofs_tris", -(2 ** 31)), ("version", 7), ("framesize", 8),
                         ("num_frames", 0), ("num_xyz", 4096), ("ofs_frames", 2 ** 30)):
        bad = blob.copy()
        bad[:68].view(synth.MD2_HEADER)[0][field] = value
        assert L.bq2_gpu_md2_tangents(gpu_ctx._h, c.ptr(bad), bad.nbytes, c.ptr(st), SMOOTH_COSINE, 1, None, c.ptr(tn), c.ptr(bn)) < 0, field
        assert L.bq2_upload_md2(gpu_ctx._h, c.ptr(bad), bad.nbytes, None, None, None, None, 0, C.byref(mid)) < 0, field
    assert L.bq2_gpu_md2_tangents(gpu_ctx._h, c.ptr(blob), blob.nbytes, c.ptr(st), SMOOTH_COSINE, 1, None, c.ptr(tn), c.ptr(bn)) == 0


@pytest.mark.gpu
def test_file_to_shadow_volume_with_device_precompute(gpu_ctx, oracle):
    """file bytes -> bq2_host_md2_from_file -> neighbours + tangent bytes on the device -> cache file -> read back
    -> upload -> frame: the same index buffers and N/T/B rows as the oracle on the oracle's own tables."""
    from scenes import world_scene
    import oracle_lib
    file = synth.md2(12, 9, 5, 33)
    rc, blob, st = md2_from_file(file, scale=1.5)
    assert rc == 0
    h = synth.md2_header(blob); V, T, F = int(h["num_xyz"]), int(h["num_tris"]), int(h["num_frames"])
    nb = gpu_ctx.gpu_md2_neighbours(blob)
    _, tn, bn = gpu_ctx.gpu_md2_tangents(blob, st, SMOOTH_COSINE, smooth=True)
    n = L.bq2_md2_cache_bytes(1, T, V, F)
    cache = np.zeros(n, np.uint8)
    assert L.bq2_md2_cache_write(c.ptr(cache), n, 1, SMOOTH_COSINE, c.ptr(nb), T, V, F, c.ptr(bn), c.ptr(tn), None) == 0
    nb2 = np.zeros((T, 3), np.int32); bn2 = np.zeros((F, V), np.uint8); tn2 = np.zeros((F, V), np.uint8); why = C.c_int32()
    assert L.bq2_md2_cache_read(c.ptr(cache), n, 1, SMOOTH_COSINE, T, V, F, c.ptr(nb2), c.ptr(bn2), c.ptr(tn2), None, C.byref(why)) == 0
    assert np.array_equal(nb2, oracle.md2_neighbours(blob))
    ot, ob = oracle.md2_tangents_smooth(blob, st, SMOOTH_COSINE)
    assert np.array_equal(tn2, ot) and np.array_equal(bn2, ob)
    mid = gpu_ctx.upload_md2(blob, nb2, tn2, bn2, None, smooth=True)
    we, wl, pe, pl = world_scene(3, 2, F, seed=44)
    we["model"] = mid
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl)
    res = gpu_ctx.frame(ents, pairs, want=bq2.WANT_SHADOW | bq2.WANT_NTB)
    idx = oracle_lib.md2_idx(blob)
    for p in range(len(pe)):
        W = we[pe[p]]
        w = oracle.md2_pair(blob, nb2, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"], W["oldorigin"],
                            W["angles"], wl["origin"][pl[p]], 300.0, idx=idx)
        assert np.array_equal(res.indices(p), w["indices"]) and np.array_equal(bits(res.lerped(int(pe[p]), V)), bits(w["lerped"]))


# ---------------------------------------------------------------------------------------------- MD3 files
def build_md3_file(meshes, frame_translate, frame_extra=None):
    """A Quake III .md3 image (defs.h:4298-4391): header, frames, then each mesh = header, skins, triangles,
    texture coordinates, vertices (shorts + packed lat/lng normal)."""
    import struct
    F = len(frame_translate)
    frames = b"".join(struct.pack("<10f16s", -1, -2, -3, 1, 2, 3, *map(float, t), 9.0, b"test") for t in frame_translate)
    body = b""
    for m in meshes:
        V, T = m["points"].shape[1], len(m["tris"])
        skins = struct.pack("<64si", b"models/x/skin.tga", 0)
        tris = np.asarray(m["tris"], "<u4").tobytes()
        tcs = np.asarray(m["st"], "<f4").tobytes()
        verts = np.zeros((F, V, 4), "<i2"); verts[:, :, :3] = m["points"]; verts[:, :, 3] = m["norm"]
        ofs_tris = 108 + len(skins); ofs_tcs = ofs_tris + len(tris); ofs_verts = ofs_tcs + len(tcs)
        size = ofs_verts + verts.nbytes
        body += struct.pack("<4s64s11i", b"IDP3", b"mesh", 0, F, 1, V, T, ofs_tris, 108, ofs_tcs, ofs_verts, size, 0)[:108]
        body += skins + tris + tcs + verts.tobytes()
    ofs_frames = 108
    hdr = struct.pack("<ii64s9i", 0x33504449, 15, b"test.md3", 0, F, 0, len(meshes), 0, ofs_frames, ofs_frames + len(frames),
                      ofs_frames + len(frames), ofs_frames + len(frames) + len(body))
    assert len(hdr) == 108
    return np.frombuffer(hdr + frames + body, np.uint8).copy()


def test_md3_file_to_memory(oracle):
    import ctypes.util
    libm = C.CDLL(ctypes.util.find_library("m"))
    libm.sincosf.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    rng = np.random.RandomState(8)
    F = 3
    meshes = []
    for V, T in ((20, 30), (7, 5)):
        meshes.append(dict(points=rng.randint(-3000, 3000, (F, V, 3)), norm=rng.randint(-32768, 32767, (F, V)),
                           tris=rng.randint(0, V, (T, 3)), st=rng.rand(V, 2).astype(np.float32)))
    ft = rng.randn(F, 3).astype(np.float32) * 10
    file = build_md3_file(meshes, ft)
    nf, nm = C.c_int32(), C.c_int32(); nv = np.zeros(32, np.int32); nt = np.zeros(32, np.int32)
    assert L.bq2_md3_file_info(c.ptr(file), file.nbytes, C.byref(nf), C.byref(nm), c.ptr(nv), c.ptr(nt)) == 0
    assert (nf.value, nm.value, list(nv[:2]), list(nt[:2])) == (3, 2, [20, 7], [30, 5])
    for invert in (False, True):
        for scale in (1.0, 0.75):
            for k, m in enumerate(meshes):
                V, T = nv[k], nt[k]
                verts = np.zeros((F, V), synth.MD3_VERTEX); idx = np.zeros((T, 3), np.uint16); st = np.zeros((V, 2), np.float32)
                tr = np.zeros((F, 3), np.float32)
                assert L.bq2_host_md3_mesh_from_file(c.ptr(file), file.nbytes, k, scale, int(invert), c.ptr(tr), c.ptr(verts), c.ptr(idx), c.ptr(st)) == 0
                inv = np.array([1.0 / 64 * np.float32(scale)] * 3).astype(np.float32)      # (1.0/64)*scale in double, stored as float
                if invert:
                    inv[1] = -inv[1]
                assert np.array_equal(bits(verts["point"]), bits(m["points"].astype(np.float32) * inv[None, None, :]))   # M:50968
                want_tr = ft * np.float32(scale)
                if invert:
                    want_tr[:, 1] = -want_tr[:, 1]
                assert np.array_equal(bits(tr), bits(want_tr))
                assert np.array_equal(idx, m["tris"][:, ::-1] if invert else m["tris"]) and np.array_equal(st, m["st"])
                for f in range(F):                                  # lat/lng -> vector -> Normal2Index, M:50972-50989
                    for v in range(V):
                        n16 = int(m["norm"][f, v])
                        lat = np.float32(np.float64(np.float32((n16 >> 8) & 0xff)) * (np.pi / 128))
                        lng = np.float32(np.float64(np.float32(n16 & 0xff)) * (np.pi / 128))
                        sl, cl, sg, cg = (C.c_float() for _ in range(4))
                        libm.sincosf(lat, C.byref(sl), C.byref(cl)); libm.sincosf(lng, C.byref(sg), C.byref(cg))
                        sl, cl, sg, cg = (np.float32(x.value) for x in (sl, cl, sg, cg))
                        vec = np.array([cl * sg, (-sl * sg) if invert else sl * sg, cg], np.float32)
                        assert verts["normal"][f, v] == oracle.normal2index(vec), (f, v)
                assert not verts["tangent"].any() and not verts["binormal"].any()
    bad = file.copy(); bad[4:8].view("<i4")[0] = 14
    assert L.bq2_md3_file_info(c.ptr(bad), bad.nbytes, None, None, None, None) == -1            # wrong version number
    bad = file.copy(); bad[108 + 3 * 56:108 + 3 * 56 + 4] = np.frombuffer(b"XXXX", np.uint8)
    assert L.bq2_md3_file_info(c.ptr(bad), bad.nbytes, None, None, None, None) == -1            # mesh has wrong id
    assert L.bq2_md3_file_info(c.ptr(file), 300, None, None, None, None) == -1                  # truncated


def test_malformed_and_unaligned_file_images(oracle):
    """File images are untrusted bytes: offsets that leave the file (negative ones included) are refused, and a file
    whose sections sit at odd offsets -- which the engine's x86 loader reads without complaint -- converts to the same
    arrays as its aligned twin (the parsers read multi-byte fields through memcpy; tools/fuzz_host_parsers.py runs
    them under ASan / UBSan)."""
    rng = np.random.RandomState(3)
    F, V, T = 2, 6, 4
    mesh = dict(points=rng.randint(-3000, 3000, (F, V, 3)), norm=rng.randint(-32768, 32767, (F, V)),
                tris=rng.randint(0, V, (T, 3)), st=rng.rand(V, 2).astype(np.float32))
    file = build_md3_file([mesh, mesh], rng.randn(F, 3).astype(np.float32))
    info = lambda f, n=None: L.bq2_md3_file_info(c.ptr(f), f.nbytes if n is None else n, None, None, None, None)
    assert info(file) == 0
    for field_ofs, value in ((100, -100), (100, -(1 << 31)), (100, file.nbytes - 50), (92, -8), (92, file.nbytes)):   # ofs_meshes, ofs_frames
        bad = file.copy(); bad[field_ofs:field_ofs + 4].view("<i4")[0] = value
        assert info(bad) == -1, (field_ofs, value)
    m0 = 108 + F * 56                                              # first mesh header: ... num_verts +80, num_tris +84, ofs_tris +88, meshsize +104
    for field_ofs, value in ((m0 + 88, -4), (m0 + 88, file.nbytes), (m0 + 96, 1 << 30), (m0 + 104, 0), (m0 + 104, -5), (m0 + 104, 1 << 30)):
        bad = file.copy(); bad[field_ofs:field_ofs + 4].view("<i4")[0] = value
        assert info(bad) == -1, (field_ofs, value)
    # the same file with one byte inserted in front of the frames: every later section sits at an odd offset
    odd = np.concatenate([file[:108], np.zeros(1, np.uint8), file[108:]])
    h = odd[:108].view("<i4"); h[23] += 1; h[24] += 1; h[25] += 1; h[26] += 1      # ofs_frames, ofs_tags, ofs_meshes, ofs_end
    assert info(odd) == 0
    for k in range(2):
        got = []
        for f in (file, odd):
            verts = np.zeros((F, V), synth.MD3_VERTEX); idx = np.zeros((T, 3), np.uint16); st = np.zeros((V, 2), np.float32)
            tr = np.zeros((F, 3), np.float32)
            assert L.bq2_host_md3_mesh_from_file(c.ptr(f), f.nbytes, k, 0.75, 1, c.ptr(tr), c.ptr(verts), c.ptr(idx), c.ptr(st)) == 0
            got.append((verts.tobytes(), idx.tobytes(), st.tobytes(), tr.tobytes()))
        assert got[0] == got[1]
    # MD2: sections at odd offsets, and offsets that leave the file
    blob = synth.md2(6, 5, 3, 9)
    hdr = blob[:68].view("<i4").copy()
    out = np.zeros(blob.nbytes + 8, np.uint8); st = np.zeros((int(hdr[7]), 2), np.float32)
    conv = lambda f, o, s: L.bq2_host_md2_from_file(c.ptr(f), f.nbytes, 0.5, 1, c.ptr(o), o.nbytes, c.ptr(s))
    assert conv(blob, out, st) == 0
    odd2 = np.concatenate([blob[:68], np.zeros(1, np.uint8), blob[68:]])
    odd2[:68].view("<i4")[11:17] += 1                                             # every ofs_* field
    out2 = np.zeros_like(out); st2 = np.zeros_like(st)
    assert conv(odd2, out2, st2) == 0
    assert np.array_equal(out[68:blob.nbytes], out2[69:blob.nbytes + 1]) and np.array_equal(bits(st), bits(st2))
    for field, value in ((13, -12), (14, -40), (12, -4), (16, blob.nbytes + 1), (14, blob.nbytes - 10), (4, -1), (10, 1 << 28)):
        bad = blob.copy(); bad[:68].view("<i4")[field] = value
        assert conv(bad, out, st) in (-1, -4), (field, value)


@pytest.mark.gpu
def test_md3_file_to_shadow_volume(gpu_ctx, oracle):
    """.md3 bytes -> meshes -> neighbours + tangent bytes on the device -> upload -> frame, against the oracle."""
    from scenes import world_scene
    F = 4
    meshes_in, raw = [], []
    for k in range(2):
        verts, idx, st = synth.md3_mesh(10, 7, F, 50 + k, centre=(30.0 * k, 0, 0))
        pts = np.round(verts["point"] * 64).astype(np.int32)
        meshes_in.append(dict(points=pts, norm=np.full(pts.shape[:2], 0x2040, np.int32), tris=idx.astype(np.int32), st=st))
    ft = np.array([[0, 0, 0], [1, 0, 2], [2, 1, 0], [0, 3, 1]], np.float32)
    file = build_md3_file(meshes_in, ft)
    tr = np.zeros((F, 3), np.float32)
    meshes = []
    for k in range(2):
        V, T = meshes_in[k]["points"].shape[1], len(meshes_in[k]["tris"])
        verts = np.zeros((F, V), synth.MD3_VERTEX); idx = np.zeros((T, 3), np.uint16); st = np.zeros((V, 2), np.float32)
        assert L.bq2_host_md3_mesh_from_file(c.ptr(file), file.nbytes, k, 1.0, 0, c.ptr(tr), c.ptr(verts), c.ptr(idx), c.ptr(st)) == 0
        gpu_ctx.gpu_md3_tangents(verts, st, idx)
        want = verts.copy(); oracle.md3_tangents(want, st, idx)
        assert np.array_equal(verts, want)
        meshes.append(dict(verts=verts, indexes=idx, neighbours="gpu"))
    mid = gpu_ctx.upload_md3(tr, meshes)
    we, wl, pe, pl = world_scene(2, 2, F, seed=21)
    we["model"] = mid
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl)
    res = gpu_ctx.frame(ents, pairs)
    for p in range(len(pe)):
        W = we[pe[p]]
        for k in range(2):
            nb = oracle.md3_neighbours(meshes[k]["indexes"])
            w = oracle.md3_mesh_pair(meshes[k]["verts"], meshes[k]["indexes"], nb, tr, int(W["frame"]), int(W["oldframe"]),
                                     float(W["backlerp"]), W["origin"], W["oldorigin"], W["angles"], wl["origin"][pl[p]], 300.0)
            assert np.array_equal(res.indices(2 * p + k), w["indices"]), (p, k)
