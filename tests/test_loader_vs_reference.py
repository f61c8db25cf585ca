"""SURVEY 8f rows 1-3 pinned by EXECUTION of the engine's own loaders: Mod_LoadAliasModel (M:51217) and
Mod_LoadAliasMD3Model (M:50584) are run whole, headless, from the unmodified reference object code on synthetic
.md2 / .md3 file images.  What they leave in memory and in their cache files is compared byte for byte with
 - the product's host C (file -> in-memory image, neighbour tables, the cache-file writers),
 - the oracle's restatement of the ASSEMBLED tangent / binormal / normal loops (M:51661-51804, 50956-51015), which is
   what tests/test_gpu_tangents.py holds the device kernels to.
CPU only; skipped where oracle/_ref is not built."""
import ctypes as C
import os

import numpy as np
import pytest

from berserkerquake2_b200 import _capi as c, synth, host
from scenes import bits
from test_formats import build_md3_file, checksum, md2_from_file

L = c.lib


def gamedir(tmp_path, fx=None, modname="models/t/tris.md2"):
    base = tmp_path / "baseq2"
    (base / os.path.dirname(modname)).mkdir(parents=True, exist_ok=True)
    if fx is not None:
        (base / (modname[:-3] + "fx" + modname[-1])).write_text(fx)      # tris.fx2 / tris.fx3 (M:51355-51358)
    return str(tmp_path)


def md2_with_coincident_verts(seed):
    """the synthetic blob with a few vertices moved onto others (same bytes in every frame): the loader merges the
    tangent sums of such vertices when their normals are within r_smoothangle (M:51719-51735)"""
    blob = synth.md2(10, 7, 4, seed)
    h = synth.md2_header(blob)
    V, F, fs, of = int(h["num_xyz"]), int(h["num_frames"]), int(h["framesize"]), int(h["ofs_frames"])
    for f in range(F):
        rows = blob[of + f * fs + 40: of + f * fs + 40 + 4 * V].reshape(V, 4)
        rows[5] = rows[6]                   # same position, same normal index: merged
        rows[20, :3] = rows[21, :3]         # same position, its own normal index: merged only if within 45 degrees
        rows[30, :3] = rows[3, :3]; rows[30, 3] = (int(rows[3, 3]) + 81) % 162      # same position, far normal
        rows[31] = rows[6]                  # a third member of the first group
    return blob


@pytest.mark.parametrize("smooth", [True, False])
@pytest.mark.parametrize("invert,scale", [(False, 1.0), (True, 1.0), (False, 0.6)])
def test_md2_loader(ref, oracle, tmp_path, smooth, invert, scale):
    file = md2_with_coincident_verts(11)
    work = gamedir(tmp_path, None if smooth else "unsmoothvecs\n")
    got = ref.load_md2(file, work, "models/t/tris.md2", scale, invert)
    assert got["smooth"] == smooth
    h = synth.md2_header(file); V, T, F = int(h["num_xyz"]), int(h["num_tris"]), int(h["num_frames"])

    # file -> in-memory image (product host C).  The engine leaves the name field of each frame and the st / glcmds
    # areas of the image unwritten (M:51297, 51329-51345 with cache = true): compare what it does write.
    rc, blob, st = md2_from_file(file, scale, invert)
    assert rc == 0
    assert np.array_equal(blob[:68], got["image"][:68])
    assert np.array_equal(synth.md2_tris(blob), synth.md2_tris(got["image"]))
    fs, of = int(h["framesize"]), int(h["ofs_frames"])
    img = got["image"]
    for f in range(F):
        o = of + f * fs
        assert np.array_equal(blob[o:o + 24], img[o:o + 24]), f                       # scale, translate
        # vertex bytes; with invert the lightnormalindex bytes went through Mod_InvertNormal (M:61850) on both sides
        assert np.array_equal(blob[o + 40:o + 40 + 4 * V], img[o + 40:o + 40 + 4 * V]), f

    # neighbour table: engine == oracle == product host C
    assert np.array_equal(got["neighbours"], oracle.md2_neighbours(blob))
    assert np.array_equal(got["neighbours"], host.md2_neighbours(blob))

    # the assembled tangent loops: engine == oracle restatement
    cosine = ref.smooth_cosine()
    if smooth:
        tn, bn = oracle.md2_tangents_smooth(blob, st, cosine); nm = None
        assert np.array_equal(got["tangents"], tn) and np.array_equal(got["binormals"], bn)
        rows = V
    else:
        nm, tn, bn = oracle.md2_tangents_flat(blob, st)
        assert np.array_equal(got["normals"], nm)
        assert np.array_equal(got["tangents"], tn) and np.array_equal(got["binormals"], bn)
        rows = T

    # the cache file the engine wrote == the product's writer, byte for byte; the product's reader accepts the engine's
    path = os.path.join(work, "baseq2", "cache", ("inv_" if invert else "") + "models/t/tris.md2")
    theirs = np.fromfile(path, np.uint8)
    n = L.bq2_md2_cache_bytes(int(smooth), T, rows, F)
    assert n == theirs.size
    buf = np.zeros(n, np.uint8)
    assert L.bq2_md2_cache_write(c.ptr(buf), n, int(smooth), cosine, c.ptr(got["neighbours"]), T, rows, F,
                                 c.ptr(bn), c.ptr(tn), c.ptr(nm)) == 0
    assert np.array_equal(buf, theirs)
    nb2 = np.zeros((T, 3), np.int32); o_ = [np.zeros((F, rows), np.uint8) for _ in range(3)]; why = C.c_int32(-1)
    assert L.bq2_md2_cache_read(c.ptr(theirs), theirs.size, int(smooth), cosine, T, rows, F, c.ptr(nb2), c.ptr(o_[0]),
                                c.ptr(o_[1]), c.ptr(o_[2]), C.byref(why)) == 0
    assert np.array_equal(nb2, got["neighbours"]) and np.array_equal(o_[0], bn) and np.array_equal(o_[1], tn)

    # second load: the engine reads its own cache file back ("loaded from cache") and ends with the same tables
    again = ref.load_md2(file, work, "models/t/tris.md2", scale, invert)
    for k in ("neighbours", "tangents", "binormals"):
        assert np.array_equal(again[k], got[k]), k

    # ... and it accepts a cache file written by the product (tables from elsewhere: it must not recompute)
    marked = tn.copy(); marked[0, 0] = (int(marked[0, 0]) + 1) % 162
    assert L.bq2_md2_cache_write(c.ptr(buf), n, int(smooth), cosine, c.ptr(got["neighbours"]), T, rows, F,
                                 c.ptr(bn), c.ptr(marked), c.ptr(nm)) == 0
    buf.tofile(path)
    third = ref.load_md2(file, work, "models/t/tris.md2", scale, invert)
    assert np.array_equal(third["tangents"], marked)


def md3_file(seed, F=3):
    rng = np.random.RandomState(seed)
    meshes = []
    for k, (S, R) in enumerate(((8, 5), (6, 4))):
        verts, idx, st = synth.md3_mesh(S, R, F, seed + k, centre=(20.0 * k, 0, 0))
        pts = np.round(verts["point"] * 64).astype(np.int32)
        meshes.append(dict(points=pts, norm=rng.randint(-32768, 32767, pts.shape[:2]), tris=idx.astype(np.int32), st=st))
    ft = (rng.randn(F, 3) * 4).astype(np.float32)
    return build_md3_file(meshes, ft), len(meshes)


@pytest.mark.parametrize("invert,scale", [(False, 1.0), (True, 1.0), (False, 0.75)])
def test_md3_loader(ref, oracle, tmp_path, invert, scale):
    file, n_meshes = md3_file(40)
    work = gamedir(tmp_path, modname="models/t/test.md3")
    meshes, tr = ref.load_md3(file, work, "models/t/test.md3", scale, invert)
    assert len(meshes) == n_meshes
    F = tr.shape[0]
    for k, m in enumerate(meshes):
        V, T = m["verts"].shape[1], m["indexes"].shape[0]
        verts = np.zeros((F, V), synth.MD3_VERTEX); idx = np.zeros((T, 3), np.uint16); st = np.zeros((V, 2), np.float32)
        mytr = np.zeros((F, 3), np.float32)
        assert L.bq2_host_md3_mesh_from_file(c.ptr(file), file.nbytes, k, scale, int(invert), c.ptr(mytr), c.ptr(verts),
                                             c.ptr(idx), c.ptr(st)) == 0
        # file -> mesh: product host C == engine (points bit for bit, lat/lng normal byte, indices, st, frame translation)
        assert np.array_equal(bits(verts["point"]), bits(m["verts"]["point"]))
        assert np.array_equal(verts["normal"], m["verts"]["normal"])
        assert np.array_equal(idx, m["indexes"]) and np.array_equal(bits(st), bits(m["st"]))
        assert np.array_equal(bits(mytr), bits(tr))
        # neighbours: engine == oracle == product host C
        assert np.array_equal(m["neighbours"], oracle.md3_neighbours(idx))
        assert np.array_equal(m["neighbours"], host.md3_neighbours(idx))
        # the assembled tangent loop M:50993-51014: engine == oracle restatement
        oracle.md3_tangents(verts, st, idx)
        assert np.array_equal(verts["tangent"], m["verts"]["tangent"])
        assert np.array_equal(verts["binormal"], m["verts"]["binormal"])
        # cache file <model>_<mesh>: engine's bytes == product writer's
        path = os.path.join(work, "baseq2", "cache", ("inv_" if invert else "") + "models/t/test.md3_%d" % k)
        theirs = np.fromfile(path, np.uint8)
        n = L.bq2_md3_cache_bytes(F, V)
        assert n == theirs.size
        buf = np.zeros(n, np.uint8)
        assert L.bq2_md3_cache_write(c.ptr(buf), n, checksum(file), c.ptr(verts), F, V) == 0
        assert np.array_equal(buf, theirs)
    # second load comes from the cache files and ends with the same bytes
    again, _ = ref.load_md3(file, work, "models/t/test.md3", scale, invert)
    for a, b in zip(again, meshes):
        for k in ("normal", "tangent", "binormal"):
            assert np.array_equal(a["verts"][k], b["verts"][k]), k
