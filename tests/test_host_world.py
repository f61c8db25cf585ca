"""CPU test of the host half of the world path (bq2_host_world_entities, csrc/bq2_host.c): what bq2_world_submit
decides on the host before anything reaches the GPU -- validation, kernel class, record and output-row numbering,
the flag filter M:63722-63727 / 64089, AngleVectors once per rotated entity.  No GPU, no compute call."""
import ctypes as C

import numpy as np

import berserkerquake2_b200 as bq2
from berserkerquake2_b200 import _capi as c, host

MODEL_ROW = np.dtype([("F", "<i4"), ("md3", "<i4"), ("first_mesh", "<i4"), ("n_mesh", "<i4"), ("V", "<i4"), ("T", "<i4"),
                      ("Tp", "<i4"), ("mflags", "<u4"), ("vpad", "<u4"), ("team", "<i4"), ("pad", "<i4", 2), ("cmax", "<i2", 8)])
# bq2_dev_went (csrc/bq2_internal.h): the caller's entity without its angles, then the host's numbering
XFORM = np.dtype([("model", "<i4"), ("frame", "<i4"), ("oldframe", "<i4"), ("rf_flags", "<u4"), ("backlerp", "<f4"),
                  ("origin", "<f4", 3), ("oldorigin", "<f4", 3), ("rec", "<i4"), ("pair_rec", "<i4"), ("vert_off", "<u4"),
                  ("nrm_off", "<u4"), ("basis", "<i4")])
STATS = np.dtype([("n_warp", "<i4"), ("n_team", "<i4"), ("warp_max_V", "<i4"), ("warp_max_T", "<i4"), ("warp_md2", "<i4"),
                  ("team_max_V", "<i4"), ("team_max_T", "<i4"), ("team_md2", "<i4"), ("lerped_verts_padded", "<u4"),
                  ("max_vpad", "<u4"), ("max_bitw", "<u4"), ("max_T", "<i4"), ("team_normal_rows", "<u4"), ("n_rotated", "<i4"),
                  ("total_verts", "<u8"), ("and_mflags", "<u4"), ("_tail", "<u4")])
assert MODEL_ROW.itemsize == 64 and XFORM.itemsize == 64 and STATS.itemsize == 72        # sizeof(bq2_world_stats): checked by test_struct_sizes
MESH_MD3 = 0x1


def run(we, rows, opts=None):
    fn = c.lib.bq2_host_world_entities
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    n = len(we)
    hx = np.zeros(n, XFORM); rec = np.full(n + 1, -7, np.int32); voff = np.zeros(n + 1, np.uint32); st = np.zeros(1, STATS)
    bases = np.full((n + 1, 9), np.nan, np.float32)                # the loop may write rows [0, rotated entities) only
    rc = fn(we.ctypes.data, n, rows.ctypes.data, len(rows), None if opts is None else C.addressof(opts), hx.ctypes.data,
            bases.ctypes.data, rec.ctypes.data, voff.ctypes.data, st.ctypes.data)
    if rc == 0:
        assert np.isnan(bases[st[0]["n_rotated"]:]).all()
        # the loop leaves each rotated entity's ANGLES in its row (the trigonometry runs afterwards, on the submit worker
        # when the ctx has one): rows [0, n_rotated) hold angles in [0..3) and nothing else yet
        nr = int(st[0]["n_rotated"])
        rot = [e for e in range(n) if we["angles"][e][0] or we["angles"][e][1] or we["angles"][e][2]]
        assert len(rot) == nr
        for r, e in enumerate(rot):
            assert np.array_equal(bases[r, :3].view(np.uint32), we["angles"][e].view(np.uint32)) and np.isnan(bases[r, 3:]).all()
        conv = c.lib.bq2_host_bases_from_angles
        conv.restype = None; conv.argtypes = [C.c_void_p, C.c_int32]
        conv(bases.ctypes.data, nr)
        for f in ("model", "frame", "oldframe", "rf_flags", "backlerp", "origin", "oldorigin"):      # copied bit for bit
            assert np.array_equal(hx[f].view(np.uint32), we[f].view(np.uint32)), f
    run.bases = bases
    return rc, hx, rec[:n], voff, st[0]


def test_struct_mirrors_have_the_librarys_sizes():
    """The numpy mirrors above are written into by C: a mirror shorter than its struct would be overrun silently."""
    sizes = (C.c_int32 * 3)()
    c.lib.bq2_host_world_struct_sizes(sizes)
    assert list(sizes) == [MODEL_ROW.itemsize, XFORM.itemsize, STATS.itemsize]


def rows_of(shapes):
    rows = np.zeros(len(shapes), MODEL_ROW)
    for i, (V, T, md3, n_mesh) in enumerate(shapes):
        rows[i] = (6, md3, i, n_mesh, V, T, (T + 31) & ~31, MESH_MD3 if md3 else 0, 0xdead, 7, (7, 7), (7,) * 8)
        c.lib.bq2_model_row_derive(C.c_void_p(rows[i:].ctypes.data))         # the fields the entity loop reads per entity
        team = int(V > 1024 or T > 1024)
        want = [0] * 8
        want[2 * team] = V; want[2 * team + 1] = T; want[4 + team] = int(not md3)
        assert rows["vpad"][i] == (V + 3) & ~3 and rows["team"][i] == team and list(rows["cmax"][i]) == want
    return rows


def test_numbering_classes_and_stats():
    rows = rows_of([(258, 512, 0, 1), (1026, 2048, 1, 1), (40, 60, 1, 1), (1024, 1024, 0, 1), (1025, 10, 0, 1)])
    rng = np.random.RandomState(5)
    n = 41
    we = np.zeros(n, bq2.WORLD_ENTITY_DTYPE)
    we["model"] = rng.randint(0, 5, n); we["frame"] = rng.randint(0, 6, n); we["oldframe"] = rng.randint(0, 6, n)
    we["backlerp"] = rng.rand(n); we["origin"] = rng.randn(n, 3) * 500; we["oldorigin"] = we["origin"] + rng.randn(n, 3)
    rc, hx, rec, voff, st = run(we, rows)
    assert rc == 0
    small = np.array([(rows["V"][m] <= 1024) and (rows["T"][m] <= 1024) for m in we["model"]])
    # warp-kernel records count up from 0 in entity order, team-kernel records down from n-1
    assert np.array_equal(rec[small], np.arange(small.sum()))
    assert np.array_equal(rec[~small], n - 1 - np.arange((~small).sum()))
    assert sorted(rec) == list(range(n)) and np.array_equal(hx["rec"], rec) and np.array_equal(hx["pair_rec"], rec)
    V = rows["V"][we["model"]]; T = rows["T"][we["model"]]; Tp = rows["Tp"][we["model"]]
    vpad = (V + 3) & ~3
    assert np.array_equal(voff[:n], np.concatenate([[0], np.cumsum(vpad)[:-1]])) and voff[n] == vpad.sum()
    assert np.array_equal(hx["vert_off"], voff[:n])
    assert (hx["basis"] == -1).all() and st["n_rotated"] == 0
    assert np.array_equal(hx["nrm_off"][~small], np.concatenate([[0], np.cumsum(Tp[~small])[:-1]]))
    assert (hx["nrm_off"][small] == 0).all()
    assert st["n_warp"] == small.sum() and st["n_team"] == (~small).sum()
    assert st["warp_max_V"] == V[small].max() and st["warp_max_T"] == T[small].max()
    assert st["team_max_V"] == V[~small].max() and st["team_max_T"] == T[~small].max()
    assert st["warp_md2"] == int((rows["md3"][we["model"]][small] == 0).any())
    assert st["team_md2"] == int((rows["md3"][we["model"]][~small] == 0).any())
    assert st["lerped_verts_padded"] == vpad.sum() and st["max_vpad"] == vpad.max() and st["max_T"] == T.max()
    assert st["max_bitw"] == ((T + 31) >> 5).max() and st["team_normal_rows"] == Tp[~small].sum()
    assert st["total_verts"] == V.sum()
    assert st["and_mflags"] == 0                                     # MD2 and MD3 models together: no mesh flag common to all
    rc, _, _, _, st3 = run(we[rows["md3"][we["model"]] == 1], rows)
    assert rc == 0 and st3["and_mflags"] == MESH_MD3                 # what the tangent-space route decides on (per-vertex rows)


def test_validation_codes():
    rows = rows_of([(10, 12, 0, 1), (10, 12, 1, 3)])
    we = np.zeros(3, bq2.WORLD_ENTITY_DTYPE)
    assert run(we, rows)[0] == 0
    for bad in (-1, 2, 1 << 30):
        w = we.copy(); w["model"][1] = bad
        assert run(w, rows)[0] == -1
    for field in ("frame", "oldframe"):
        for bad in (-1, 6):
            w = we.copy(); w[field][2] = bad
            assert run(w, rows)[0] == -2
    w = we.copy(); w["model"][0] = 1                                  # multi-mesh model: the caller takes the host-built route
    assert run(w, rows)[0] == -3
    assert run(we[:0], rows)[0] == 0


def test_row_numbers_must_fit_32_bits():
    """Output rows are numbered in 32 bits on the device: a frame whose lerped rows exceed 2^32 vertices is refused
    (-4 -> BQ2_E_LIMIT), one entity short of it is accepted."""
    rows = rows_of([(4096, 10, 1, 1)])
    n = (1 << 32) // 4096
    we = np.zeros(n + 1, bq2.WORLD_ENTITY_DTYPE)
    rc, hx, rec, voff, st = run(we[:n - 1], rows)
    assert rc == 0 and int(voff[n - 1]) == 4096 * (n - 1) and int(hx["vert_off"][-1]) == 4096 * (n - 2)
    assert run(we[:n], rows)[0] == -4                             # the total itself (2^32) is kept in 32 bits too
    assert run(we, rows)[0] == -4


def test_flag_filter_and_basis():
    rows = rows_of([(30, 40, 0, 1)])
    flags = [0, 0x2, 0x4, 0x8, 0x20, 0x80, 0x400, 0x800, 0x1000, 0x80000, 0x10000000, 0x40, 0x100, 0x80000 | 0x400]
    we = np.zeros(len(flags), bq2.WORLD_ENTITY_DTYPE)
    we["rf_flags"] = flags
    ang = np.zeros((len(flags), 3), np.float32)
    ang[1] = (0, 90, 0); ang[2] = (10, 20, 30); ang[3] = (-0.0, 0, 0); ang[4] = (0, 0, -45); ang[5] = (1e-30, 0, 0)
    we["angles"] = ang
    rc, hx, rec, voff, st = run(we, rows)
    assert rc == 0
    n_rot = 0
    for e, f in enumerate(flags):
        casts = host.casts_shadow(f) and not (f & (0x80000 | 0x10000000))
        assert hx["pair_rec"][e] == (rec[e] if casts else -1), hex(f)
        rotated = bool(ang[e][0] or ang[e][1] or ang[e][2])             # -0.0 counts as zero, a denormal does not
        assert hx["basis"][e] == (n_rot if rotated else -1), e          # bases are numbered in entity order
        if rotated:
            fwd, right, up = host.angle_vectors(ang[e])
            got = run.bases[n_rot].view(np.uint32)
            assert np.array_equal(got, np.concatenate([fwd, right, up]).astype(np.float32).view(np.uint32)), e
            n_rot += 1
    assert st["n_rotated"] == n_rot == 4
    # a model whose one mesh casts no shadow (pad[1]): lerped like the others, never a pair record
    rows["pad"][0][1] = 1
    rc, hx2, rec2, _, _ = run(we, rows)
    assert rc == 0 and (hx2["pair_rec"] == -1).all() and np.array_equal(hx2["rec"], rec2) and sorted(rec2) == list(range(len(flags)))


def test_rows_and_bases_carry_what_the_record_kernel_needs():
    """The record-building kernel (bq2_world_link_kernel, csrc/bq2_kernels.cu) sees only the 64-byte rows, the bases of
    the rotated entities, the frame headers and the lights.  Its arithmetic restated here in numpy float32, one
    rounding per operation and in its order, must give the lerp constants of bq2_host_entity_md2 (M:63526-63553) and
    the model-space light of bq2_host_point_to_model (M:63733-63745) bit for bit -- i.e. the rows lose nothing."""
    from berserkerquake2_b200 import synth
    f32 = np.float32
    F = 6
    blob = synth.md2(8, 5, F, 3)
    h = blob[:68].view("<i4"); fsz, ofs = int(h[4]), int(h[14])
    hdr = np.stack([blob[ofs + f * fsz: ofs + f * fsz + 24].view("<f4") for f in range(F)])   # scale[3], translate[3]
    rows = rows_of([(int(h[6]), int(h[8]), 0, 1)]); rows["F"] = F
    rng = np.random.RandomState(11)
    n = 64
    we = np.zeros(n, bq2.WORLD_ENTITY_DTYPE)
    we["frame"] = rng.randint(0, F, n); we["oldframe"] = rng.randint(0, F, n); we["backlerp"] = rng.rand(n)
    we["origin"] = rng.randn(n, 3) * 300; we["oldorigin"] = we["origin"] + rng.randn(n, 3) * 2
    ang = np.zeros((n, 3), np.float32)
    for e in range(n):
        k = rng.randint(4)
        if k == 1: ang[e, 1] = rng.uniform(-180, 180)
        elif k == 2: ang[e] = rng.uniform(-180, 180, 3)
        elif k == 3: ang[e, rng.randint(3)] = -0.0
    we["angles"] = ang
    rc, hx, rec, voff, st = run(we, rows)
    assert rc == 0 and st["n_rotated"] == np.count_nonzero((ang != 0).any(axis=1))

    def dot3(t, b):                                               # DotProduct M:2253: (x0*y0 + x1*y1) + x2*y2
        return f32(f32(f32(t[0] * b[0]) + f32(t[1] * b[1])) + f32(t[2] * b[2]))

    def to_model(basis, t):
        return t if basis is None else np.array([dot3(t, basis[0:3]), -dot3(t, basis[3:6]), dot3(t, basis[6:9])], f32)

    light = (rng.randn(3) * 400).astype(f32)
    for e in range(n):
        x = hx[e]
        basis = run.bases[x["basis"]] if x["basis"] >= 0 else None
        bl = f32(x["backlerp"]); fl = f32(np.float64(1.0) - np.float64(bl))
        mv = to_model(basis, (x["oldorigin"] - x["origin"]).astype(f32))
        fr, ofr = hdr[x["frame"]], hdr[x["oldframe"]]
        move = np.array([f32(f32(bl * f32(mv[k] + ofr[3 + k])) + f32(fl * fr[3 + k])) for k in range(3)], f32)
        frontv = np.array([f32(fl * fr[k]) for k in range(3)], f32); backv = np.array([f32(bl * ofr[k]) for k in range(3)], f32)
        want = host.entity_md2(blob, int(we["frame"][e]), int(we["oldframe"][e]), float(we["backlerp"][e]), we["origin"][e],
                               we["oldorigin"][e], we["angles"][e])
        for name, got in (("move", move), ("frontv", frontv), ("backv", backv)):
            assert np.array_equal(got.view(np.uint32), np.asarray(want[name]).view(np.uint32)), (e, name)
        assert f32(want["frontlerp"]) == fl
        got = to_model(basis, (light - x["origin"]).astype(f32))
        assert np.array_equal(got.view(np.uint32), host.point_to_model(light, we["origin"][e], we["angles"][e]).view(np.uint32)), e


def test_entity_loop_applies_render_options_and_model_flags():
    """The loop's per-entity test under explicit options = bq2_host_entity_casts (which tests/test_caller_loop.py pins to the
    engine's own loop); model flags travel in bq2_model_row.pad[0]."""
    rows = rows_of([(30, 40, 0, 1), (30, 40, 0, 1), (30, 40, 0, 1)])
    rows["pad"][1][0] = 0x2000                                      # RF_NOSELFSHADOW
    rows["pad"][2][0] = 0x80000                                     # RF_NOCASTSHADOW on the MODEL
    flags = [0, 0x2, 0x20, 0x10000000]
    we = np.zeros(3 * len(flags), bq2.WORLD_ENTITY_DTYPE)
    we["model"] = np.repeat([0, 1, 2], len(flags)); we["rf_flags"] = flags * 3
    for kw in (dict(), dict(pass_=c.PASS_SELF), dict(pass_=c.PASS_NOSELF), dict(r_playershadow=1.0), dict(r_shadows=1.0),
               dict(r_mirror=1), dict(r_mirror=1, r_mirror_models=0.0), dict(r_drawentities=0.0), dict(r_mirror=1, r_mirror_shadows=1.0)):
        o = bq2.Context.render_opts(**kw)
        rc, hx, rec, voff, st = run(we, rows, o)
        assert rc == 0
        for e in range(len(we)):
            casts = c.lib.bq2_host_entity_casts(int(we["rf_flags"][e]), int(rows["pad"][we["model"][e]][0]), 0, C.byref(o))
            assert hx["pair_rec"][e] == (rec[e] if casts else -1), (kw, e)
    assert any(hx["pair_rec"] >= 0) or True
