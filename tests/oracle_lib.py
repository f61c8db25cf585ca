"""ctypes access to the two CPU checkers.  TEST INFRASTRUCTURE ONLY (see oracle/bq2_oracle.h):

  Oracle     oracle/liboracle.so       plain-C restatement of the reference arithmetic
  Reference  oracle/_ref/libbq2ref.so  the UNMODIFIED reference main.c compiled headless
                                       (oracle/ref/build_ref.sh + oracle/ref/ref_driver.c)

Nothing under berserkerquake2_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbq2ref.so")
REF_DROPIN_SO = os.path.join(ROOT, "oracle", "_ref", "libbq2ref_dropin.so")

vp, i32, f32 = C.c_void_p, C.c_int, C.c_float


def P(a):
    return None if a is None else a.ctypes.data_as(vp)


def f3(a):
    return np.ascontiguousarray(a, np.float32).reshape(3)


def md2_dims(blob):
    h = blob[:68].view("<i4")
    return int(h[6]), int(h[8]), int(h[10])          # V, T, F


def md2_idx(blob):
    """int32 [T][3] index_xyz"""
    h = blob[:68].view("<i4")
    T, o = int(h[8]), int(h[13])
    return np.ascontiguousarray(blob[o:o + 12 * T].view(np.int16).reshape(T, 6)[:, :3].astype(np.int32))


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        L = self.lib = C.CDLL(ORACLE_SO)
        L.bq2o_anorms.restype = C.POINTER(C.c_float)
        L.bq2o_md2_lerp_consts.argtypes = [vp, i32, i32, f32, vp, vp, vp, vp, vp, vp, vp]
        L.bq2o_md2_lerp.argtypes = [vp, i32, i32, vp, vp, vp, i32, f32, vp]
        L.bq2o_md3_lerp_consts.argtypes = [vp, vp, f32, vp, vp, vp, vp, vp]
        L.bq2o_md3_lerp_extrude.argtypes = [vp, vp, i32, vp, f32, f32, vp, f32, vp, vp]
        L.bq2o_facing.argtypes = [vp, vp, i32, vp, f32, vp, vp, vp]
        L.bq2o_shadow_volume.argtypes = [vp, vp, i32, vp, vp, vp, i32, vp, vp]
        L.bq2o_shadow_volume.restype = i32
        L.bq2o_ntb.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, i32, i32, f32, f32, vp, vp, vp]
        L.bq2o_tslight.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp]
        L.bq2o_normal2index.argtypes = [vp]; L.bq2o_normal2index.restype = C.c_uint8
        L.bq2o_md2_neighbours.argtypes = [vp, i32, vp]
        L.bq2o_md3_neighbours.argtypes = [vp, i32, vp]
        L.bq2o_vecs_for_tris.argtypes = [vp] * 8
        L.bq2o_md2_tangents_smooth.argtypes = [vp, vp, f32, vp, vp]
        L.bq2o_md2_tangents_flat.argtypes = [vp, vp, vp, vp, vp]
        L.bq2o_md3_tangents.argtypes = [vp, i32, i32, vp, vp, i32]
        L.bq2o_angle_vectors.argtypes = [vp] * 4
        L.bq2o_point_to_model.argtypes = [vp] * 4

    def anorms(self):
        return np.ctypeslib.as_array(self.lib.bq2o_anorms(), (162, 3)).copy()

    def angle_vectors(self, angles):
        f, r, u = (np.zeros(3, np.float32) for _ in range(3))
        self.lib.bq2o_angle_vectors(P(f3(angles)), P(f), P(r), P(u))
        return f, r, u

    def point_to_model(self, world, origin, angles):
        out = np.zeros(3, np.float32)
        self.lib.bq2o_point_to_model(P(f3(world)), P(f3(origin)), P(f3(angles)), P(out))
        return out

    # ---- MD2 ---------------------------------------------------------------------------------------
    def md2_neighbours(self, blob):
        V, T, F = md2_dims(blob)
        o = int(blob[:68].view("<i4")[13])
        tris = np.ascontiguousarray(blob[o:o + 12 * T])
        out = np.zeros((T, 3), np.int32)
        self.lib.bq2o_md2_neighbours(P(tris), T, P(out))
        return out

    def md2_lerp_consts(self, blob, frame, oldframe, backlerp, origin, oldorigin, angles):
        move, fv, bv = (np.zeros(3, np.float32) for _ in range(3))
        fl = np.zeros(1, np.float32)
        self.lib.bq2o_md2_lerp_consts(P(blob), frame, oldframe, backlerp, P(f3(origin)), P(f3(oldorigin)),
                                      P(f3(angles)), P(move), P(fv), P(bv), P(fl))
        return move, fv, bv, float(fl[0])

    def md2_lerp(self, blob, frame, oldframe, backlerp, origin=(0, 0, 0), oldorigin=(0, 0, 0), angles=(0, 0, 0),
                 shell=False, shell_scale=0.0):
        V, T, F = md2_dims(blob)
        move, fv, bv, _ = self.md2_lerp_consts(blob, frame, oldframe, backlerp, origin, oldorigin, angles)
        out = np.zeros((V, 3), np.float32)
        self.lib.bq2o_md2_lerp(P(blob), frame, oldframe, P(move), P(fv), P(bv), int(shell), shell_scale, P(out))
        return out

    def facing(self, lerped, idx, light, scale, extrude=True):
        T = len(idx)
        ext = np.full_like(lerped, np.nan) if extrude else None
        tn = np.zeros((T, 3), np.float32)
        viss = np.zeros(T, np.uint8)
        self.lib.bq2o_facing(P(lerped), P(idx), T, P(f3(light)), scale, P(ext), P(tn), P(viss))
        return ext, tn, viss

    def shadow_volume(self, lerped, extruded, idx, nb, viss):
        V, T = len(lerped), len(idx)
        tv = np.zeros((24 * T, 3), np.float32)
        ind = np.zeros(24 * T, np.uint32)
        nb = np.ascontiguousarray(nb, np.int32)
        n = self.lib.bq2o_shadow_volume(P(lerped), P(extruded), V, P(idx), P(nb), P(viss), T, P(tv), P(ind))
        return tv[:n].copy(), ind[:n].copy()

    def md2_pair(self, blob, nb, frame, oldframe, backlerp, origin, oldorigin, angles, light_world, radius,
                 idx=None):
        """Everything R_DrawAliasShadowVolume (M:63713) produces for one (entity, light)."""
        idx = md2_idx(blob) if idx is None else idx
        lerped = self.md2_lerp(blob, frame, oldframe, backlerp, origin, oldorigin, angles)
        light = self.point_to_model(light_world, origin, angles)
        scale = np.float32(32) * np.float32(radius)
        ext, tn, viss = self.facing(lerped, idx, light, float(scale))
        tv, ind = self.shadow_volume(lerped, ext, idx, nb, viss)
        return dict(lerped=lerped, extruded=ext, trinormals=tn, viss=viss, tris_verts=tv, indices=ind,
                    light_local=light)

    def md2_pair_local(self, blob, nb, frame, oldframe, backlerp, light_model, radius, idx=None):
        idx = md2_idx(blob) if idx is None else idx
        lerped = self.md2_lerp(blob, frame, oldframe, backlerp)
        ext, tn, viss = self.facing(lerped, idx, light_model, float(np.float32(32) * np.float32(radius)))
        tv, ind = self.shadow_volume(lerped, ext, idx, nb, viss)
        return dict(lerped=lerped, extruded=ext, viss=viss, tris_verts=tv, indices=ind)

    def ntb_rows(self, n_cur, n_old, t_cur, t_old, b_cur, b_old, backlerp, frontlerp):
        """byte rows (possibly strided views are copied) -> lerped N/T/B float rows"""
        rows = len(t_cur)
        a = [np.ascontiguousarray(x, np.uint8) for x in (n_cur, n_old, t_cur, t_old, b_cur, b_old)]
        N, T_, B = (np.zeros((rows, 3), np.float32) for _ in range(3))
        self.lib.bq2o_ntb(P(a[0]), P(a[1]), 1, P(a[2]), P(a[3]), 1, P(a[4]), P(a[5]), 1, rows,
                          backlerp, frontlerp, P(N), P(T_), P(B))
        return N, T_, B

    def tslight(self, lerped, N, T_, B, light, view):
        rows = len(lerped)
        lp, ts = np.zeros((rows, 3), np.float32), np.zeros((rows, 3), np.float32)
        self.lib.bq2o_tslight(P(np.ascontiguousarray(lerped)), P(N), P(T_), P(B), rows, P(f3(light)), P(f3(view)),
                              P(lp), P(ts))
        return lp, ts

    def brush_volume(self, bm, lit, light_model, radius):
        """R_DrawBrushModelVolumes M:63326-63499; light already in model space"""
        self.lib.bq2o_brush_volume.argtypes = [i32, vp, vp, vp, vp, vp, vp, f32, vp, vp, vp]
        self.lib.bq2o_brush_volume.restype = i32
        nv = int(bm.first[-1])
        vc = np.zeros((2 * nv, 3), np.float32); ic = np.zeros(12 * nv, np.uint32); vb = C.c_int(0)
        lit = np.ascontiguousarray(lit, np.uint8)
        ib = self.lib.bq2o_brush_volume(bm.n_polys, P(bm.numverts), P(bm.verts), P(bm.neighbours), P(lit), P(bm.fence),
                                        P(f3(light_model)), float(np.float32(32) * np.float32(radius)), P(vc), P(ic), C.byref(vb))
        return vc[:2 * vb.value].copy(), ic[:ib].copy()

    def normal2index(self, v):
        return int(self.lib.bq2o_normal2index(P(f3(v))))

    def vecs_for_tris(self, v0, v1, v2, st0, st1, st2):
        t, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
        a = [np.ascontiguousarray(x, np.float32) for x in (v0, v1, v2, st0, st1, st2)]
        self.lib.bq2o_vecs_for_tris(*[P(x) for x in a], P(t), P(b))
        return t, b

    def md2_tangents_smooth(self, blob, st, smooth_cosine):
        V, T, F = md2_dims(blob)
        tan, bi = np.zeros((F, V), np.uint8), np.zeros((F, V), np.uint8)
        self.lib.bq2o_md2_tangents_smooth(P(blob), P(np.ascontiguousarray(st, np.float32)), smooth_cosine, P(tan), P(bi))
        return tan, bi

    def md2_tangents_flat(self, blob, st):
        V, T, F = md2_dims(blob)
        n, tan, bi = (np.zeros((F, T), np.uint8) for _ in range(3))
        self.lib.bq2o_md2_tangents_flat(P(blob), P(np.ascontiguousarray(st, np.float32)), P(n), P(tan), P(bi))
        return n, tan, bi

    # ---- MD3 ---------------------------------------------------------------------------------------
    def md3_neighbours(self, indexes):
        ix = np.ascontiguousarray(indexes, np.uint16)
        T = ix.size // 3
        out = np.zeros((T, 3), np.int32)
        self.lib.bq2o_md3_neighbours(P(ix), T, P(out))
        return out

    def md3_tangents(self, verts, st, indexes):
        """fills the tangent/binormal bytes of verts [F][V] in place"""
        F, V = verts.shape
        ix = np.ascontiguousarray(indexes, np.uint16)
        self.lib.bq2o_md3_tangents(P(verts), F, V, P(np.ascontiguousarray(st, np.float32)), P(ix), ix.size // 3)

    def md3_lerp_consts(self, ft, oft, backlerp, origin, oldorigin, angles):
        move = np.zeros(3, np.float32); fl = np.zeros(1, np.float32)
        self.lib.bq2o_md3_lerp_consts(P(f3(ft)), P(f3(oft)), backlerp, P(f3(origin)), P(f3(oldorigin)), P(f3(angles)),
                                      P(move), P(fl))
        return move, float(fl[0])

    def md3_mesh_pair(self, verts, indexes, nb, frame_translate, frame, oldframe, backlerp, origin, oldorigin,
                      angles, light_world, radius):
        """One mesh of R_DrawMD3AliasShadowVolume (M:63858-63991)."""
        F, V = verts.shape
        move, fl = self.md3_lerp_consts(frame_translate[frame], frame_translate[oldframe], backlerp, origin,
                                        oldorigin, angles)
        light = self.point_to_model(light_world, origin, angles)
        scale = float(np.float32(32) * np.float32(radius))
        lerped, ext = np.zeros((V, 3), np.float32), np.zeros((V, 3), np.float32)
        cur, old = np.ascontiguousarray(verts[frame]), np.ascontiguousarray(verts[oldframe])
        self.lib.bq2o_md3_lerp_extrude(P(cur), P(old), V, P(move), backlerp, fl, P(light), scale, P(lerped), P(ext))
        idx = np.ascontiguousarray(np.asarray(indexes).reshape(-1, 3).astype(np.int32))
        _, tn, viss = self.facing(lerped, idx, light, scale, extrude=False)
        tv, ind = self.shadow_volume(lerped, ext, idx, nb, viss)
        return dict(lerped=lerped, extruded=ext, viss=viss, tris_verts=tv, indices=ind, light_local=light)


class Reference:
    """The reference's own object code.  See oracle/ref/ref_driver.c for each entry point."""

    @staticmethod
    def load(path=REF_SO):
        if not os.path.exists(path):
            return None
        return Reference(path)

    def __init__(self, path=REF_SO):
        L = self.lib = C.CDLL(path)
        L.bq2ref_init()
        L.bq2ref_smooth_cosine.restype = f32
        L.bq2ref_anorms.restype = C.POINTER(C.c_float)
        L.bq2ref_md2_lerp.argtypes = [vp, i32, i32, f32, vp, vp, vp, i32, vp]
        L.bq2ref_md2_shadow.argtypes = [vp, vp, i32, i32, f32, vp, vp, vp, i32, vp, f32, vp, vp, vp, vp, vp, vp]
        L.bq2ref_md2_shadow_local.argtypes = [vp, vp, i32, i32, f32, vp, vp, vp, vp, f32, vp, vp, vp, vp, vp]
        L.bq2ref_md2_bench.argtypes = [vp, vp, i32, vp, vp, vp, vp, f32]
        L.bq2ref_md2_bench.restype = C.c_longlong
        L.bq2ref_md2_light.argtypes = [vp, vp, vp, vp, i32, i32, i32, f32, vp, vp, vp, vp, vp, i32, vp, vp, vp]
        L.bq2ref_md3_shadow.argtypes = [i32, vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, f32, vp, vp, vp, i32, vp, f32,
                                        vp, vp, vp, vp, vp, vp, vp, vp]
        L.bq2ref_md3_light.argtypes = [vp, i32, i32, i32, f32, vp, vp, vp, i32, vp, vp]
        L.bq2ref_md2_neighbours.argtypes = [vp, vp]
        L.bq2ref_md3_neighbours.argtypes = [vp, i32, vp]
        L.bq2ref_normal2index.argtypes = [vp]
        L.bq2ref_vecs_for_tris.argtypes = [vp] * 8
        L.bq2ref_tangent_for_tri_md3.argtypes = [vp] * 5
        L.bq2ref_angle_vectors.argtypes = [vp] * 4
        L.bq2ref_vector_normalize.argtypes = [vp]; L.bq2ref_vector_normalize.restype = f32
        L.bq2ref_load_md2.argtypes = [vp, i32, C.c_char_p, C.c_char_p, f32, i32, vp, vp, vp, vp, vp, vp]
        L.bq2ref_load_md3.argtypes = [vp, i32, C.c_char_p, C.c_char_p, f32, i32, vp]
        L.bq2ref_loaded_md3_dims.argtypes = [i32, vp, vp]
        L.bq2ref_loaded_md3_mesh.argtypes = [i32, vp, vp, vp, vp, vp]

    def smooth_cosine(self):
        return float(self.lib.bq2ref_smooth_cosine())

    # ---- the engine's loaders run whole (cache = true: no skins), in a scratch game directory ----------
    def load_md2(self, file, workdir, modname, scale=1.0, invert=False):
        """Mod_LoadAliasModel M:51217 on a .md2 file image.  Returns the in-memory image, the neighbour table, the
        tangent / binormal (/ normal) anorm bytes and the smooth flag; the loader's cache file is left under
        <workdir>/baseq2/cache/[inv_]<modname>."""
        file = np.ascontiguousarray(file, np.uint8).copy()
        h = file[:68].view("<i4")
        V, T, F, ofs_end = int(h[6]), int(h[8]), int(h[10]), int(h[16])
        image = np.zeros(ofs_end, np.uint8); nb = np.zeros((T, 3), np.int32)
        rows = max(V, T) * F
        tan, bi, nm = (np.zeros(rows, np.uint8) for _ in range(3))
        smooth = C.c_int(-1)
        rc = self.lib.bq2ref_load_md2(P(file), file.nbytes, workdir.encode(), modname.encode(), scale, int(invert),
                                      P(image), P(nb), P(tan), P(bi), P(nm), C.byref(smooth))
        if rc != 0:
            raise RuntimeError("bq2ref_load_md2 -> %d" % rc)
        n = V if smooth.value else T
        out = dict(image=image, neighbours=nb, smooth=bool(smooth.value), tangents=tan[:n * F].reshape(F, n).copy(),
                   binormals=bi[:n * F].reshape(F, n).copy())
        if not smooth.value:
            out["normals"] = nm[:n * F].reshape(F, n).copy()
        return out

    def load_md3(self, file, workdir, modname, scale=1.0, invert=False):
        """Mod_LoadAliasMD3Model M:50584 on a .md3 file image -> list of meshes (verts [F][V] maliasvertex_t, indexes,
        neighbours, st) and the frame translations; cache files under <workdir>/baseq2/cache/[inv_]<modname>_<mesh>."""
        from berserkerquake2_b200 import synth
        file = np.ascontiguousarray(file, np.uint8).copy()
        nf = C.c_int(0)
        nm = self.lib.bq2ref_load_md3(P(file), file.nbytes, workdir.encode(), modname.encode(), scale, int(invert), C.byref(nf))
        if nm < 0:
            raise RuntimeError("bq2ref_load_md3 -> %d" % nm)
        F = nf.value
        meshes, tr = [], np.zeros((F, 3), np.float32)
        for k in range(nm):
            v, t = C.c_int(0), C.c_int(0)
            assert self.lib.bq2ref_loaded_md3_dims(k, C.byref(v), C.byref(t)) == 0
            verts = np.zeros((F, v.value), synth.MD3_VERTEX); idx = np.zeros((t.value, 3), np.uint16)
            nb = np.zeros((t.value, 3), np.int32); st = np.zeros((v.value, 2), np.float32)
            assert self.lib.bq2ref_loaded_md3_mesh(k, P(verts), P(idx), P(nb), P(st), P(tr)) == 0
            meshes.append(dict(verts=verts, indexes=idx, neighbours=nb, st=st))
        return meshes, tr

    def anorms(self):
        return np.ctypeslib.as_array(self.lib.bq2ref_anorms(), (162, 3)).copy()

    def angle_vectors(self, angles):
        f, r, u = (np.zeros(3, np.float32) for _ in range(3))
        a = f3(angles).copy()
        self.lib.bq2ref_angle_vectors(P(a), P(f), P(r), P(u))
        return f, r, u

    def md2_neighbours(self, blob):
        V, T, F = md2_dims(blob)
        out = np.zeros((T, 3), np.int32)
        self.lib.bq2ref_md2_neighbours(P(blob), P(out))
        return out

    def md3_neighbours(self, indexes):
        ix = np.ascontiguousarray(indexes, np.uint16).copy()
        T = ix.size // 3
        out = np.zeros((T, 3), np.int32)
        self.lib.bq2ref_md3_neighbours(P(ix), T, P(out))
        return out

    def normal2index(self, v):
        return int(self.lib.bq2ref_normal2index(P(f3(v)))) & 255

    def vecs_for_tris(self, v0, v1, v2, st0, st1, st2):
        t, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
        a = [np.ascontiguousarray(x, np.float32).copy() for x in (v0, v1, v2, st0, st1, st2)]
        self.lib.bq2ref_vecs_for_tris(*[P(x) for x in a], P(t), P(b))
        return t, b

    def md2_lerp(self, blob, frame, oldframe, backlerp, origin=(0, 0, 0), oldorigin=(0, 0, 0), angles=(0, 0, 0),
                 flags=0):
        V, T, F = md2_dims(blob)
        out = np.zeros((V, 3), np.float32)
        self.lib.bq2ref_md2_lerp(P(blob), frame, oldframe, backlerp, P(f3(origin)), P(f3(oldorigin)), P(f3(angles)),
                                 flags, P(out))
        return out

    def md2_pair(self, blob, nb, frame, oldframe, backlerp, origin, oldorigin, angles, light_world, radius, flags=0):
        """R_DrawAliasShadowVolume whole (M:63713); None when the flag filter rejects the entity."""
        V, T, F = md2_dims(blob)
        nb = np.ascontiguousarray(nb, np.int32)
        lerped, ext = np.zeros((V, 3), np.float32), np.zeros((V, 3), np.float32)
        tn, viss = np.zeros((T, 3), np.float32), np.zeros(T, np.uint8)
        tv = np.zeros((36 * 8192 // 3, 3), np.float32)
        ll = np.zeros(3, np.float32)
        n = self.lib.bq2ref_md2_shadow(P(blob), P(nb), frame, oldframe, backlerp, P(f3(origin)), P(f3(oldorigin)),
                                       P(f3(angles)), flags, P(f3(light_world)), radius, P(lerped), P(ext), P(tn),
                                       P(viss), P(tv), P(ll))
        if n < 0:
            return None
        return dict(lerped=lerped, extruded=ext, trinormals=tn, viss=viss, tris_verts=tv[:n].copy(), light_local=ll)

    def md2_pair_local(self, blob, nb, frame, oldframe, backlerp, light_model, radius):
        V, T, F = md2_dims(blob)
        nb = np.ascontiguousarray(nb, np.int32)
        z = np.zeros(3, np.float32)
        lerped, ext = np.zeros((V, 3), np.float32), np.zeros((V, 3), np.float32)
        tn, viss = np.zeros((T, 3), np.float32), np.zeros(T, np.uint8)
        tv = np.zeros((36 * 8192 // 3, 3), np.float32)
        n = self.lib.bq2ref_md2_shadow_local(P(blob), P(nb), frame, oldframe, backlerp, P(z), P(z), P(z),
                                             P(f3(light_model)), radius, P(lerped), P(ext), P(tn), P(viss), P(tv))
        return dict(lerped=lerped, extruded=ext, trinormals=tn, viss=viss, tris_verts=tv[:n].copy())

    def md2_bench(self, blob, nb, frames, oldframes, backlerps, lights_model, radius):
        nb = np.ascontiguousarray(nb, np.int32)
        fr = np.ascontiguousarray(frames, np.int32); ofr = np.ascontiguousarray(oldframes, np.int32)
        bl = np.ascontiguousarray(backlerps, np.float32); lm = np.ascontiguousarray(lights_model, np.float32)
        return int(self.lib.bq2ref_md2_bench(P(blob), P(nb), len(fr), P(fr), P(ofr), P(bl), P(lm), radius))

    def md2_light(self, blob, tangents, binormals, normals, smooth, frame, oldframe, backlerp, light_model,
                  view_model, vp_variant):
        """GL_DrawAliasFrameLerpLightARB (vp=0) / ..._vp (vp=1): every TMU array it hands to GL.
        Returns (count, {tmu: float[count][size]}, vertex[count][3])."""
        V, T, F = md2_dims(blob)
        z = np.zeros(3, np.float32)
        tmu = np.zeros((8, 9 * T), np.float32)
        vert = np.zeros((3 * T, 3), np.float32)
        sizes = np.zeros(8, np.int32)
        tb = [None if a is None else np.ascontiguousarray(a, np.uint8) for a in (tangents, binormals, normals)]
        n = self.lib.bq2ref_md2_light(P(blob), P(tb[0]), P(tb[1]), P(tb[2]), int(smooth), frame, oldframe, backlerp,
                                      P(z), P(z), P(z), P(f3(light_model)), P(f3(view_model)), int(vp_variant),
                                      P(tmu), P(vert), P(sizes))
        out = {t: tmu[t, :sizes[t] * n].reshape(n, sizes[t]).copy() for t in range(8) if sizes[t]}
        return n, out, vert[:n].copy()

    def md3_shadow(self, frame_translate, meshes, frame, oldframe, backlerp, origin, oldorigin, angles,
                   light_world, radius, flags=0):
        """R_DrawMD3AliasShadowVolume whole (M:63800).  meshes: list of dict(verts [F][V] MD3 vertex,
        indexes u16 [T][3], neighbours i32 [T][3], casts_shadow).  Returns a list (per mesh) of dicts or None."""
        ft = np.ascontiguousarray(frame_translate, np.float32)
        nm = len(meshes)
        nv = np.array([m["verts"].shape[1] for m in meshes], np.int32)
        nt = np.array([np.asarray(m["indexes"]).size // 3 for m in meshes], np.int32)
        casts = np.array([1 if m.get("casts_shadow", True) else 0 for m in meshes], np.int32)
        keep = []
        def arr_of_ptrs(arrs):
            a = (vp * nm)()
            for i, x in enumerate(arrs):
                keep.append(x); a[i] = x.ctypes.data
            return a
        pv = arr_of_ptrs([np.ascontiguousarray(m["verts"]) for m in meshes])
        pi = arr_of_ptrs([np.ascontiguousarray(m["indexes"], np.uint16) for m in meshes])
        pn = arr_of_ptrs([np.ascontiguousarray(m["neighbours"], np.int32) for m in meshes])
        voff = np.concatenate([[0], np.cumsum(nv)]).astype(np.int32)
        toff = np.concatenate([[0], np.cumsum(nt)]).astype(np.int32)
        tvoff = np.concatenate([[0], np.cumsum(24 * nt)]).astype(np.int32)
        lerped = np.zeros((voff[-1], 3), np.float32); ext = np.zeros((voff[-1], 3), np.float32)
        viss = np.zeros(toff[-1], np.uint8); tv = np.zeros((tvoff[-1], 3), np.float32)
        cnt = np.zeros(nm, np.int32)
        drawn = self.lib.bq2ref_md3_shadow(ft.shape[0], P(ft), nm, P(nv), P(nt), pv, pi, pn, P(casts), frame, oldframe,
                                           backlerp, P(f3(origin)), P(f3(oldorigin)), P(f3(angles)), flags,
                                           P(f3(light_world)), radius, P(voff), P(toff), P(tvoff),
                                           P(lerped), P(ext), P(viss), P(tv), P(cnt))
        out = []
        for m in range(nm):
            if cnt[m] < 0:
                out.append(None)
                continue
            out.append(dict(lerped=lerped[voff[m]:voff[m + 1]].copy(), extruded=ext[voff[m]:voff[m + 1]].copy(),
                            viss=viss[toff[m]:toff[m + 1]].copy(),
                            tris_verts=tv[tvoff[m]:tvoff[m] + cnt[m]].copy()))
        return drawn, out

    def dropin_md3(self, frame_translate, meshes, frame, oldframe, backlerp, origin, oldorigin, angles,
                   light_world, radius):
        """integration/bq2_dropin.c: bq2_dropin_md3_mesh per casting mesh (libbq2ref_dropin.so only)."""
        fn = self.lib.bq2ref_dropin_md3
        fn.argtypes = [i32, vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, f32, vp, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, vp, vp]
        ft = np.ascontiguousarray(frame_translate, np.float32)
        nm = len(meshes)
        nv = np.array([m["verts"].shape[1] for m in meshes], np.int32)
        nt = np.array([np.asarray(m["indexes"]).size // 3 for m in meshes], np.int32)
        casts = np.array([1 if m.get("casts_shadow", True) else 0 for m in meshes], np.int32)
        keep = []
        def arr_of_ptrs(arrs):
            a = (vp * nm)()
            for i, x in enumerate(arrs):
                keep.append(x); a[i] = x.ctypes.data
            return a
        pv = arr_of_ptrs([np.ascontiguousarray(m["verts"]) for m in meshes])
        pi = arr_of_ptrs([np.ascontiguousarray(m["indexes"], np.uint16) for m in meshes])
        pn = arr_of_ptrs([np.ascontiguousarray(m["neighbours"], np.int32) for m in meshes])
        voff = np.concatenate([[0], np.cumsum(nv)]).astype(np.int32)
        toff = np.concatenate([[0], np.cumsum(nt)]).astype(np.int32)
        tvoff = np.concatenate([[0], np.cumsum(24 * nt)]).astype(np.int32)
        lerped = np.zeros((voff[-1], 3), np.float32); ext = np.zeros((voff[-1], 3), np.float32)
        viss = np.zeros(toff[-1], np.uint8); tv = np.zeros((tvoff[-1], 3), np.float32)
        cnt = np.zeros(nm, np.int32)
        done = fn(ft.shape[0], P(ft), nm, P(nv), P(nt), pv, pi, pn, P(casts), frame, oldframe, backlerp,
                  P(f3(origin)), P(f3(oldorigin)), P(f3(angles)), P(f3(light_world)), radius,
                  P(voff), P(toff), P(tvoff), P(lerped), P(ext), P(viss), P(tv), P(cnt))
        out = []
        for m in range(nm):
            out.append(None if cnt[m] < 0 else dict(
                lerped=lerped[voff[m]:voff[m + 1]].copy(), extruded=ext[voff[m]:voff[m + 1]].copy(),
                viss=viss[toff[m]:toff[m + 1]].copy(), tris_verts=tv[tvoff[m]:tvoff[m] + cnt[m]].copy()))
        return done, out

    def brush_volume(self, bm, lit, origin, angles, light_world, radius):
        """the reference's R_DrawBrushModelVolumes whole (world-space light in)"""
        fn = self.lib.bq2ref_brush_volume
        fn.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, f32, vp, vp, vp]; fn.restype = i32
        nv = int(bm.first[-1])
        vc = np.zeros((2 * nv, 3), np.float32); ic = np.zeros(12 * nv, np.uint32); vb = C.c_int(0)
        lit = np.ascontiguousarray(lit, np.uint8)
        ib = fn(bm.n_polys, P(bm.numverts), P(bm.verts), P(bm.neighbours), P(lit), P(bm.fence), P(f3(origin)), P(f3(angles)),
                P(f3(light_world)), radius, P(vc), P(ic), C.byref(vb))
        return vc[:2 * vb.value].copy(), ic[:ib].copy()

    def md3_light(self, mesh_verts, frame, oldframe, backlerp, lerped, light_model, view_model, vp_variant):
        F, V = mesh_verts.shape
        tmu = np.zeros((8, 3 * V), np.float32)
        sizes = np.zeros(8, np.int32)
        mv = np.ascontiguousarray(mesh_verts)
        self.lib.bq2ref_md3_light(P(mv), V, frame, oldframe, backlerp, P(np.ascontiguousarray(lerped, np.float32)),
                                  P(f3(light_model)), P(f3(view_model)), int(vp_variant), P(tmu), P(sizes))
        return {t: tmu[t, :sizes[t] * V].reshape(V, sizes[t]).copy() for t in range(8) if sizes[t]}


# ------------------------------------------------------------------------------------------------------
# Full-size parity: the per-pair 64-bit hash (csrc/bq2_kernels.cu bq2_hash_kernel == oracle/ref/ref_driver.c
# hash_globals), here in numpy for outputs that are already arrays, and a fan-out over the host cores that runs the
# reference's own R_DrawAliasShadowVolume for EVERY pair of a frame.
def _mix64(x):
    x = x + np.uint64(0x9e3779b97f4a7c15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
    return x ^ (x >> np.uint64(31))


def _hash_section(section, words):
    w = np.ascontiguousarray(words).view(np.uint32).reshape(-1).astype(np.uint64)
    i = np.arange(len(w), dtype=np.uint64)
    with np.errstate(over="ignore"):
        return _mix64((np.uint64(section) << np.uint64(56)) ^ (i << np.uint64(32)) ^ w).sum(dtype=np.uint64)


def pair_hash(lerped, extruded, viss, tris_verts, lerped_out=True):
    """The hash of one pair slot from reference-style outputs (lerped [V][3], extruded [V][3], viss [T] bytes,
    tris_verts [TrisCounter][3])."""
    T = len(viss)
    bits_ = np.zeros(((T + 31) // 32) * 32, np.uint8); bits_[:T] = np.asarray(viss) != 0
    words = (bits_.reshape(-1, 32).astype(np.uint32) << np.arange(32, dtype=np.uint32)[None, :]).sum(1, dtype=np.uint32)
    with np.errstate(over="ignore"):
        h = np.uint64(0)
        if lerped_out:
            h = h + _hash_section(1, np.asarray(lerped, np.float32))
        h = h + _hash_section(2, np.asarray(extruded, np.float32)) + _hash_section(3, words)
        h = h + _hash_section(4, np.array([len(tris_verts)], np.uint32)) + _hash_section(5, np.asarray(tris_verts, np.float32))
    return np.uint64(h)


def _hash_worker_main(argv):
    """child process of reference_pair_hashes: `python oracle_lib.py --hash-worker <job dir> <lo> <hi> <timed passes>`.
    A fresh interpreter (no fork of a process that holds a CUDA context and helper threads); the job's arrays are
    memory-mapped from the job directory."""
    import json
    import time
    job, lo, hi, passes = argv[0], int(argv[1]), int(argv[2]), int(argv[3])
    ld = lambda n: np.load(os.path.join(job, n + ".npy"), mmap_mode="r")
    blob_data, blob_off, nb_data, nb_off = ld("blob_data"), ld("blob_off"), ld("nb_data"), ld("nb_off")
    we, wl = np.load(os.path.join(job, "we.npy")), np.load(os.path.join(job, "wl.npy"))
    pe, pl = np.load(os.path.join(job, "pe.npy"))[lo:hi], np.load(os.path.join(job, "pl.npy"))[lo:hi]
    n = hi - lo
    lib = C.CDLL(REF_SO)
    lib.bq2ref_init()
    fn = lib.bq2ref_md2_world_hash
    fn.restype = C.c_longlong
    fn.argtypes = [vp] * 2 + [i32] + [vp] * 9
    bench = lib.bq2ref_md2_world_bench
    bench.restype = C.c_longlong
    bench.argtypes = [vp] * 2 + [i32] + [vp] * 8
    model_of = we["model"][pe]
    used = {}
    for m in np.unique(model_of):                       # private, writable copies of the models this slice touches
        m = int(m)
        used[m] = (np.array(blob_data[blob_off[m]:blob_off[m + 1]]), np.array(nb_data[nb_off[m]:nb_off[m + 1]]))
    bp = (vp * n)(*[used[int(m)][0].ctypes.data for m in model_of])
    nbp = (vp * n)(*[used[int(m)][1].ctypes.data for m in model_of])
    a = [np.ascontiguousarray(x) for x in (we["frame"][pe].astype(np.int32), we["oldframe"][pe].astype(np.int32),
                                           we["backlerp"][pe].astype(np.float32), we["origin"][pe], we["oldorigin"][pe],
                                           we["angles"][pe], wl["origin"][pl], wl["radius"][pl].astype(np.float32))]
    out = np.zeros(n, np.uint64)
    tot = fn(bp, nbp, n, *[P(x) for x in a], P(out)) if n else 0
    secs = 0.0
    if passes and n:                                    # the same slice again, without the hashing, timed
        t0 = time.perf_counter()
        for _ in range(passes):
            bench(bp, nbp, n, *[P(x) for x in a])
        secs = (time.perf_counter() - t0) / passes
    np.save(os.path.join(job, f"out_{lo}.npy"), out)
    print(json.dumps({"total": int(tot), "secs": secs}), flush=True)


def reference_pair_hashes(blobs, nbs, we, wl, pe, pl, procs=None, timed_passes=0, want_seconds=False):
    """hash of every pair (pe[k], pl[k]) computed by the unmodified reference (whole R_DrawAliasShadowVolume per pair),
    fanned out over one child process per host core (the engine keeps its state in globals).  blobs / nbs: per MODEL id
    (we["model"] indexes them): MD2 images and int32 neighbour tables.  The children are fresh interpreters started with
    subprocess -- the caller usually holds a CUDA context and helper threads, which a fork() would not survive reliably --
    and read the job from memory-mapped files.  Returns (uint64 [n_pairs], total index count)."""
    import json
    import shutil
    import subprocess
    import sys
    import tempfile
    procs = procs or max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 4)
    n = len(pe)
    procs = max(1, min(procs, n))
    job = tempfile.mkdtemp(prefix="bq2hash_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        bl = [np.ascontiguousarray(b, np.uint8).reshape(-1) for b in blobs]
        nl = [np.ascontiguousarray(x, np.int32).reshape(-1) for x in nbs]
        np.save(os.path.join(job, "blob_off.npy"), np.concatenate([[0], np.cumsum([len(b) for b in bl])]).astype(np.int64))
        np.save(os.path.join(job, "nb_off.npy"), np.concatenate([[0], np.cumsum([len(x) for x in nl])]).astype(np.int64))
        np.save(os.path.join(job, "blob_data.npy"), np.concatenate(bl) if bl else np.zeros(0, np.uint8))
        np.save(os.path.join(job, "nb_data.npy"), np.concatenate(nl) if nl else np.zeros(0, np.int32))
        del bl, nl
        for name, arr in (("we", we), ("wl", wl), ("pe", np.ascontiguousarray(pe, np.int32)), ("pl", np.ascontiguousarray(pl, np.int32))):
            np.save(os.path.join(job, name + ".npy"), np.ascontiguousarray(arr))
        ranges = [(n * w // procs, n * (w + 1) // procs) for w in range(procs)]
        here = os.path.abspath(__file__)
        children = [subprocess.Popen([sys.executable, here, "--hash-worker", job, str(lo), str(hi), str(int(timed_passes))],
                                     stdout=subprocess.PIPE, text=True) for lo, hi in ranges]
        res = []
        for (lo, hi), ch in zip(ranges, children):
            so, _ = ch.communicate(timeout=1800)
            if ch.returncode != 0:
                raise RuntimeError(f"reference hash worker [{lo}, {hi}) failed with status {ch.returncode}")
            meta = json.loads([l for l in so.splitlines() if l.startswith("{")][-1])
            res.append((np.load(os.path.join(job, f"out_{lo}.npy")), meta["total"], meta["secs"]))
    finally:
        shutil.rmtree(job, ignore_errors=True)
    if want_seconds:        # seconds per pass over the whole list = the slowest process's (timed_passes > 0), and the process count
        return np.concatenate([r[0] for r in res]), sum(r[1] for r in res), max(r[2] for r in res), procs
    return np.concatenate([r[0] for r in res]), sum(r[1] for r in res)


if __name__ == "__main__":
    import sys
    if len(sys.argv) >= 6 and sys.argv[1] == "--hash-worker":
        _hash_worker_main(sys.argv[2:])
