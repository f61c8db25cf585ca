"""Procedural MD2 / MD3 inputs (bq2_synth.c).  Pure data generation."""
import numpy as np
from . import _capi as c

MD2_HEADER = np.dtype([(n, "<i4") for n in (
    "ident", "version", "skinwidth", "skinheight", "framesize", "num_skins", "num_xyz", "num_st",
    "num_tris", "num_glcmds", "num_frames", "ofs_skins", "ofs_st", "ofs_tris", "ofs_frames",
    "ofs_glcmds", "ofs_end")])                       # dmdl_t, defs.h:3682-3706
MD3_VERTEX = np.dtype([("point", "<f4", 3), ("normal", "u1"), ("tangent", "u1"), ("binormal", "u1"),
                       ("pad", "u1")])               # maliasvertex_t, defs.h:4401-4407
assert MD3_VERTEX.itemsize == 16


def sphere_dims(slices, rings):
    return c.lib.bq2_synth_sphere_verts(slices, rings), c.lib.bq2_synth_sphere_tris(slices, rings)


def md2(slices=16, rings=17, num_frames=198, phase_seed=0):
    """In-memory MD2 image (uint8 array) of a closed blob: V = S*(R-1)+2, T = 2*S*(R-1)."""
    blob = np.zeros(c.lib.bq2_synth_md2_bytes(slices, rings, num_frames), np.uint8)
    if c.lib.bq2_synth_md2(slices, rings, num_frames, phase_seed, c.ptr(blob)) != 0:
        raise ValueError("bq2_synth_md2: bad arguments")
    return blob


def md2_header(blob):
    return blob[:MD2_HEADER.itemsize].view(MD2_HEADER)[0]


def md2_tris(blob):
    """int16 [T][6]: index_xyz[3], index_st[3] (dtriangle_t)"""
    h = md2_header(blob)
    o = int(h["ofs_tris"])
    return blob[o:o + 12 * int(h["num_tris"])].view(np.int16).reshape(-1, 6)


def md2_frame_verts(blob, f):
    h = md2_header(blob)
    o = int(h["ofs_frames"]) + f * int(h["framesize"]) + 40
    return blob[o:o + 4 * int(h["num_xyz"])].reshape(-1, 4)


def md2_st(blob):
    h = md2_header(blob)
    st = np.zeros((int(h["num_st"]), 2), np.float32)
    c.lib.bq2_synth_md2_st(c.ptr(blob), c.ptr(st))
    return st


def md3_mesh(slices=32, rings=33, num_frames=32, phase_seed=0, centre=(0.0, 0.0, 0.0)):
    V, T = sphere_dims(slices, rings)
    verts = np.zeros((num_frames, V), MD3_VERTEX)
    idx = np.zeros((T, 3), np.uint16)
    st = np.zeros((V, 2), np.float32)
    ctr = np.asarray(centre, np.float32)
    if c.lib.bq2_synth_md3_mesh(slices, rings, num_frames, phase_seed, c.ptr(ctr), c.ptr(verts), c.ptr(idx), c.ptr(st)) != 0:
        raise ValueError("bq2_synth_md3_mesh: bad arguments")
    return verts, idx, st


def scene(n_ents, lights_per_ent, num_frames, n_models=1, seed=12345, step=0):
    """World entities + per-entity world-space light origins, SURVEY 8d LCG recipe."""
    ents = np.zeros(n_ents, c.WORLD_ENTITY_DTYPE)
    lights = np.zeros((n_ents * lights_per_ent, 3), np.float32)
    c.lib.bq2_synth_scene(n_ents, lights_per_ent, num_frames, n_models, seed, step, c.ptr(ents), c.ptr(lights))
    return ents, lights


def normal2index(v):
    return int(c.lib.bq2_host_normal2index(c.ptr(np.ascontiguousarray(v, np.float32))))
