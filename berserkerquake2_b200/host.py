"""Host-side C helpers (bq2_host.c): the trigonometry and per-entity constants the reference computes
on the CPU (main.c:63531-63553, 63733-63745, 63825-63856) -- same libm, same operation order."""
import numpy as np
from . import _capi as c


def _f3(a):
    return np.ascontiguousarray(a, dtype=np.float32).reshape(3)


def angle_vectors(angles):
    f, r, u = (np.zeros(3, np.float32) for _ in range(3))
    c.lib.bq2_host_angle_vectors(c.ptr(_f3(angles)), c.ptr(f), c.ptr(r), c.ptr(u))
    return f, r, u


def entity_md2(blob, frame, oldframe, backlerp, origin, oldorigin, angles, rf_flags=0, model=0):
    out = np.zeros(1, c.ENTITY_DTYPE)
    c.lib.bq2_host_entity_md2(c.ptr(blob), frame, oldframe, float(backlerp), c.ptr(_f3(origin)),
                              c.ptr(_f3(oldorigin)), c.ptr(_f3(angles)), int(rf_flags), c.ptr(out))
    out["model"] = model
    return out[0]


def entity_md3(frame_translate, frame, oldframe, backlerp, origin, oldorigin, angles, model=0):
    ft = np.ascontiguousarray(frame_translate, np.float32)
    out = np.zeros(1, c.ENTITY_DTYPE)
    c.lib.bq2_host_entity_md3(c.ptr(ft[frame]), c.ptr(ft[oldframe]), frame, oldframe, float(backlerp),
                              c.ptr(_f3(origin)), c.ptr(_f3(oldorigin)), c.ptr(_f3(angles)), c.ptr(out))
    out["model"] = model
    return out[0]


def point_to_model(world, origin, angles):
    out = np.zeros(3, np.float32)
    c.lib.bq2_host_point_to_model(c.ptr(_f3(world)), c.ptr(_f3(origin)), c.ptr(_f3(angles)), c.ptr(out))
    return out


def casts_shadow(rf_flags, r_mirror=False, r_playershadow=0.0, r_shadows=2.0):
    return bool(c.lib.bq2_host_casts_shadow(int(rf_flags), int(r_mirror), float(r_playershadow), float(r_shadows)))


def shard_entities(n_ents, rank, world):
    import ctypes as C
    first, count = C.c_int32(), C.c_int32()
    c.lib.bq2_shard_entities(n_ents, rank, world, C.byref(first), C.byref(count))
    return first.value, count.value


def md2_neighbours(blob):
    """model_t.triangles for an in-memory MD2 image (findneighbour rules, main.c:61782-61824, 51643-51653)."""
    h = blob[:68].view("<i4")
    T, o = int(h[8]), int(h[13])
    tris = np.ascontiguousarray(blob[o:o + 12 * T]).view(np.int16)
    out = np.zeros((T, 3), np.int32)
    if c.lib.bq2_host_md2_neighbours(c.ptr(tris), T, c.ptr(out)) != 0:
        raise MemoryError("bq2_host_md2_neighbours")
    return out


def md3_neighbours(indexes):
    """maliasmesh_t.triangles (R_BuildTriangleNeighbors, main.c:68646-68696)."""
    ix = np.ascontiguousarray(indexes, np.uint16)
    T = ix.size // 3
    out = np.zeros((T, 3), np.int32)
    if c.lib.bq2_host_md3_neighbours(c.ptr(ix), T, c.ptr(out)) != 0:
        raise MemoryError("bq2_host_md3_neighbours")
    return out
