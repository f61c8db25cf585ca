"""Context / FrameResult: the Python face of ``bq2_ctx`` (include/bq2_gpu.h).

One Context per GPU (one process per GPU under torch.distributed).  Inputs are numpy structured
arrays with the C record layouts; results stay in HBM and are fetched on demand with
``FrameResult.lerped(slot)`` etc. (tests) or consumed in place through the device pointers in
``FrameResult.out`` (a renderer).  Nothing here computes: no numpy arithmetic stands in for a kernel.
"""
import ctypes as C
import numpy as np
from . import _capi as c


class Context:
    def __init__(self, device=0):
        h = C.c_void_p()
        st = c.lib.bq2_create(int(device), C.byref(h))
        if st != c.OK:
            raise c.Bq2Error(st, (c.lib.bq2_last_error(None) or b"").decode())
        self._h = h
        self.device = int(device)
        self._keep = []          # host arrays borrowed by the C side for the duration of a call

    # -- lifetime ------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            lib = getattr(c, "lib", None)          # module globals may already be gone at interpreter shutdown
            if lib is not None:
                lib.bq2_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st, allow=()):
        if st != c.OK and st not in allow:
            raise c.Bq2Error(st, (c.lib.bq2_last_error(self._h) or b"").decode())
        return st

    @property
    def stream(self):
        """cudaStream_t (int) of the most recently submitted frame (fixed unless frames are overlapped)."""
        return c.lib.bq2_stream(self._h) or 0

    @property
    def kernel_launches(self):
        return int(c.lib.bq2_kernel_launches(self._h))

    # -- models --------------------------------------------------------------------------------------
    def upload_md2(self, blob, neighbours, tangents=None, binormals=None, normals=None, smooth=True):
        """blob: uint8 MD2 image as the reference holds it after Mod_LoadAliasModel (main.c:51216);
        neighbours: int32 [T][3] (model_t.triangles)."""
        blob = np.ascontiguousarray(blob, np.uint8)
        nb = None if neighbours is None or isinstance(neighbours, str) else np.ascontiguousarray(neighbours, np.int32)
        tb = [None if a is None else np.ascontiguousarray(a, np.uint8) for a in (tangents, binormals, normals)]
        mid = C.c_int32(-1)
        flags = (c.MODEL_SMOOTHVECS if smooth else 0) | (c.MODEL_BUILD_NEIGHBOURS if isinstance(neighbours, str) and neighbours == "gpu" else 0)
        self._check(c.lib.bq2_upload_md2(self._h, c.ptr(blob), blob.nbytes, c.ptr(nb), c.ptr(tb[0]), c.ptr(tb[1]),
                                         c.ptr(tb[2]), flags, C.byref(mid)))
        return mid.value

    def upload_md3(self, frame_translate, meshes):
        """meshes: list of dicts {verts: MD3_VERTEX [F][V], indexes: uint16 [T][3], neighbours: int32 [T][3],
        casts_shadow: bool}."""
        ft = np.ascontiguousarray(frame_translate, np.float32)
        arr = (c.Md3Mesh * len(meshes))()
        keep = []
        for i, m in enumerate(meshes):
            v = np.ascontiguousarray(m["verts"]); ix = np.ascontiguousarray(m["indexes"], np.uint16)
            gpu_nb = isinstance(m.get("neighbours"), str) and m["neighbours"] == "gpu"
            nb = None if m.get("neighbours") is None or gpu_nb else np.ascontiguousarray(m["neighbours"], np.int32)
            keep += [v, ix, nb]
            assert v.shape[0] == ft.shape[0]
            arr[i].num_verts = v.shape[1]; arr[i].num_tris = ix.shape[0]
            arr[i].vertexes = v.ctypes.data; arr[i].indexes = ix.ctypes.data
            arr[i].neighbours = None if nb is None else nb.ctypes.data
            arr[i].casts_shadow = (1 if m.get("casts_shadow", True) else 0) | (2 if gpu_nb else 0)
        mid = C.c_int32(-1)
        self._check(c.lib.bq2_upload_md3(self._h, ft.shape[0], c.ptr(ft), arr, len(meshes), C.byref(mid)))
        return mid.value

    def gpu_md2_neighbours(self, blob):
        """model_t.triangles computed on the device (findneighbour rules, main.c:61782-61824)."""
        h = np.ascontiguousarray(blob, np.uint8)[:68].view("<i4")
        T, o = int(h[8]), int(h[13])
        tris = np.ascontiguousarray(blob[o:o + 12 * T]).view(np.int16)
        out = np.zeros((T, 3), np.int32)
        self._check(c.lib.bq2_gpu_md2_neighbours(self._h, c.ptr(tris), T, c.ptr(out)))
        return out

    def gpu_md3_neighbours(self, indexes):
        """maliasmesh_t.triangles computed on the device (R_BuildTriangleNeighbors, main.c:68646-68696)."""
        ix = np.ascontiguousarray(indexes, np.uint16)
        out = np.zeros((ix.size // 3, 3), np.int32)
        self._check(c.lib.bq2_gpu_md3_neighbours(self._h, c.ptr(ix), ix.size // 3, c.ptr(out)))
        return out

    def gpu_md2_tangents(self, blob, st, smooth_cosine, smooth=True):
        """model_t.tangents / binormals (/ normals when not smooth) computed on the device
        (main.c:51661-51804).  Returns (normals or None, tangents, binormals), uint8 [frames][V or T]."""
        blob = np.ascontiguousarray(blob, np.uint8)
        h = blob[:68].view("<i4")
        V, T, F = int(h[6]), int(h[8]), int(h[10])
        rows = V if smooth else T
        st = np.ascontiguousarray(st, np.float32)
        n = None if smooth else np.zeros((F, rows), np.uint8)
        t, b = np.zeros((F, rows), np.uint8), np.zeros((F, rows), np.uint8)
        self._check(c.lib.bq2_gpu_md2_tangents(self._h, c.ptr(blob), blob.nbytes, c.ptr(st), float(smooth_cosine),
                                               1 if smooth else 0, c.ptr(n), c.ptr(t), c.ptr(b)))
        return n, t, b

    def gpu_md3_tangents(self, verts, st, indexes):
        """fills the tangent / binormal bytes of verts [F][V] (MD3 vertex dtype) in place (main.c:50993-51014)."""
        assert verts.flags["C_CONTIGUOUS"]
        F, V = verts.shape
        st = np.ascontiguousarray(st, np.float32); ix = np.ascontiguousarray(indexes, np.uint16)
        self._check(c.lib.bq2_gpu_md3_tangents(self._h, c.ptr(verts), F, V, c.ptr(st), c.ptr(ix), ix.size // 3))

    def selftest_ratio(self, mode, num=1.0, count=0, seed=0):
        """bq2_selftest_ratio: mismatches between the kernels' num / sqrtf(len2) sequence and div.rn(sqrt.rn)."""
        bad = C.c_uint64(0)
        self._check(c.lib.bq2_selftest_ratio(self._h, int(mode), float(num), int(count), int(seed), C.byref(bad)))
        return int(bad.value)

    # -- brush-model volumes (main.c:63326-63499) -------------------------------------------------------
    def upload_brush(self, numverts, verts, neighbours, fence=None):
        nv = np.ascontiguousarray(numverts, np.int32); v = np.ascontiguousarray(verts, np.float32)
        nb = np.ascontiguousarray(neighbours, np.int32)
        fe = None if fence is None else np.ascontiguousarray(fence, np.uint8)
        bid = C.c_int32(-1)
        self._check(c.lib.bq2_upload_brush(self._h, len(nv), c.ptr(nv), c.ptr(v), c.ptr(nb), c.ptr(fe), C.byref(bid)))
        return bid.value

    def brush_volumes(self, pairs, lit):
        """pairs: BRUSH_PAIR_DTYPE array (light in model space); lit: uint8, the pairs' polygon flags concatenated.
        Returns a list of (vcache [vb*2][3], icache [ib]) per pair."""
        pr = np.ascontiguousarray(pairs, c.BRUSH_PAIR_DTYPE); li = np.ascontiguousarray(lit, np.uint8)
        out = c.BrushOut()
        self._check(c.lib.bq2_brush_volumes(self._h, c.ptr(pr), len(pr), c.ptr(li), C.byref(out)))
        res = []
        for k in range(len(pr)):
            vb, ib = int(out.counts[2 * k]), int(out.counts[2 * k + 1])
            vc = np.zeros((2 * vb, 3), np.float32); ic = np.zeros(ib, np.uint32)
            self._check(c.lib.bq2_brush_download(self._h, k, c.ptr(vc), c.ptr(ic)))
            res.append((vc, ic))
        return res

    # -- per frame -----------------------------------------------------------------------------------
    def set_model_rf_flags(self, model_id, rf_flags):
        """model_t.flags bits the caller loop reads (RF_NOCASTSHADOW, RF_DISTORT, RF_NOSELFSHADOW)."""
        self._check(c.lib.bq2_model_set_rf_flags(self._h, int(model_id), int(rf_flags)))

    @staticmethod
    def render_opts(**kw):
        """bq2_render_opts at the engine's defaults, fields overridden by keyword (pass_ = PASS_ALL / SELF / NOSELF)."""
        o = c.RenderOpts()
        c.lib.bq2_render_opts_default(C.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def build_records(self, world_ents, world_lights, pair_ent, pair_light, view_world=None, opts=None):
        """The renderer's (light x entity) loop nest flattened into bq2_entity / bq2_pair records, on the
        host in C (bq2_build_records / bq2_build_records_opts)."""
        we = np.ascontiguousarray(world_ents, c.WORLD_ENTITY_DTYPE)
        wl = np.ascontiguousarray(world_lights, c.WORLD_LIGHT_DTYPE)
        pe = np.ascontiguousarray(pair_ent, np.int32); pl = np.ascontiguousarray(pair_light, np.int32)
        vw = None if view_world is None else np.ascontiguousarray(view_world, np.float32)
        ents = np.zeros(len(we), c.ENTITY_DTYPE); pairs = np.zeros(len(pe), c.PAIR_DTYPE)
        kept = C.c_int32(0)
        if opts is None:
            self._check(c.lib.bq2_build_records(self._h, c.ptr(we), len(we), c.ptr(wl), len(wl), c.ptr(pe), c.ptr(pl),
                                                len(pe), c.ptr(vw), c.ptr(ents), c.ptr(pairs), C.byref(kept)))
        else:
            self._check(c.lib.bq2_build_records_opts(self._h, c.ptr(we), len(we), c.ptr(wl), len(wl), c.ptr(pe), c.ptr(pl),
                                                     len(pe), c.ptr(vw), C.cast(C.byref(opts), C.c_void_p), c.ptr(ents),
                                                     c.ptr(pairs), C.byref(kept)))
        return ents, pairs[:kept.value]

    def _desc(self, ents, pairs, want, index_stride):
        ents = np.ascontiguousarray(ents, c.ENTITY_DTYPE)
        pairs = np.zeros(0, c.PAIR_DTYPE) if pairs is None else np.ascontiguousarray(pairs, c.PAIR_DTYPE)
        self._keep = [ents, pairs]
        d = c.FrameDesc()
        d.ents = ents.ctypes.data if len(ents) else None; d.n_ents = len(ents)
        d.pairs = pairs.ctypes.data if len(pairs) else None; d.n_pairs = len(pairs)
        d.want = int(want); d.index_stride = int(index_stride)
        return d

    def frame(self, ents, pairs=None, want=c.WANT_SHADOW, index_stride=0, allow_overflow=False):
        d = self._desc(ents, pairs, want, index_stride)
        out = c.FrameOut()
        st = self._check(c.lib.bq2_frame(self._h, C.byref(d), C.byref(out)),
                         allow=(c.E_OVERFLOW,) if allow_overflow else ())
        return FrameResult(self, out, st)

    def submit(self, ents, pairs=None, want=c.WANT_SHADOW, index_stride=0):
        d = self._desc(ents, pairs, want, index_stride)
        self._check(c.lib.bq2_frame_submit(self._h, C.byref(d)))

    def wait(self, allow_overflow=False):
        out = c.FrameOut()
        st = self._check(c.lib.bq2_frame_wait(self._h, C.byref(out)), allow=(c.E_OVERFLOW,) if allow_overflow else ())
        return FrameResult(self, out, st)

    def world_desc(self, world_ents, world_lights, pair_ent, pair_light, view_world=None, want=c.WANT_SHADOW,
                   index_stride=0, opts=None):
        we = np.ascontiguousarray(world_ents, c.WORLD_ENTITY_DTYPE)
        wl = np.ascontiguousarray(world_lights, c.WORLD_LIGHT_DTYPE)
        pe = np.ascontiguousarray(pair_ent, np.int32); pl = np.ascontiguousarray(pair_light, np.int32)
        vw = None if view_world is None else np.ascontiguousarray(view_world, np.float32)
        self._keep = [we, wl, pe, pl, vw, opts]
        d = c.WorldDesc()
        d.opts = None if opts is None else C.addressof(opts)
        d.ents = we.ctypes.data if len(we) else None; d.n_ents = len(we)
        d.lights = wl.ctypes.data if len(wl) else None; d.n_lights = len(wl)
        d.pair_ent = pe.ctypes.data if len(pe) else None; d.pair_light = pl.ctypes.data if len(pl) else None
        d.n_pairs = len(pe); d.view_world = None if vw is None else vw.ctypes.data
        d.want = int(want); d.index_stride = int(index_stride)
        return d

    def world_frame(self, world_ents, world_lights, pair_ent, pair_light, view_world=None,
                    want=c.WANT_SHADOW | c.WANT_TABLES, index_stride=0, allow_overflow=False, opts=None):
        """bq2_world_frame: records built on the device; pair slot p is caller pair p."""
        d = self.world_desc(world_ents, world_lights, pair_ent, pair_light, view_world, want, index_stride, opts)
        out = c.FrameOut()
        st = self._check(c.lib.bq2_world_frame(self._h, C.byref(d), C.byref(out)),
                         allow=(c.E_OVERFLOW,) if allow_overflow else ())
        return FrameResult(self, out, st)

    def host_array(self, like):
        """A copy of `like` in page-locked memory from bq2_host_alloc (numpy view; valid until close()).  World
        descriptors whose arrays all come from here are copied to the device without the staging copy."""
        like = np.ascontiguousarray(like)
        p = c.lib.bq2_host_alloc(self._h, max(like.nbytes, 16))
        if not p:
            raise MemoryError(c.lib.bq2_last_error(self._h))
        buf = (C.c_char * max(like.nbytes, 16)).from_address(p)
        a = np.frombuffer(buf, dtype=like.dtype, count=like.size).reshape(like.shape)
        a[...] = like
        return a

    def world_submit(self, desc):
        """bq2_world_submit of a descriptor made by world_desc (keep it alive); pair with wait()."""
        self._check(c.lib.bq2_world_submit(self._h, C.byref(desc)))

    def replay(self):
        """Re-launch the last frame's kernel on the records already in HBM (no copies, no sync)."""
        self._check(c.lib.bq2_frame_replay(self._h))

    def keep(self, own_results=False):
        """bq2_frame_keep / bq2_frame_keep_own: a handle to the last (host-built, completed) frame's records, for
        replay_kept() / run_ring(); own_results: the kept frame gets result arrays of its own."""
        h = C.c_void_p()
        self._check((c.lib.bq2_frame_keep_own if own_results else c.lib.bq2_frame_keep)(self._h, C.byref(h)))
        return h

    def replay_kept(self, handle):
        self._check(c.lib.bq2_resident_replay(self._h, handle))

    def run_ring(self, handles, start, k, timed=False, streams=1):
        """bq2_resident_run: k replays launched from C, handles[(start + i) % len]; timed: returns the milliseconds
        between two CUDA events on the stream around the k launches (the call waits for the second).  streams > 1:
        bq2_resident_run_streams, the launches dealt round-robin onto that many streams."""
        arr = (C.c_void_p * len(handles))(*[h.value if isinstance(h, C.c_void_p) else h for h in handles])
        ms = C.c_float(0.0)
        if streams > 1:
            self._check(c.lib.bq2_resident_run_streams(self._h, arr, len(handles), int(start), int(k), int(streams), C.byref(ms)))
            return float(ms.value) if timed else None
        self._check(c.lib.bq2_resident_run(self._h, arr, len(handles), int(start), int(k), C.byref(ms) if timed else None))
        return float(ms.value) if timed else None

    def download_kept(self, handle, which, first, count, dtype):
        buf = np.zeros(int(count), dtype)
        if count:
            self._check(c.lib.bq2_resident_download(self._h, handle, which, int(first), int(count), c.ptr(buf)))
        return buf

    def hash_results(self, n_pair_slots, handle=None):
        """bq2_hash_results: one uint64 per pair slot over the frame's lerped / extruded / facing / count / stream."""
        out = np.zeros(int(n_pair_slots), np.uint64)
        self._check(c.lib.bq2_hash_results(self._h, handle, c.ptr(out), int(n_pair_slots)))
        return out

    def profiler_range(self, on):
        self._check(c.lib.bq2_profiler_range(self._h, 1 if on else 0))

    def free_kept(self, handle):
        c.lib.bq2_resident_free(self._h, handle)

    def download(self, which, first, count, dtype):
        buf = np.zeros(int(count), dtype)
        assert buf.itemsize == 4
        if count:
            self._check(c.lib.bq2_download(self._h, which, int(first), int(count), c.ptr(buf)))
        return buf


class FrameResult:
    """Where one frame's results live (device) plus host-side slot tables."""

    def __init__(self, ctx, out, status):
        self.ctx, self.out, self.status = ctx, out, status
        nes, nps = out.n_ent_slots, out.n_pair_slots
        self.n_ent_slots, self.n_pair_slots = nes, nps
        self.ent_vert_offset = np.ctypeslib.as_array(out.ent_vert_offset, (nes + 1,)).copy()
        self.ent_tb_offset = np.ctypeslib.as_array(out.ent_tb_offset, (nes + 1,)).copy()
        if out.pair_vert_offset and nps:      # absent for device-built frames submitted without WANT_TABLES
            n_tab = nps + 1 if out.pair_slot_first else nps
            self.pair_vert_offset = np.ctypeslib.as_array(out.pair_vert_offset, (n_tab,)).copy()
            self.pair_bits_offset = np.ctypeslib.as_array(out.pair_bits_offset, (n_tab,)).copy()
        else:
            self.pair_vert_offset = self.pair_bits_offset = None
        self.pair_ts_offset = (np.ctypeslib.as_array(out.pair_ts_offset, (nps + 1,)).copy()
                               if out.pair_ts_offset and nps else None)
        self.index_stride = out.index_stride
        self.total_lerped_verts = out.total_lerped_verts
        self.total_indices = out.total_indices
        self.overflow_pairs = out.overflow_pairs
        hc = c.lib.bq2_host_index_counts(ctx._h)
        self.index_counts = np.ctypeslib.as_array(hc, (nps,)).copy() if nps else np.zeros(0, np.uint32)

    # rows are padded to 4 vertices; callers pass the true V / T
    def lerped(self, ent_slot, V):
        return self.ctx.download(c.A_LERPED, 3 * int(self.ent_vert_offset[ent_slot]), 3 * V, np.float32).reshape(V, 3)

    def extruded(self, pair_slot, V):
        return self.ctx.download(c.A_EXTRUDED, 3 * int(self.pair_vert_offset[pair_slot]), 3 * V, np.float32).reshape(V, 3)

    def facing(self, pair_slot, T):
        w = self.ctx.download(c.A_FACING_BITS, int(self.pair_bits_offset[pair_slot]), (T + 31) // 32, np.uint32)
        return ((w[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(np.uint8).reshape(-1)[:T]

    def indices(self, pair_slot):
        n = min(int(self.index_counts[pair_slot]), int(self.index_stride))
        return self.ctx.download(c.A_INDICES, int(pair_slot) * int(self.index_stride), n, np.uint32)

    def ntb(self, ent_slot, rows):
        o = 3 * int(self.ent_tb_offset[ent_slot])
        return tuple(self.ctx.download(w, o, 3 * rows, np.float32).reshape(rows, 3)
                     for w in (c.A_NORMALS, c.A_TANGENTS, c.A_BINORMALS))

    def tslight(self, pair_slot, rows):
        """rows = V (per-vertex rows) or 3T (per-corner rows of a non-smooth MD2)"""
        o = 3 * int(self.pair_ts_offset[pair_slot])
        return tuple(self.ctx.download(w, o, 3 * rows, np.float32).reshape(rows, 3) for w in (c.A_LIGHTP, c.A_TSH))

    def trisverts(self, pair_slot, max_verts=36 * 8192 // 3):
        """The reference's expanded stream (TrisVerts, TrisCounter) for one pair slot."""
        buf = np.zeros((max_verts, 3), np.float32)
        n = C.c_int32(0)
        self.ctx._check(c.lib.bq2_expand_trisverts(self.ctx._h, int(pair_slot), c.ptr(buf), max_verts, C.byref(n)))
        return buf[:n.value].copy()
