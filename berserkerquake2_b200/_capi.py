"""ctypes binding of libbq2gpu.so (include/bq2_gpu.h + include/bq2_synth.h).  Fails loudly when the
library has not been built -- there is no pure-Python or CPU substitute."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BQ2_LIB: a diagnostic build of the SAME library (make TIMELINE=1 / CHECK=1 into another file) for tools/ -- still the
# CUDA library, still no fallback
lib_path = os.environ.get("BQ2_LIB") or os.path.join(_HERE, "libbq2gpu.so")
if not os.path.exists(lib_path):
    raise ImportError(
        f"{lib_path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C berserkerquake2_b200/csrc`). The CUDA extension is the product; there is no fallback.")
lib = C.CDLL(lib_path)


class Bq2Error(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"bq2 status {status}: {text}")
        self.status = status


OK, E_INVALID, E_CUDA, E_NOMEM, E_LIMIT, E_OVERFLOW, E_NODEVICE = 0, -1, -2, -3, -4, -5, -6
WANT_SHADOW, WANT_NTB, WANT_TSLIGHT, WANT_NOLERPOUT, WANT_TABLES = 1, 2, 4, 8, 16
MODEL_SMOOTHVECS = 1
MODEL_BUILD_NEIGHBOURS = 2
ENT_SHELL = 1
(A_LERPED, A_EXTRUDED, A_FACING_BITS, A_INDICES, A_INDEX_COUNT, A_NORMALS, A_TANGENTS, A_BINORMALS,
 A_LIGHTP, A_TSH) = range(10)

# bq2_entity (64 B), bq2_pair (32 B), bq2_world_entity (56 B), bq2_world_light (16 B)
ENTITY_DTYPE = np.dtype([("model", "<i4"), ("frame", "<i4"), ("oldframe", "<i4"), ("flags", "<u4"),
                         ("backlerp", "<f4"), ("frontlerp", "<f4"), ("move", "<f4", 3),
                         ("frontv", "<f4", 3), ("backv", "<f4", 3), ("shell_scale", "<f4")])
PAIR_DTYPE = np.dtype([("entity", "<i4"), ("light", "<f4", 3), ("scale", "<f4"), ("view", "<f4", 3)])
WORLD_ENTITY_DTYPE = np.dtype([("model", "<i4"), ("frame", "<i4"), ("oldframe", "<i4"), ("rf_flags", "<u4"),
                               ("backlerp", "<f4"), ("origin", "<f4", 3), ("oldorigin", "<f4", 3),
                               ("angles", "<f4", 3)])
WORLD_LIGHT_DTYPE = np.dtype([("origin", "<f4", 3), ("radius", "<f4")])
assert ENTITY_DTYPE.itemsize == 64 and PAIR_DTYPE.itemsize == 32
assert WORLD_ENTITY_DTYPE.itemsize == 56 and WORLD_LIGHT_DTYPE.itemsize == 16


class FrameDesc(C.Structure):
    _fields_ = [("ents", C.c_void_p), ("n_ents", C.c_int32), ("pairs", C.c_void_p), ("n_pairs", C.c_int32),
                ("want", C.c_uint32), ("index_stride", C.c_uint32)]


class FrameOut(C.Structure):
    _fields_ = [("n_ent_slots", C.c_int32), ("n_pair_slots", C.c_int32),
                ("ent_slot_first", C.POINTER(C.c_int32)), ("pair_slot_first", C.POINTER(C.c_int32)),
                ("ent_vert_offset", C.POINTER(C.c_uint32)), ("pair_vert_offset", C.POINTER(C.c_uint32)),
                ("pair_bits_offset", C.POINTER(C.c_uint32)), ("ent_tb_offset", C.POINTER(C.c_uint32)),
                ("lerped", C.c_void_p), ("extruded", C.c_void_p), ("facing_bits", C.c_void_p),
                ("indices", C.c_void_p), ("index_count", C.c_void_p),
                ("normals", C.c_void_p), ("tangents", C.c_void_p), ("binormals", C.c_void_p),
                ("lightp", C.c_void_p), ("tsh", C.c_void_p),
                ("index_stride", C.c_uint32),
                ("total_lerped_verts", C.c_uint64), ("total_indices", C.c_uint64), ("total_pairs", C.c_uint64),
                ("overflow_pairs", C.c_int32), ("pair_ts_offset", C.POINTER(C.c_uint32))]


class WorldDesc(C.Structure):
    _fields_ = [("ents", C.c_void_p), ("n_ents", C.c_int32), ("lights", C.c_void_p), ("n_lights", C.c_int32),
                ("pair_ent", C.c_void_p), ("pair_light", C.c_void_p), ("n_pairs", C.c_int32),
                ("view_world", C.c_void_p), ("want", C.c_uint32), ("index_stride", C.c_uint32), ("opts", C.c_void_p)]


class RenderOpts(C.Structure):
    """bq2_render_opts: the cvars / r_mirror / pass of the caller loops (main.c:64041-64133, 63722-63727)"""
    _fields_ = [("r_mirror", C.c_int32), ("r_mirror_shadows", C.c_float), ("r_mirror_models", C.c_float),
                ("r_playershadow", C.c_float), ("r_shadows", C.c_float), ("r_drawentities", C.c_float), ("pass_", C.c_int32)]


PASS_ALL, PASS_SELF, PASS_NOSELF = 0, 1, 2
RF_NOSELFSHADOW, RF_NOCASTSHADOW, RF_DISTORT, RF_VIEWERMODEL = 0x2000, 0x80000, 0x10000000, 0x2


class BrushOut(C.Structure):
    _fields_ = [("vcache", C.c_void_p), ("icache", C.c_void_p), ("counts", C.POINTER(C.c_uint32)),
                ("vert_stride", C.c_uint32), ("index_stride", C.c_uint32)]


BRUSH_PAIR_DTYPE = np.dtype([("brush", "<i4"), ("light", "<f4", 3), ("scale", "<f4")])


class Md3Mesh(C.Structure):
    _fields_ = [("num_verts", C.c_int32), ("num_tris", C.c_int32), ("vertexes", C.c_void_p),
                ("indexes", C.c_void_p), ("neighbours", C.c_void_p), ("casts_shadow", C.c_int32)]


def _sig(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


vp, i32, u32, u64, f32 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_float
# every symbol include/bq2_gpu.h declares
EXPORTS_GPU_H = [
    _sig("bq2_abi_version", C.c_int),
    _sig("bq2_max_frames_in_flight", C.c_int),
    _sig("bq2_create", C.c_int, C.c_int, C.POINTER(vp)),
    _sig("bq2_destroy", None, vp),
    _sig("bq2_last_error", C.c_char_p, vp),
    _sig("bq2_stream", vp, vp),
    _sig("bq2_device", C.c_int, vp),
    _sig("bq2_kernel_launches", u64, vp),
    _sig("bq2_upload_md2", C.c_int, vp, vp, C.c_size_t, vp, vp, vp, vp, u32, C.POINTER(i32)),
    _sig("bq2_upload_md3", C.c_int, vp, i32, vp, C.POINTER(Md3Mesh), i32, C.POINTER(i32)),
    _sig("bq2_model_set_rf_flags", C.c_int, vp, i32, u32),
    _sig("bq2_render_opts_default", None, vp),
    _sig("bq2_host_entity_casts", C.c_int, u32, u32, C.c_int, vp),
    _sig("bq2_build_records_opts", C.c_int, vp, vp, i32, vp, i32, vp, vp, i32, vp, vp, vp, vp, C.POINTER(i32)),
    _sig("bq2_model_info", C.c_int, vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), vp, vp),
    _sig("bq2_frame", C.c_int, vp, C.POINTER(FrameDesc), C.POINTER(FrameOut)),
    _sig("bq2_frame_submit", C.c_int, vp, C.POINTER(FrameDesc)),
    _sig("bq2_frame_wait", C.c_int, vp, C.POINTER(FrameOut)),
    _sig("bq2_frame_replay", C.c_int, vp),
    _sig("bq2_frame_keep", C.c_int, vp, C.POINTER(vp)),
    _sig("bq2_frame_keep_own", C.c_int, vp, C.POINTER(vp)),
    _sig("bq2_resident_result_bytes", u64, vp),
    _sig("bq2_resident_replay", C.c_int, vp, vp),
    _sig("bq2_resident_run", C.c_int, vp, C.POINTER(vp), i32, i32, i32, C.POINTER(f32)),
    _sig("bq2_resident_run_streams", C.c_int, vp, C.POINTER(vp), i32, i32, i32, i32, C.POINTER(f32)),
    _sig("bq2_resident_download", C.c_int, vp, vp, C.c_int, u64, u64, vp),
    _sig("bq2_hash_results", C.c_int, vp, vp, vp, i32),
    _sig("bq2_profiler_range", C.c_int, vp, C.c_int),
    _sig("bq2_resident_free", None, vp, vp),
    _sig("bq2_download", C.c_int, vp, C.c_int, u64, u64, vp),
    _sig("bq2_download_indices", C.c_int, vp, i32, i32, u32, vp),
    _sig("bq2_host_index_counts", C.POINTER(u32), vp),
    _sig("bq2_expand_trisverts", C.c_int, vp, i32, vp, i32, C.POINTER(i32)),
    _sig("bq2_host_angle_vectors", None, vp, vp, vp, vp),
    _sig("bq2_host_entity_md2", None, vp, i32, i32, f32, vp, vp, vp, u32, vp),
    _sig("bq2_host_entity_md2_hdr", None, vp, vp, i32, i32, f32, vp, vp, vp, u32, vp),
    _sig("bq2_host_entity_md3", None, vp, vp, i32, i32, f32, vp, vp, vp, vp),
    _sig("bq2_host_point_to_model", None, vp, vp, vp, vp),
    _sig("bq2_host_entity_basis", C.c_int, vp, vp),
    _sig("bq2_host_point_to_model_basis", None, vp, vp, vp, vp),
    _sig("bq2_host_casts_shadow", C.c_int, u32, C.c_int, f32, f32),
    _sig("bq2_build_records", C.c_int, vp, vp, i32, vp, i32, vp, vp, i32, vp, vp, vp, C.POINTER(i32)),
    _sig("bq2_host_md2_neighbours", C.c_int, vp, i32, vp),
    _sig("bq2_host_md3_neighbours", C.c_int, vp, i32, vp),
    _sig("bq2_gpu_md2_neighbours", C.c_int, vp, vp, i32, vp),
    _sig("bq2_gpu_md3_neighbours", C.c_int, vp, vp, i32, vp),
    _sig("bq2_gpu_md2_tangents", C.c_int, vp, vp, C.c_size_t, vp, f32, C.c_int, vp, vp, vp),
    _sig("bq2_gpu_md3_tangents", C.c_int, vp, vp, i32, i32, vp, vp, i32),
    _sig("bq2_selftest_ratio", C.c_int, vp, C.c_int, C.c_float, C.c_uint64, C.c_uint64, vp),
    _sig("bq2_block_checksum", C.c_uint32, vp, C.c_size_t),
    _sig("bq2_host_md2_from_file", C.c_int, vp, C.c_size_t, f32, C.c_int, vp, C.c_size_t, vp),
    _sig("bq2_md2_cache_bytes", C.c_size_t, C.c_int, i32, i32, i32),
    _sig("bq2_md2_cache_write", C.c_int, vp, C.c_size_t, C.c_int, f32, vp, i32, i32, i32, vp, vp, vp),
    _sig("bq2_md2_cache_read", C.c_int, vp, C.c_size_t, C.c_int, f32, i32, i32, i32, vp, vp, vp, vp, C.POINTER(i32)),
    _sig("bq2_md3_cache_bytes", C.c_size_t, i32, i32),
    _sig("bq2_md3_cache_write", C.c_int, vp, C.c_size_t, u32, vp, i32, i32),
    _sig("bq2_md3_cache_read", C.c_int, vp, C.c_size_t, u32, vp, i32, i32),
    _sig("bq2_md3_file_info", C.c_int, vp, C.c_size_t, C.POINTER(i32), C.POINTER(i32), vp, vp),
    _sig("bq2_host_md3_mesh_from_file", C.c_int, vp, C.c_size_t, i32, f32, C.c_int, vp, vp, vp, vp),
    _sig("bq2_upload_brush", C.c_int, vp, i32, vp, vp, vp, vp, C.POINTER(i32)),
    _sig("bq2_brush_volumes", C.c_int, vp, vp, i32, vp, C.POINTER(BrushOut)),
    _sig("bq2_brush_download", C.c_int, vp, i32, vp, vp),
    _sig("bq2_host_alloc", vp, vp, C.c_size_t),
    _sig("bq2_host_free", None, vp, vp),
    _sig("bq2_world_submit", C.c_int, vp, C.POINTER(WorldDesc)),
    _sig("bq2_world_frame", C.c_int, vp, C.POINTER(WorldDesc), C.POINTER(FrameOut)),
    _sig("bq2_shard_entities", None, i32, i32, i32, C.POINTER(i32), C.POINTER(i32)),
]
EXPORTS_SYNTH_H = [
    _sig("bq2_synth_sphere_verts", i32, i32, i32),
    _sig("bq2_synth_sphere_tris", i32, i32, i32),
    _sig("bq2_synth_md2_bytes", C.c_size_t, i32, i32, i32),
    _sig("bq2_synth_md2", C.c_int, i32, i32, i32, u32, vp),
    _sig("bq2_synth_md2_st", None, vp, vp),
    _sig("bq2_synth_md3_mesh", C.c_int, i32, i32, i32, u32, vp, vp, vp, vp),
    _sig("bq2_host_normal2index", C.c_uint8, vp),
    _sig("bq2_host_anorms", vp),
    _sig("bq2_synth_scene", None, i32, i32, i32, i32, u32, u32, vp, vp),
]


def ptr(a):
    """numpy array (or None) -> void*"""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
