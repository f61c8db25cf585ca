/*
 * bq2_gpu.h -- C ABI of the B200-native alias-model lerp + stencil-shadow-volume path.
 *
 * This is the drop-in boundary for ONE hot path of Berserker@Quake2 (reference citations are
 * file:line under the reference tree; "M:" = main.c):
 *
 *   R_CalcAliasFrameLerp          M:63510-63583   MD2 keyframe lerp (+ shell push-out)
 *   R_CalcAliasFrameLerpShadow    M:63586-63631   extrusion + triangle normal + light-facing
 *   R_CalcAliasFrameShadowVolume  M:63634-63710   silhouette edges + caps  -> vertex stream
 *   R_DrawAliasShadowVolume       M:63713-63797   light -> model space, drives the two above
 *   R_DrawMD3AliasShadowVolume    M:63800-64038   the same, fused, per MD3 mesh
 *   GL_DrawAliasFrameLerpLightARB_vp  M:65140-65267 / GL_DrawMD3...ARB_vp M:65671-65828  N/T/B lerp
 *   GL_DrawAliasFrameLerpLightARB     M:64889-65017 / GL_DrawMD3...ARB    M:65839-65898  LightP / tsH
 *
 * The reference has no plugin layer: those functions talk through file-scope globals
 * (currententity, currentmodel, currentshadowlight, tempVertexArray, extrudedVerts, viss,
 * TrisVerts, TrisCounter; data.h:624, 2626, 2709, 2731-2736, 2827).  The replacement is a batched
 * API: the renderer's per-light / per-entity loops (R_DrawEntsShadowVolumes M:64041-64133) collect
 * (entity, light) pairs, the trigonometry stays on the host in C (bq2_host_* below restate the
 * reference's own host lines so libm's sincosf is the same code on both sides), and one CUDA
 * launch does every pair.  integration/bq2_dropin.c shows the reference
 * signatures re-implemented on top of this header (batch of one).
 *
 * Of the five signatures SURVEY 8(b) lists, THREE are exported as symbols by the shim (integration/bq2_dropin.c, linked
 * in front of the engine): R_CalcAliasFrameLerp, R_CalcAliasFrameLerpShadow, R_CalcAliasFrameShadowVolume -- the engine's
 * own, unmodified R_DrawAliasShadowVolume then runs on the GPU (tests/test_dropin.py, bench.py e2e.dropin).  The other
 * TWO are NOT symbols of this project, on purpose: R_DrawAliasShadowVolume (M:63713) and R_DrawMD3AliasShadowVolume
 * (M:63800) are mostly OpenGL state and draw calls, which stay the engine's; the first needs no change at all, the second
 * is monolithic and takes the five-line patch of INTEGRATION.md section 1 that calls bq2_dropin_md3_mesh() for its
 * per-mesh compute block (M:63885-63991).  Their filters and loops are available batched: bq2_render_opts /
 * bq2_host_entity_casts (M:63722-63727, 64041-64133), bq2_build_records / bq2_world_submit.
 *
 * Plain C, plain pointers and sizes; no torch / C++ types.  All functions return BQ2_OK (0) or a
 * negative bq2_status; bq2_last_error() gives text.  Nothing here ever longjmps or aborts and
 * there is NO CPU fallback: without a CUDA device bq2_create() fails with BQ2_E_NODEVICE.
 * A ctx is bound to one device and one stream and is not thread-safe; multi-GPU = one ctx per
 * device (one process per GPU under torch.distributed, or one host thread per ctx).
 */
#ifndef BQ2_GPU_H
#define BQ2_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BQ2_ABI_VERSION 2      /* 2: bq2_world_desc.opts (render options of the caller loops) */

typedef enum bq2_status {
    BQ2_OK          = 0,
    BQ2_E_INVALID   = -1,   /* bad argument / malformed model blob                                  */
    BQ2_E_CUDA      = -2,   /* CUDA runtime error, text in bq2_last_error                           */
    BQ2_E_NOMEM     = -3,   /* host or device allocation failed                                     */
    BQ2_E_LIMIT     = -4,   /* exceeds a reference limit (MAX_VERTS 2048, MAX_TRIANGLES 4096, ...)   */
                            /* ... or a frame whose output rows do not fit 32-bit row numbers (2^32 vertex rows) */
    BQ2_E_OVERFLOW  = -5,   /* some pair produced more indices than index_stride holds              */
    BQ2_E_NODEVICE  = -6    /* no usable CUDA device: there is no CPU path                          */
} bq2_status;

/* Reference limits (defs.h:3455-3456 MD2; defs.h:560-565 MD3 per mesh). */
#define BQ2_MD2_MAX_VERTS      2048
#define BQ2_MD2_MAX_TRIANGLES  4096
#define BQ2_MD3_MAX_VERTS      4096
#define BQ2_MD3_MAX_TRIANGLES  8192
#define BQ2_MD3_MAX_MESHES     32
#define BQ2_NUM_ANORMS         162

/* Entity render flags that matter on this path (values as in the reference's defs.h RF_* set). */
#define BQ2_RF_VIEWERMODEL   0x00000002u
#define BQ2_RF_WEAPONMODEL   0x00000004u
#define BQ2_RF_FULLBRIGHT    0x00000008u
#define BQ2_RF_TRANSLUCENT   0x00000020u
#define BQ2_RF_BEAM          0x00000080u
#define BQ2_RF_SHELL_RED     0x00000400u
#define BQ2_RF_SHELL_GREEN   0x00000800u
#define BQ2_RF_SHELL_BLUE    0x00001000u
#define BQ2_RF_NOSELFSHADOW  0x00002000u   /* model_t.flags: which of the two self-shadow passes draws it, M:64078-64083 */
#define BQ2_RF_NOCASTSHADOW  0x00080000u   /* entity OR model flag, filtered by the caller loop, M:64089 */
#define BQ2_RF_DISTORT       0x10000000u   /* entity OR model flag, M:64089 */

/* ---- the caller loops' switches: R_DrawEntsShadowVolumes M:64041-64133 and the early-outs of
 *      R_DrawAliasShadowVolume / R_DrawMD3AliasShadowVolume M:63722-63727, 63811-63816 -------------------------- */
typedef enum bq2_pass {
    BQ2_PASS_ALL    = 0,    /* R_DrawEntsShadowVolumes(false, true): every caster                                 */
    BQ2_PASS_SELF   = 1,    /* (true, false): brush models + alias models WITHOUT RF_NOSELFSHADOW (M:64078-64080) */
    BQ2_PASS_NOSELF = 2     /* (false, false): alias models WITH RF_NOSELFSHADOW only (M:64081-64082)             */
} bq2_pass;
typedef struct bq2_render_opts {
    int32_t r_mirror;               /* data.h r_mirror: the mirrored scene is being drawn                        */
    float   r_mirror_shadows;       /* cvar, default 2 (M:4536)                                                   */
    float   r_mirror_models;        /* cvar, default 1 (M:4525)                                                   */
    float   r_playershadow;         /* cvar, default 0 (M:4390)                                                   */
    float   r_shadows;              /* cvar, default 2 (M:4383): alias volumes only when > 1                      */
    float   r_drawentities;         /* cvar, default 1 (M:4556)                                                   */
    int32_t pass;                   /* bq2_pass                                                                   */
} bq2_render_opts;
void bq2_render_opts_default(bq2_render_opts *o);       /* the engine's defaults above, r_mirror 0, BQ2_PASS_ALL */
/* 1 = the loop reaches the entity's volume builder and the builder does not return early.  model_kind: 0 alias
 * (MD2 / MD3; the per-mesh skin filter M:63862-63880 is bq2_md3_mesh.casts_shadow), 1 brush model (M:64066-64071:
 * entity / model flags are not looked at).  model_rf_flags = model_t.flags.  opts NULL = defaults. */
int  bq2_host_entity_casts(uint32_t ent_rf_flags, uint32_t model_rf_flags, int model_kind, const bq2_render_opts *opts);

/* bq2_entity.flags (ours) */
#define BQ2_ENT_SHELL        0x1u   /* add anorm[cur.lightnormalindex]*shell_scale  (M:63565-63572) */

/* model flags */
#define BQ2_MODEL_BUILD_NEIGHBOURS 0x2u /* neighbours == NULL: build the table on the device (bq2_gpu_md2_neighbours) */
#define BQ2_MODEL_SMOOTHVECS 0x1u   /* RF_SMOOTHVECS: tangent/binormal bytes are per vertex (M:51661)
                                       otherwise per triangle, plus per-triangle normal bytes (M:51747) */

typedef struct bq2_ctx bq2_ctx;

/* One entity for one frame.  move/frontv/backv are the reference's host-side lerp constants
 * (M:63531-63553 for MD2, M:63825-63856 for MD3 where frontv = frontlerp, backv = backlerp in all
 * three lanes); fill them with bq2_host_entity_md2 / bq2_host_entity_md3.  64 bytes. */
typedef struct bq2_entity {
    int32_t  model;        /* id returned by bq2_upload_md2 / bq2_upload_md3        */
    int32_t  frame;        /* currententity->frame                                   */
    int32_t  oldframe;     /* currententity->oldframe                                */
    uint32_t flags;        /* BQ2_ENT_*                                              */
    float    backlerp;     /* currententity->backlerp                                */
    float    frontlerp;    /* (float)(1.0 - backlerp), computed as the reference does */
    float    move[3];
    float    frontv[3];
    float    backv[3];
    float    shell_scale;  /* POWERSUIT_SCALE(_WEAPON) when BQ2_ENT_SHELL            */
} bq2_entity;

/* One (entity, light) pair.  light[] is the light origin ALREADY in the entity's model space
 * (M:63733-63745; bq2_host_light_to_model), scale = 32 * light radius (M:63597), view[] is
 * r_origin in model space (M:66689-66715), only read when BQ2_WANT_TSLIGHT.  32 bytes. */
typedef struct bq2_pair {
    int32_t entity;        /* index into the frame's entity array; ~index (negative) = the pair keeps its
                              slot(s) but casts nothing (index_count 0)                                   */
    float   light[3];
    float   scale;
    float   view[3];
} bq2_pair;

/* What one bq2_frame call computes. */
#define BQ2_WANT_SHADOW   0x1u  /* extruded verts, facing bits, canonical index buffer per pair   */
#define BQ2_WANT_NTB      0x2u  /* lerped normal / tangent / binormal per entity (a10)            */
#define BQ2_WANT_TSLIGHT  0x4u  /* tangent-space light vector LightP and half vector tsH per pair */
#define BQ2_WANT_NOLERPOUT 0x8u /* do not write lerped[] to memory (shadow-only callers)          */
/* every bit a caller may set (BQ2_WANT_TABLES 0x10 is declared with bq2_world_desc below); anything else in
   `want` is refused with BQ2_E_INVALID */
#define BQ2_WANT_PUBLIC_MASK 0x1fu

typedef struct bq2_frame_desc {
    const bq2_entity *ents;    int32_t n_ents;
    const bq2_pair   *pairs;   int32_t n_pairs;   /* may be 0 / NULL: lerp only                    */
    uint32_t want;                                 /* BQ2_WANT_*; lerped positions always computed  */
    uint32_t index_stride;                         /* u32 words reserved per pair (per mesh-pair for
                                                      MD3); 0 = 12 * max triangles of any model used,
                                                      the reference's own TrisVerts ratio
                                                      (data.h:2736: 36*8192 floats = 12 verts/tri)   */
} bq2_frame_desc;

/* Where the results of the last frame live.  All pointers are DEVICE pointers owned by the ctx,
 * valid until the next bq2_frame*(ctx) (see "frames in flight" below) or bq2_destroy.  "Slots": MD2 entities/pairs occupy one
 * slot; an MD3 entity/pair occupies one slot per mesh (the reference draws per mesh, M:63858).
 *   ent_slot_first[e]  .. +n meshes : slots of entity e        (n_ents + 1 entries, host array)
 *   pair_slot_first[p] .. +n meshes : slots of pair p          (n_pairs + 1 entries, host array)
 * Vertex rows are padded to a multiple of 4 vertices so every row starts 16-byte aligned. */
typedef struct bq2_frame_out {
    int32_t  n_ent_slots, n_pair_slots;
    const int32_t  *ent_slot_first;     /* HOST, n_ents+1                                            */
    const int32_t  *pair_slot_first;    /* HOST, n_pairs+1                                           */
    const uint32_t *ent_vert_offset;    /* HOST, n_ent_slots+1 : first vertex of the slot's row      */
    const uint32_t *pair_vert_offset;   /* HOST, n_pair_slots+1                                      */
    const uint32_t *pair_bits_offset;   /* HOST, n_pair_slots+1 : first u32 word of the facing bits  */
    const uint32_t *ent_tb_offset;      /* HOST, n_ent_slots+1 : first row of N/T/B (verts or tris)  */
    /* device */
    float    *lerped;          /* [ent verts][3]   tempVertexArray                                    */
    float    *extruded;        /* [pair verts][3]  extrudedVerts                                      */
    uint32_t *facing_bits;     /* viss[] packed, bit (t&31) of word t>>5                              */
    uint32_t *indices;         /* [n_pair_slots][index_stride] canonical index form, see DESIGN.md    */
    uint32_t *index_count;     /* [n_pair_slots]   == TrisCounter                                     */
    float    *normals, *tangents, *binormals;   /* [ent rows][3]  (BQ2_WANT_NTB)                      */
    float    *lightp, *tsh;                     /* [pair ts rows][3] (BQ2_WANT_TSLIGHT), see pair_ts_offset */
    uint32_t index_stride;
    /* totals, valid after bq2_frame / bq2_frame_wait */
    uint64_t total_lerped_verts;   /* sum of V over entity slots                                      */
    uint64_t total_indices;        /* sum of index_count                                              */
    uint64_t total_pairs;          /* pair slots processed                                            */
    int32_t  overflow_pairs;       /* pair slots whose true count exceeded index_stride (count is the
                                      true count; only the first index_stride words were written)     */
    const uint32_t *pair_ts_offset;     /* HOST, n_pair_slots+1 : first row of the slot in lightp / tsh.  Rows are
                                           per vertex (MD3, MD2 with RF_SMOOTHVECS) or per triangle corner, 3T of
                                           them in (triangle, corner) order, for a non-smooth MD2 (M:64943-65017) */
} bq2_frame_out;

/* ---- lifetime ------------------------------------------------------------------------------ */
int          bq2_abi_version(void);
int          bq2_max_frames_in_flight(void);               /* submits that may precede the wait for the oldest frame */
bq2_status   bq2_create(int device, bq2_ctx **out);
void         bq2_destroy(bq2_ctx *ctx);
const char  *bq2_last_error(const bq2_ctx *ctx);          /* ctx may be NULL: last create error   */
void        *bq2_stream(bq2_ctx *ctx);                     /* cudaStream_t of the most recently submitted frame (each of
                                                              the frame states has its own -- up to sixteen frames in
                                                              flight; fixed for callers that never overlap frames) */
int          bq2_device(const bq2_ctx *ctx);
uint64_t     bq2_kernel_launches(const bq2_ctx *ctx);      /* kernels launched by this ctx so far  */

/* ---- model upload (replaces what Mod_LoadAliasModel M:51216 / Mod_LoadAliasMD3Model M:50584
 *      leave in model_t for this path) ------------------------------------------------------- */
/* dmdl_blob: MD2 image in memory exactly as the reference holds it after load (dmdl_t header,
 * frames at ofs_frames with scale/translate already multiplied by the load scale, dtriangle_t at
 * ofs_tris; defs.h:3682-3726).  neighbours: int[num_tris][3] (model_t.triangles, defs.h:3486).
 * tangents/binormals(/normals): model_t.tangents/binormals/normals bytes, [frames][num_xyz] when
 * BQ2_MODEL_SMOOTHVECS else [frames][num_tris]; may be NULL when N/T/B is never requested. */
bq2_status bq2_upload_md2(bq2_ctx *ctx, const void *dmdl_blob, size_t blob_bytes,
                          const int32_t *neighbours, const uint8_t *tangents,
                          const uint8_t *binormals, const uint8_t *normals,
                          uint32_t model_flags, int32_t *model_id);

/* One MD3 mesh as the reference keeps it in memory (maliasmesh_t, defs.h:4426-4461). */
typedef struct bq2_md3_mesh {
    int32_t num_verts, num_tris;
    const void     *vertexes;   /* maliasvertex_t[num_frames*num_verts]: float point[3];
                                   u8 normal, tangent, binormal, pad  (16 bytes, defs.h:4401-4407) */
    const uint16_t *indexes;    /* WORD[3*num_tris]                                                 */
    const int32_t  *neighbours; /* neighbours_t[num_tris]                                           */
    int32_t casts_shadow;       /* bit 0: 0 = the per-mesh skin filter M:63862-63880 rejects this mesh;
                                   bit 1: neighbours == NULL -> build them on the device               */
} bq2_md3_mesh;

bq2_status bq2_upload_md3(bq2_ctx *ctx, int32_t num_frames, const float *frame_translate /*[F][3]*/,
                          const bq2_md3_mesh *meshes, int32_t num_meshes, int32_t *model_id);

/* model_t.flags of an uploaded model: the bits the caller loop reads (BQ2_RF_NOCASTSHADOW, BQ2_RF_DISTORT,
 * BQ2_RF_NOSELFSHADOW; M:64078-64089); 0 after upload. */
bq2_status bq2_model_set_rf_flags(bq2_ctx *ctx, int32_t model_id, uint32_t model_rf_flags);

bq2_status bq2_model_info(const bq2_ctx *ctx, int32_t model_id, int32_t *is_md3,
                          int32_t *num_frames, int32_t *num_meshes,
                          int32_t *num_verts /*[meshes]*/, int32_t *num_tris /*[meshes]*/);

/* ---- per frame ------------------------------------------------------------------------------ */
/* Host records in, one fused launch per model family, per-slot counts back.  Equivalent to
 * bq2_frame_submit + bq2_frame_wait. */
bq2_status bq2_frame(bq2_ctx *ctx, const bq2_frame_desc *desc, bq2_frame_out *out);
/* Up to SIXTEEN frames may be in flight (bq2_max_frames_in_flight): bq2_frame_submit (or bq2_world_submit) for frames
 * i+1 .. i+15 may be called before bq2_frame_wait for frame i, so the host packs the next frames while the GPU works;
 * bq2_frame_wait always completes the OLDEST submitted frame.  (A frame is a chain of dependent stream operations with
 * 50-60 us of latency end to end; a pipelined caller's frame rate is that latency / frames in flight until the host or
 * the SMs become the limit.)  The ctx holds seventeen complete frame states, taken on demand -- pinned host
 * buffers, device records, device RESULT arrays, host offset tables, stream, CUDA graph -- and a submit takes the
 * next one when a frame is in flight: the copies and kernels of consecutive frames overlap on the GPU, and what
 * bq2_frame_wait returned for frame i (device pointers, offset tables, bq2_download / bq2_expand_trisverts) stays
 * valid while the next frames are computed: if another frame was still in flight when the wait returned, until the
 * first submit after the NEXT bq2_frame_wait; if that wait emptied the pipeline, until the next submit (which reuses
 * the same state).  So a caller that never overlaps (bq2_frame / bq2_world_frame, or submit then wait) uses one
 * state only -- results valid until the next submit, no further set of arrays allocated -- and bq2_stream is fixed. */
bq2_status bq2_frame_submit(bq2_ctx *ctx, const bq2_frame_desc *desc);   /* async on bq2_stream   */
bq2_status bq2_frame_wait(bq2_ctx *ctx, bq2_frame_out *out);             /* sync + counts readback */
/* Re-run the kernels of the last submitted frame on the records already resident in HBM
 * (no host->device or device->host traffic, no sync): the "inputs resident" measurement leg. */
bq2_status bq2_frame_replay(bq2_ctx *ctx);
/* Resident frames (measurement): bq2_frame_keep copies the device records of the most recently submitted frame --
 * which must have come from bq2_frame / bq2_frame_submit and been waited for -- into an object of its own, so that a
 * set of DIFFERENT frames can be re-run back to back with every input already in HBM (bench.py cycles through more
 * of them than fit in L2).  bq2_resident_replay launches that frame's kernels on bq2_stream (async, no copies); the
 * results go to the arrays of the frame state it was computed on, which must not have been re-allocated since.
 * Replays (this one and bq2_frame_replay) are launched with programmatic stream serialization: every global write
 * keeps its stream order -- frames that share result arrays may follow each other without a sync -- while the
 * read-only front of a frame's kernel runs under the tail of the kernel in front of it. */
typedef struct bq2_resident bq2_resident;
bq2_status bq2_frame_keep(bq2_ctx *ctx, bq2_resident **out);
/* The same, with a set of RESULT arrays of the kept frame's own (a copy of the state's, same sizes): a ring of such
 * frames writes a different set of output lines on every step, so that a measurement cycling through more result
 * bytes than the L2 holds streams its outputs to HBM instead of overwriting one L2-resident set. */
bq2_status bq2_frame_keep_own(bq2_ctx *ctx, bq2_resident **out);
uint64_t   bq2_resident_result_bytes(const bq2_resident *r);   /* bytes of result arrays owned by the frame (0: shared) */
bq2_status bq2_resident_replay(bq2_ctx *ctx, const bq2_resident *r);
/* k replays launched from C, ring[(start + i) % n_ring] for i = 0 .. k-1 (all kept on one frame state).  ms != NULL:
 * two CUDA events on the stream bracket the k launches; the call waits for the second and returns their distance. */
bq2_status bq2_resident_run(bq2_ctx *ctx, bq2_resident *const *ring, int32_t n_ring, int32_t start, int32_t k, float *ms);
/* The same with the k launches dealt round-robin onto n_streams of the ctx's streams, as frames in flight run: the
 * kept frames are independent (bq2_frame_keep_own), so the tail of one launch overlaps the
 * body of the next instead of the next waiting for it (programmatic dependent launch orders them on ONE stream).  The
 * two events are on the first stream; the other streams start after the first event and are joined before the second. */
bq2_status bq2_resident_run_streams(bq2_ctx *ctx, bq2_resident *const *ring, int32_t n_ring, int32_t start, int32_t k,
                                    int32_t n_streams, float *ms);
void       bq2_resident_free(bq2_ctx *ctx, bq2_resident *r);
/* cudaProfilerStart (on != 0) / cudaProfilerStop after a device synchronize: the range ncu's range replay measures */
bq2_status bq2_profiler_range(bq2_ctx *ctx, int on);

/* ---- result download helpers (tests, the drop-in shim) --------------------------------------- */
/* Copies `count` 32-bit words starting at word offset `first` of a device result array to host. */
typedef enum bq2_array {
    BQ2_A_LERPED = 0, BQ2_A_EXTRUDED, BQ2_A_FACING_BITS, BQ2_A_INDICES, BQ2_A_INDEX_COUNT,
    BQ2_A_NORMALS, BQ2_A_TANGENTS, BQ2_A_BINORMALS, BQ2_A_LIGHTP, BQ2_A_TSH
} bq2_array;
bq2_status bq2_download(bq2_ctx *ctx, bq2_array which, uint64_t first, uint64_t count, void *host);
/* The first words_per_slot index words of pair slots [first_slot, first_slot + n_slots), packed into
 * host[n_slots][words_per_slot] by one strided copy (host memory from bq2_host_alloc makes it a DMA). */
bq2_status bq2_download_indices(bq2_ctx *ctx, int32_t first_slot, int32_t n_slots, uint32_t words_per_slot, uint32_t *host);
/* the same from the arrays a kept frame writes to (its own after bq2_frame_keep_own) */
bq2_status bq2_resident_download(bq2_ctx *ctx, const bq2_resident *r, bq2_array which, uint64_t first, uint64_t count, void *host);
const uint32_t *bq2_host_index_counts(const bq2_ctx *ctx);  /* pinned host copy, n_pair_slots      */

/* Verification at full size: one 64-bit hash per pair slot over everything the path produced for it -- the entity's
 * lerped row, the pair's extruded row, the facing words, index_count and the expanded stream pool[indices[k]] (the
 * reference's tempVertexArray, extrudedVerts, viss, TrisCounter, TrisVerts) -- computed on the device, so that every
 * pair of a 65,536-pair frame can be compared with the reference without downloading the geometry.  The definition is
 * in csrc/bq2_kernels.cu (bq2_hash_kernel); the test harness computes the same number from the engine's globals.  r == NULL: the frame bq2_frame_wait returned last; n_pair_slots must match it.  Slots that cast nothing
 * hash to 0. */
bq2_status bq2_hash_results(bq2_ctx *ctx, const bq2_resident *r, uint64_t *host_hashes, int32_t n_pair_slots);

/* Expand one pair slot into the reference's non-indexed stream: TrisVerts[3*k..] =
 * pool[indices[k]] with pool = lerped ++ extruded (M:63661-63708).  Host-side, for the shim. */
bq2_status bq2_expand_trisverts(bq2_ctx *ctx, int32_t pair_slot, float *tris_verts,
                                int32_t max_verts, int32_t *tris_counter);

/* ---- host-side C that stays on the CPU (libm trig must be the reference's) ------------------- */
/* AngleVectors (M:37566-37620). */
void bq2_host_angle_vectors(const float angles[3], float forward[3], float right[3], float up[3]);
/* M:63531-63553: move / frontv / backv of an MD2 entity from the two frame headers. */
void bq2_host_entity_md2(const void *dmdl_blob, int32_t frame, int32_t oldframe, float backlerp,
                         const float origin[3], const float oldorigin[3], const float angles[3],
                         uint32_t rf_flags, bq2_entity *out);
/* The same from the two {scale[3], translate[3]} frame headers alone (24 bytes each). */
void bq2_host_entity_md2_hdr(const float *frame_hdr, const float *oldframe_hdr, int32_t frame,
                             int32_t oldframe, float backlerp, const float origin[3],
                             const float oldorigin[3], const float angles[3], uint32_t rf_flags,
                             bq2_entity *out);
/* M:63825-63856 (also M:66802-66839): the MD3 flavour. */
void bq2_host_entity_md3(const float frame_translate[3], const float oldframe_translate[3],
                         int32_t frame, int32_t oldframe, float backlerp,
                         const float origin[3], const float oldorigin[3], const float angles[3],
                         bq2_entity *out);
/* M:63733-63745: world-space point (light origin, or r_origin per M:66689-66715) -> model space. */
void bq2_host_point_to_model(const float world[3], const float origin[3], const float angles[3],
                             float out[3]);
/* The same with the entity's AngleVectors evaluated once (basis = forward, right, up; returns 0 for the
 * identity) and reused for every light of that entity. */
int  bq2_host_entity_basis(const float angles[3], float basis[9]);
void bq2_host_point_to_model_basis(const float world[3], const float origin[3], const float *basis /* or NULL */,
                                   float out[3]);
/* Flag filter of R_DrawAliasShadowVolume M:63722-63727 (+ M:64089): 1 = entity casts a volume. */
int  bq2_host_casts_shadow(uint32_t rf_flags, int r_mirror, float r_playershadow, float r_shadows);

/* Load-time neighbour tables with the reference's exact rules, from one sorted edge list instead of the
 * reference's O(T^2) scans: MD2 findneighbour M:61782-61824 under the loader loop M:51643-51653
 * (dtriangles = dtriangle_t[T]); MD3 R_BuildTriangleNeighbors M:68646-68696.  out = int[T][3].
 * Return 0, or -1 when out of memory. */
int  bq2_host_md2_neighbours(const int16_t *dtriangles, int32_t num_tris, int32_t *out);
int  bq2_host_md3_neighbours(const uint16_t *indexes, int32_t num_tris, int32_t *out);

/* The same tables computed ON THE DEVICE (one CTA per mesh, edges bucketed by vertex in shared memory) and
 * returned to the host; bit-identical to the reference's builders.  num_verts = 1 + the largest index. */
bq2_status bq2_gpu_md2_neighbours(bq2_ctx *ctx, const int16_t *dtriangles, int32_t num_tris, int32_t *out);
bq2_status bq2_gpu_md3_neighbours(bq2_ctx *ctx, const uint16_t *indexes, int32_t num_tris, int32_t *out);

/* Load-time tangent-space bytes on the device, bit-identical to the loader loops:
 * MD2 (M:51661-51804): st = the loader's float st (fstvert_t, M:51322-51329), [num_st][2].  smooth != 0:
 *   tangents / binormals = [num_frames][num_xyz] (RF_SMOOTHVECS, incl. the coincident-vertex merge with
 *   smooth_cosine); smooth == 0: normals / tangents / binormals = [num_frames][num_tris].
 * MD3 (M:50993-51014): fills the tangent / binormal bytes of vertexes (maliasvertex_t[num_frames][num_verts])
 *   in place; st = [num_verts][2]. */
bq2_status bq2_gpu_md2_tangents(bq2_ctx *ctx, const void *dmdl_blob, size_t blob_bytes, const float *st,
                                float smooth_cosine, int smooth, uint8_t *normals, uint8_t *tangents,
                                uint8_t *binormals);
bq2_status bq2_gpu_md3_tangents(bq2_ctx *ctx, void *vertexes, int32_t num_frames, int32_t num_verts,
                                const float *st, const uint16_t *indexes, int32_t num_tris);

/* Device self-test of the kernels' `num / sqrtf(len2)` sequence (`scale / VectorLength` M:63607, `1 / VectorLength`
 * M:63620; DESIGN.md 5, Numerics) against div.rn(sqrt.rn) on the same device.  mode 0: every one of the 2^32 bit
 * patterns of len2 with numerator `num` (count ignored); mode 1: `count` random pairs with exponents across the
 * guarded ranges and their edges; mode 2: `count` raw random bit patterns.  *mismatches = results whose bits differ
 * (two NaNs compare equal). */
bq2_status bq2_selftest_ratio(bq2_ctx *ctx, int mode, float num, uint64_t count, uint64_t seed, uint64_t *mismatches);

/* ---- brush-model dynamic volumes: R_DrawBrushModelVolumes M:63326-63499 ---------------- */
/* A brush model as the renderer holds it for this function: polygon p has numverts[p] vertices (glpoly_t.verts,
 * x/y/z only) stored consecutively in verts, neighbours[slot] = polygon across edge j or -1 (glpoly_t.neighbours),
 * fence[p] != 0 = non-casting SURF_FENCE texture (M:63387-63389; NULL = none).  At most 4,096 polygons, 24,576 polygon vertices in all and 255 per polygon. */
bq2_status bq2_upload_brush(bq2_ctx *ctx, int32_t n_polys, const int32_t *numverts, const float *verts,
                            const int32_t *neighbours, const uint8_t *fence, int32_t *brush_id);
/* One (brush entity, light): light already in the entity's model space (bq2_host_point_to_model, the same lines
 * M:63341-63356), scale = 32 * radius (M:63361). */
typedef struct bq2_brush_pair { int32_t brush; float light[3]; float scale; } bq2_brush_pair;
typedef struct bq2_brush_out {
    float    *vcache;           /* DEVICE [n_pairs][vert_stride] x {vertex, extruded vertex} x 3 floats          */
    uint32_t *icache;           /* DEVICE [n_pairs][index_stride] indices into the pair's vcache rows             */
    const uint32_t *counts;     /* HOST   [n_pairs][2]: vertex pairs written (vb), indices (ib)                  */
    uint32_t vert_stride, index_stride;
} bq2_brush_out;
/* lit = for pair k the n_polys bytes of its brush, concatenated in pair order: polygon stamped by the light's
 * marking pass (lightTimestamp == r_lightTimestamp, M:63384).  Synchronous. */
bq2_status bq2_brush_volumes(bq2_ctx *ctx, const bq2_brush_pair *pairs, int32_t n_pairs, const uint8_t *lit,
                             bq2_brush_out *out);
bq2_status bq2_brush_download(bq2_ctx *ctx, int32_t pair, float *vcache /*[vb][2][3]*/, uint32_t *icache /*[ib]*/);

/* ---- the data formats either side of the path (host C, memory buffers) ------------------ */
/* Com_BlockChecksum, M:40589-40602: the four words of an MD4 digest XOR-ed. */
uint32_t   bq2_block_checksum(const void *buffer, size_t length);
/* An MD2 FILE image -> the in-memory image the engine keeps (Mod_LoadAliasModel M:51216-51345: the load-time
 * checks, frame scale/translate multiplied by `scale`; when `invert`: winding reversed, y mirrored and every
 * lightnormalindex mapped to the anorm of its y-mirrored normal, Mod_InvertNormal M:61850) and its float st
 * (st_out = float[num_st][2], may be NULL).  blob_out needs header.ofs_end bytes. */
bq2_status bq2_host_md2_from_file(const void *file, size_t file_bytes, float scale, int invert,
                                  void *blob_out, size_t blob_capacity, float *st_out);
/* The engine's MD2 precompute cache file <gamedir>/cache/<model> (write M:51811-51838, read M:51501-51606).
 * rows = num_xyz when smooth (RF_SMOOTHVECS) else num_tris.  read: *why = 0 ok, 1 "invalid cache" (flag, angle,
 * size or neighbour checksum), 2 "invalid checksum!" (a tangent-space array). */
size_t     bq2_md2_cache_bytes(int smooth, int32_t num_tris, int32_t rows, int32_t num_frames);
bq2_status bq2_md2_cache_write(void *out, size_t capacity, int smooth, float smooth_cosine,
                               const int32_t *neighbours, int32_t num_tris, int32_t rows, int32_t num_frames,
                               const uint8_t *binormals, const uint8_t *tangents, const uint8_t *normals);
bq2_status bq2_md2_cache_read(const void *in, size_t bytes, int smooth, float smooth_cosine,
                              int32_t num_tris, int32_t rows, int32_t num_frames,
                              int32_t *neighbours, uint8_t *binormals, uint8_t *tangents, uint8_t *normals,
                              int32_t *why);
/* The per-mesh MD3 cache file <gamedir>/cache/<model>_<mesh> (M:51017-51034, 50919-50947): the .md3 file's
 * checksum, then {normal, tangent, binormal} per vertex per frame, to / from maliasvertex_t rows. */
size_t     bq2_md3_cache_bytes(int32_t num_frames, int32_t num_verts);
bq2_status bq2_md3_cache_write(void *out, size_t capacity, uint32_t file_checksum, const void *vertexes,
                               int32_t num_frames, int32_t num_verts);
bq2_status bq2_md3_cache_read(const void *in, size_t bytes, uint32_t file_checksum, void *vertexes,
                              int32_t num_frames, int32_t num_verts);

/* An MD3 FILE image -> the engine's in-memory meshes (Mod_LoadAliasMD3Model M:50584-51060).  bq2_md3_file_info:
 * the loader's header checks and the mesh dimensions (num_verts / num_tris = int32[num_meshes], may be NULL).
 * bq2_host_md3_mesh_from_file, one mesh: vertexes = maliasvertex_t[num_frames * num_verts], points de-quantised
 * (short * (1/64 * scale), y mirrored when invert), normal byte from the lat/lng pair (M:50963-50989), tangent /
 * binormal bytes 0 (fill with bq2_gpu_md3_tangents or bq2_md3_cache_read); indexes = u16[3 * num_tris] (winding
 * reversed when invert); st = float[num_verts][2]; frame_translate = float[num_frames][3] or NULL. */
bq2_status bq2_md3_file_info(const void *file, size_t bytes, int32_t *num_frames, int32_t *num_meshes,
                             int32_t *num_verts, int32_t *num_tris);
bq2_status bq2_host_md3_mesh_from_file(const void *file, size_t bytes, int32_t mesh, float scale, int invert,
                                       float *frame_translate, void *vertexes, uint16_t *indexes, float *st);

/* ---- the renderer's loop nest, flattened (host, C) -------------------------------------------- */
/* The fields of entity_t (defs.h:573-618) and shadowlight_t (defs.h:3927-3973) this path reads. */
typedef struct bq2_world_entity {
    int32_t model, frame, oldframe; uint32_t rf_flags;
    float backlerp; float origin[3]; float oldorigin[3]; float angles[3];
} bq2_world_entity;                                     /* 56 bytes */
typedef struct bq2_world_light { float origin[3]; float radius; } bq2_world_light;

/* For k < n_pairs: entity pair_ent[k] is lit by light pair_light[k] (what the per-light loop
 * M:72113-72280 + R_DrawEntsShadowVolumes M:64062 enumerate).  Fills out_ents[n_ents] (lerp constants,
 * M:63531-63553 / M:63825-63856) and out_pairs (light and view_world moved to model space,
 * M:63733-63745 / M:66689-66715; scale = 32*radius M:63597), dropping pairs whose entity fails the
 * flag filter M:63722-63727.  *n_pairs_out = pairs kept.  Uses the frame headers kept at upload. */
bq2_status bq2_build_records(bq2_ctx *ctx, const bq2_world_entity *ents, int32_t n_ents,
                             const bq2_world_light *lights, int32_t n_lights,
                             const int32_t *pair_ent, const int32_t *pair_light, int32_t n_pairs,
                             const float *view_world /* [3] or NULL */,
                             bq2_entity *out_ents, bq2_pair *out_pairs, int32_t *n_pairs_out);
/* The same under explicit render options (cvars, r_mirror, pass) and the models' own flags (bq2_model_set_rf_flags):
 * a pair is kept exactly when the engine's loop would have produced a volume for it.  opts NULL = defaults, which is
 * what bq2_build_records uses. */
bq2_status bq2_build_records_opts(bq2_ctx *ctx, const bq2_world_entity *ents, int32_t n_ents,
                                  const bq2_world_light *lights, int32_t n_lights,
                                  const int32_t *pair_ent, const int32_t *pair_light, int32_t n_pairs,
                                  const float *view_world, const bq2_render_opts *opts,
                                  bq2_entity *out_ents, bq2_pair *out_pairs, int32_t *n_pairs_out);

/* ---- the same, fused, with the per-pair records built ON THE DEVICE ----------------------------------- */
/* bq2_build_records + bq2_frame in one call.  The host does only what needs libm or the model table (per
 * ENTITY: lerp constants, AngleVectors for rotated entities, the flag filter); the light and pair lists are
 * copied to HBM as they are and one small kernel builds the per-pair records there (light -> model space
 * is plain float arithmetic, bit-identical to the host's).  Differences from bq2_build_records + bq2_frame:
 *   - pair slot p IS caller pair p (a pair whose entity casts no volume keeps its slot, index_count 0);
 *   - the pair offset tables of bq2_frame_out are produced only with BQ2_WANT_TABLES (tests, the shim);
 *   - an out-of-range pair_ent / pair_light is reported by bq2_frame_wait (BQ2_E_INVALID), not up front.
 * BQ2_WANT_NTB / BQ2_WANT_TSLIGHT: when every model of the frame has per-VERTEX N/T/B rows (MD3; MD2 uploaded smooth
 * with tangent / binormal bytes) an entity's N/T/B rows are numbered like its lerped row (ent_tb_offset ==
 * ent_vert_offset) and a pair's LightP / tsH rows like its extruded row (pair_ts_offset == pair_vert_offset).
 * Frames that use a multi-mesh MD3 model, or ask for tangent-space outputs of a non-smooth MD2 (per-triangle rows) or of
 * a model without tangent / binormal bytes, take the host-built route internally, with the same slot numbering and the
 * host-built route's tables (an order of magnitude slower end to end: 154 against 16-22 us per C2 frame). */
#define BQ2_WANT_TABLES  0x10u
typedef struct bq2_world_desc {
    const bq2_world_entity *ents;    int32_t n_ents;
    const bq2_world_light  *lights;  int32_t n_lights;
    const int32_t *pair_ent, *pair_light;  int32_t n_pairs;
    const float   *view_world;       /* [3] or NULL */
    uint32_t want, index_stride;
    const bq2_render_opts *opts;     /* NULL = the engine's defaults (bq2_render_opts_default) */
} bq2_world_desc;
/* Page-locked host memory for the arrays a bq2_world_desc points at.  When ents, lights, pair_ent and pair_light of a
 * submit ALL lie inside blocks from bq2_host_alloc and together exceed 1 MiB, bq2_world_submit copies the light and
 * pair arrays to the device from where they are instead of staging them through its own pinned buffer (the entities
 * are read during the call either way: they travel inside the library's own per-entity rows; at 1M pairs the staging copy is
 * half of the host's frame time; small frames are staged either way, one copy being cheaper than five); the caller must then leave them unchanged until that frame's bq2_frame_wait has returned (an engine keeps
 * one set per frame in flight).  Any other memory works as before (borrowed only for the duration of the call).
 * NULL when out of memory.  bq2_host_free waits for the frames in flight; bq2_destroy frees what is left. */
void        *bq2_host_alloc(bq2_ctx *ctx, size_t bytes);
void         bq2_host_free(bq2_ctx *ctx, void *p);

bq2_status bq2_world_submit(bq2_ctx *ctx, const bq2_world_desc *desc);     /* async; pair with bq2_frame_wait */
bq2_status bq2_world_frame(bq2_ctx *ctx, const bq2_world_desc *desc, bq2_frame_out *out);

/* ---- multi-GPU plan (entity-major sharding) ---------------------------------------- */
/* Rank r of n owns entities [first, first+count).  Pairs follow their entity. */
void bq2_shard_entities(int32_t n_ents, int32_t rank, int32_t world, int32_t *first, int32_t *count);

#ifdef __cplusplus
}
#endif
#endif /* BQ2_GPU_H */
