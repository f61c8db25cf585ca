/*
 * bq2_internal.h -- records shared by the host API (bq2_api.cu) and the kernels (bq2_kernels.cu).
 * Not part of the public ABI.
 *
 * HBM layout of one model "mesh" (an MD2 model is one mesh; an MD3 model is one per maliasmesh_t):
 *
 *   verts   MD2: frames x vstride bytes; a frame row is V x {u8 x, y, z, lightnormalindex}
 *                (dtrivertx_t, defs.h:3714-3718) padded to 16 B.  The 40-byte frame header
 *                (scale/translate/name) is NOT resident: the host folds it into move/frontv/backv.
 *           MD3: frames x V x 16 B maliasvertex_t rows (defs.h:4401-4407), already 16-byte records.
 *   tri     Tp = round_up(T, 32) records {u16 a, b, c, n0+1} (index_xyz[0..2] of dtriangle_t / WORD
 *           indexes, first neighbour) followed by Tp records {u16 n1+1, n2+1}: neighbours_t with 0 = open
 *           edge.  12 bytes per triangle instead of the reference's 12 + 12, one 16-byte-aligned block so
 *           a single TMA bulk copy stages it; a triangle's corners are one LDS.64, its other two
 *           neighbours one LDS.32.  T <= 8192 and V <= 4096 fit in 16 bits.
 *   tb/bn/nm  tangent / binormal (/ flat normal) anorm-index bytes, frames x tbstride (MD2 only;
 *           MD3 carries them inside the vertex rows).
 */
#ifndef BQ2_INTERNAL_H
#define BQ2_INTERNAL_H
#include <stdint.h>

#define BQ2_THREADS 256

#define BQ2_MESH_MD3     0x1u
#define BQ2_MESH_SMOOTH  0x2u   /* N/T/B rows are per vertex (MD3 always; MD2 with RF_SMOOTHVECS) */
#define BQ2_MESH_HAS_TB  0x4u

typedef struct bq2_dev_mesh {
    const uint8_t *verts;
    const int16_t *tri;
    const uint8_t *tb, *bn, *nm;
    int32_t  V, T, F, Tp;
    uint32_t vstride;       /* bytes per frame of verts                          */
    uint32_t tbstride;      /* bytes per frame of each tangent-space byte array  */
    uint32_t flags;
    uint32_t pad;
} bq2_dev_mesh;             /* 72 bytes */

/* one entity slot = (entity, mesh).  Self-contained: the two key-frame rows and the triangle table are
 * addressed directly, so a CTA needs ONE record load before it can start its TMA copies. */
typedef struct bq2_dev_ent {
    const uint8_t *cur, *old;   /* vertex rows of frame / oldframe                    */
    const int16_t *tri;         /* index + neighbour table                            */
    int32_t  V, T, Tp;
    uint32_t mflags;            /* BQ2_MESH_*                                         */
    int32_t  mesh;              /* index into the device mesh table (N/T/B byte rows) */
    int32_t  frame, oldframe;
    uint32_t flags;             /* BQ2_ENT_* */
    float    backlerp, frontlerp;
    float    move[3], frontv[3], backv[3];
    float    shell_scale;
    uint32_t vert_off;          /* first vertex row of this slot in lerped[]            */
    uint32_t pair_begin;        /* first entry of this slot in the sorted pair-slot list */
    uint32_t pair_count;
    uint32_t tb_off;            /* first row in normals/tangents/binormals              */
    uint32_t vstride;           /* bytes per frame row                                  */
    uint32_t nrm_off;           /* team kernel: first float4 of this slot's triangle-normal scratch row */
} bq2_dev_ent;                  /* 128 bytes */

typedef struct bq2_dev_pair {
    float    light[3];
    float    scale;
    float    view[3];
    uint32_t vert_off;      /* first vertex row in extruded[]                     */
    uint32_t bits_off;      /* first u32 word in facing_bits[]                   */
    uint32_t slot;          /* pair slot in caller order: indices + slot*stride  */
    uint32_t pad[2];        /* [0] = first row in lightp[] / tsh[] (host-built records) */
} bq2_dev_pair;             /* 48 bytes */

/* ---- device-side record building ("world" path) ------------------------------------------------
 * The host ships what it alone can decide (per ENTITY: validation, record and output-row numbering, the flag filter,
 * and for rotated entities the AngleVectors basis -- libm trig) in one row per entity that also carries the caller's
 * entity fields, plus the caller's light and pair lists as they are; ONE kernel turns that into the records above: an entity thread writes the entity record (lerp constants
 * M:63526-63553 / 63825-63856), a pair thread writes the pair record (light moved to model space, M:63733-63745) and
 * files the pair under its entity (bucket + overflow chain).  Plain mul/add in the reference's order: bit-exact. */
typedef struct bq2_dev_went {       /* one per world entity, written by the host's entity loop                  */
    int32_t  model, frame, oldframe;/* the caller's bq2_world_entity without its angles (44 bytes) ...         */
    uint32_t rf_flags;
    float    backlerp, origin[3], oldorigin[3];
    int32_t  rec;                   /* ... and what the HOST decides: index of the entity's record (warp-kernel
                                       records first)                                                          */
    int32_t  pair_rec;              /* = rec, or -1 when the entity casts no volume (M:63722-63727, 64089)   */
    uint32_t vert_off;              /* lerped row of the entity (vertices)                                    */
    uint32_t nrm_off;               /* team kernel: row of the entity in the triangle-normal scratch          */
    int32_t  basis;                 /* row of bases[] holding forward, right, up (AngleVectors M:37566: libm), or
                                       -1 when all three angles are zero                                       */
} bq2_dev_went;                     /* 64 bytes */

typedef struct bq2_dev_model {      /* device mirror of a single-mesh model, for the record-building kernel    */
    const float *hdr;               /* MD2: F x {scale[3], translate[3]}; MD3: F x translate[3]               */
    int32_t  first_mesh, md3;
} bq2_dev_model;                    /* 16 bytes */

struct bq2_world_entity;
typedef struct bq2_world_launch {
    const bq2_dev_went *went;       /* [n_ents] the host's entity rows                                      */
    const float    *bases;          /* [rotated entities][9]                                                */
    const bq2_dev_model *models;
    const bq2_dev_mesh  *meshes;
    bq2_dev_ent    *ents_out;       /* [n_ents] entity records, written by the kernel (lerp constants M:63526-63553) */
    const float    *lights;         /* [n_lights][4]: origin, radius (bq2_world_light)                     */
    const int32_t  *pair_ent, *pair_light;
    int32_t  n_pairs, n_ents, n_lights, n_recs;
    bq2_dev_pair *dpairs;           /* [n_pairs], caller order                                              */
    uint32_t *bucket_cnt;           /* [n_recs] pairs of each entity record, zeroed before the kernel       */
    uint32_t *bucket;               /* [n_recs][BQ2_BUCKET] pair indices                                    */
    uint32_t *ohead, *onext;        /* overflow chains: [n_recs] heads, [n_pairs] links                     */
    uint32_t vert_stride, bits_stride;   /* output rows of pair p start at p * stride                       */
    uint32_t *index_count;
    unsigned long long *totals;     /* [0] sum of index counts, [1] overflowing pairs, [2] bad pairs        */
    const float *view;              /* device [4]: r_origin in world space, [3] != 0 when present (M:66689)  */
} bq2_world_launch;
#define BQ2_BUCKET 8                /* pairs per entity held in its bucket; further ones chain              */

/* internal bit of bq2_launch.want, set by bq2_launch_frame: the team kernel runs with two sets of its per-light arrays */
#define BQ2_WANT_TEAM_DOUBLE 0x40000000u
typedef struct bq2_launch {
    const bq2_dev_mesh *meshes;
    const bq2_dev_ent  *ents;
    const bq2_dev_pair *pairs;
    int32_t  n_ent_slots;
    uint32_t want;
    uint32_t index_stride;
    float    *lerped, *extruded;
    uint32_t *facing_bits, *indices, *index_count;
    float    *normals, *tangents, *binormals, *lightp, *tsh;
    unsigned long long *totals;     /* world path: [0] += index count, [1] += overflowing pair; or NULL     */
    float    *tri_normals;          /* team kernel scratch: float4 {unit normal, P0.n} per triangle per slot  */
    /* device-built pair lists (world path), NULL for host-built records */
    const uint32_t *bucket_cnt, *bucket, *ohead, *onext;
    uint32_t rec_base;              /* record index of this launch's first CTA                                */
    uint32_t seq;                   /* launch sequence number (diagnostic builds: make TIMELINE=1)            */
} bq2_launch;

/* host-side flat model table for the per-entity loop in C (bq2_host.c) */
typedef struct bq2_model_row {      /* host mirror of a model for the C entity loop (V, T, flags of its first mesh) */
    int32_t F, md3, first_mesh, n_mesh, V, T, Tp; uint32_t mflags;
    /* derived by bq2_model_row_derive, so that the loop has nothing to decide per entity: */
    uint32_t vpad;                  /* lerped row length, V rounded up to 4                                    */
    int32_t  team;                  /* 0: warp-kernel mesh (V <= BQ2_WARP_MAX_V and T <= BQ2_WARP_MAX_T), 1: team kernel */
    int32_t  pad[2];                /* [0] = model_t.flags bits of the caller loop's filter (bq2_model_set_rf_flags),
                                       [1] != 0: the model's one mesh casts no shadow (its skin's flag, M:63811-63816) */
    int16_t  cmax[8];               /* the mesh's contribution to {warp max V, warp max T, team max V, team max T,
                                       warp has an MD2, team has an MD2, 0, 0}: the frame's values are the
                                       lane-wise maxima over its entities (V <= 4,096 and T <= 8,192 fit 16 bits) */
} bq2_model_row;                    /* 64 bytes */
typedef struct bq2_world_stats {
    int32_t n_warp, n_team;                 /* records [0, n_warp) -> warp kernel, [n - n_team, n) -> team kernel */
    int32_t warp_max_V, warp_max_T, warp_md2, team_max_V, team_max_T, team_md2;
    uint32_t lerped_verts_padded, max_vpad, max_bitw; int32_t max_T;
    uint32_t team_normal_rows;              /* sum of Tp over team-kernel records */
    int32_t  n_rotated;                     /* rows of bases[] written */
    uint64_t total_verts;
    uint32_t and_mflags;                    /* AND of the mesh flags of every entity's model (BQ2_MESH_*; all ones for no entity) */
} bq2_world_stats;

#ifdef __cplusplus
extern "C" {
#endif
struct bq2_world_entity;
/* One pass over the world entities: validation, kernel class, record / output-row numbering, flag filter
 * (M:63722-63727) and -- the one thing that needs the host's libm -- AngleVectors of rotated entities.  The lerp
 * constants themselves (M:63526-63553 / 63825-63856) are computed by the record-building kernel from the caller's
 * entities and these rows.  Returns 0, -1 unknown model, -2 frame out of range, -3 a model is not single-mesh,
 * -4 the frame's rows do not fit 32-bit row numbers. */
void bq2_model_row_derive(bq2_model_row *row);
void bq2_host_world_struct_sizes(int32_t out[3]);   /* sizeof model row, entity row, stats: for harnesses that mirror them */
struct bq2_render_opts;
/* bases[] as bq2_host_world_entities leaves it holds each rotated entity's ANGLES in the first three floats of its row;
   this turns the rows into forward / right / up (libm).  Run by the submit worker, or inline. */
void bq2_host_bases_from_angles(float *bases, int32_t n_rot);
void bq2_host_filter_masks(const struct bq2_render_opts *opts, int *off, uint32_t *ent_mask, uint32_t *model_mask, uint32_t *model_xor, int *brush);
int  bq2_host_world_entities(const struct bq2_world_entity *ents, int32_t n, const bq2_model_row *models, int32_t n_models,
                             const struct bq2_render_opts *opts, bq2_dev_went *rows, float *bases /* [n][9] */, int32_t *rec_of_ent, uint32_t *ent_vert_off,
                             bq2_world_stats *st);
/* Two kernels share the records.  "warp" kernel: one warp per (entity, light) pair, one CTA per entity,
 * for meshes with V <= BQ2_WARP_MAX_V and T <= BQ2_WARP_MAX_T (an MD2 player model is ~500 triangles).
 * "team" kernel: the whole CTA works on one pair at a time; any mesh size, and the N/T/B + tangent-space
 * light outputs.  Both return cudaError_t as int; max_V / max_T size the dynamic shared memory. */
#define BQ2_WARP_MAX_V   1024
#define BQ2_WARP_MAX_T   1024
#define BQ2_WARP_WARPS   4
/* overlap: launch with programmatic stream serialization (the kernel's prologue may run under the tail of the kernel
   in front of it on the stream); never inside a graph capture */
int  bq2_launch_frame(const bq2_launch *L, int max_V, int max_T, int any_md2, void *stream, int overlap);       /* team */
int  bq2_launch_frame_warp(const bq2_launch *L, int max_V, int max_T, int any_md2, void *stream, int overlap);  /* warp */
int  bq2_launch_world_setup(const bq2_world_launch *W, void *stream);   /* returns cudaError_t; one kernel */
int  bq2_world_setup_launches(const bq2_world_launch *W);
/* words [copy_from, n_copy) of the counts + totals to mapped host memory; words [keep_from, n_all) saved to keep[]; words
   [zero_from, n_all) cleared for the next frame (copy_from <= zero_from <= keep_from) */
int  bq2_launch_frame_tail(uint32_t *cnt, uint32_t *host, uint32_t *keep, uint32_t copy_from, uint32_t n_copy, uint32_t zero_from,
                           uint32_t keep_from, uint32_t n_all, void *stream, int overlap);
/* neighbour tables of one mesh on the device: idx = u16 vertex indices, in_stride u16s per triangle (6 for
   dtriangle_t, 3 for WORD indexes); out = int32[3*T] (device) */
int  bq2_launch_selftest_ratio(int mode, float num, uint64_t count, uint64_t seed, unsigned long long *mismatches, void *stream);
int  bq2_launch_neighbours(const uint16_t *idx, int in_stride, int T, int V, int md3, int32_t *out, void *stream);
size_t bq2_neighbours_smem_bytes(int T, int V);
/* tangent-space bytes of one mesh, one CTA per key-frame; device pointers; mode 0 MD2 smooth, 1 MD2 flat, 2 MD3 */
int  bq2_launch_tangents(const uint16_t *idx_xyz, int xyz_stride, const uint16_t *idx_st, int st_stride,
                         const float *st, int num_st, const uint8_t *verts, size_t frame_stride, int num_frames,
                         int T, int V, int mode, float smooth_cosine,
                         uint8_t *out_n, uint8_t *out_t, uint8_t *out_b, size_t out_stride, void *stream);
size_t bq2_tangents_smem_bytes(int T, int V, int num_st, int mode);
/* brush-model volumes (8f.4): device mirrors of the records the kernel reads */
typedef struct bq2_brush_dev_h { const void *slot_info; const int32_t *nbr; const void *verts; int32_t n_polys, n_slots; } bq2_brush_dev_h;
typedef struct bq2_brush_pair_dev_h { int32_t brush; float light[3]; float scale; uint32_t lit_off; } bq2_brush_pair_dev_h;
#define BQ2_BRUSH_MAX_POLYS 4096        /* 1 byte per polygon + 2.75 per polygon vertex of shared memory         */
#define BQ2_BRUSH_MAX_SLOTS 24576       /* slot numbers and running sums are kept in 16 bits                     */
int  bq2_launch_brush(const void *brushes, const void *pairs, const uint8_t *lit, float *vcache, uint32_t *icache,
                      uint32_t *counts, uint32_t vert_stride, uint32_t index_stride, int n_pairs, int max_polys, int max_slots,
                      void *stream);
/* one 64-bit hash per pair slot over the frame's results (verification; out = device [n_pair_slots], zeroed by the caller) */
int  bq2_launch_hash(const bq2_launch *L, unsigned long long *out, void *stream);
int  bq2_kernel_prepare(void);   /* opt in to large dynamic shared memory once per process */
size_t bq2_kernel_smem_bytes(int max_V, int max_T, int md2);
size_t bq2_kernel_warp_smem_bytes(int max_V, int max_T, int md2);
#ifdef __cplusplus
}
#endif
#endif
