/*
 * bq2_oracle.h -- CPU restatement of the reference's alias lerp / shadow-volume arithmetic.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load liboracle.so.  The shipped library (libbq2gpu.so)
 * never links or calls anything declared here.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against the reference's own
 * object code (oracle/_ref/libbq2ref.so = unmodified main.c compiled by oracle/ref/build_ref.sh)
 * in tests/test_oracle_vs_reference.py, and against the committed golden vectors that the same
 * reference build produced (tests/golden/, made by tools/make_golden.py).  The reference ships no
 * tests or fixtures of its own, so execution of its code is the only pin.
 *
 * Conventions: "M:" = reference main.c line numbers.  Arithmetic is IEEE-754 binary32, one
 * rounding per operation, no FMA contraction (compile with -ffp-contract=off, no -march), Sqrt is
 * sqrtf (the reference's start-up default, M:79665-79669).
 */
#ifndef BQ2_ORACLE_H
#define BQ2_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const float *bq2o_anorms(void);                       /* [162][3]  data.h:2234-2239            */

/* M:37566-37620 */
void bq2o_angle_vectors(const float angles[3], float fwd[3], float right[3], float up[3]);
/* M:63733-63745 (same lines as M:66689-66715 for r_origin) */
void bq2o_point_to_model(const float world[3], const float origin[3], const float angles[3], float out[3]);

/* M:63526-63553.  blob = in-memory dmdl_t image. */
void bq2o_md2_lerp_consts(const void *blob, int frame, int oldframe, float backlerp,
                          const float origin[3], const float oldorigin[3], const float angles[3],
                          float move[3], float frontv[3], float backv[3], float *frontlerp);
/* M:63555-63582.  shell!=0 selects the M:63565-63572 branch. */
void bq2o_md2_lerp(const void *blob, int frame, int oldframe,
                   const float move[3], const float frontv[3], const float backv[3],
                   int shell, float shell_scale, float *lerped /*[V][3]*/);

/* M:63825-63856 */
void bq2o_md3_lerp_consts(const float frame_translate[3], const float oldframe_translate[3],
                          float backlerp, const float origin[3], const float oldorigin[3],
                          const float angles[3], float move[3], float *frontlerp);
/* M:63885-63898.  verts/oldverts: maliasvertex_t rows (16 B each). */
void bq2o_md3_lerp_extrude(const void *verts, const void *oldverts, int num_verts,
                           const float move[3], float backlerp, float frontlerp,
                           const float light[3], float scale,
                           float *lerped, float *extruded /* may be NULL */);

/* M:63597-63630 (MD2: per-corner extrusion, only referenced vertices written) and M:63901-63916.
 * idx = int32[T][3]. extruded may be NULL (MD3 already extruded per vertex). */
void bq2o_facing(const float *lerped, const int32_t *idx, int num_tris,
                 const float light[3], float scale,
                 float *extruded, float *trinormals /* may be NULL */, uint8_t *viss);

/* M:63642-63709 / M:63919-63991.  Emits the reference's expanded stream and the canonical index
 * form: pool = lerped[0..V) ++ extruded[0..V).  Either output may be NULL.
 * Returns TrisCounter. */
int bq2o_shadow_volume(const float *lerped, const float *extruded, int num_verts,
                       const int32_t *idx, const int32_t *neighbours, const uint8_t *viss,
                       int num_tris, float *tris_verts, uint32_t *indices);

/* a10: lerped N/T/B.  MD2 smooth M:65213-65237 (per vertex), MD2 flat M:65239-65260 (per tri),
 * MD3 M:65710-65732 (per vertex).  rows = V or T.  nidx etc. are byte rows of the two frames. */
void bq2o_ntb(const uint8_t *n_cur, const uint8_t *n_old, int n_stride,
              const uint8_t *t_cur, const uint8_t *t_old, int t_stride,
              const uint8_t *b_cur, const uint8_t *b_old, int b_stride,
              int rows, float backlerp, float frontlerp,
              float *normals, float *tangents, float *binormals);
/* a11/a12: LightP = (ld.T, ld.B, ld.N), tsH = LightP + (H.T, H.B, H.N)   M:64952-65003, 65888-65903 */
void bq2o_tslight(const float *lerped, const float *normals, const float *tangents,
                  const float *binormals, int rows, const float light[3], const float view[3],
                  float *lightp, float *tsh);

/* CPU-baseline loop over world-space pairs, one model per pair (bench.py cpu_baseline "port" kind). */
long long bq2o_md2_world_bench(void *const *blobs, void *const *neighbours, int n_pairs,
                               const int32_t *frames, const int32_t *oldframes, const float *backlerps,
                               const float *origins, const float *oldorigins, const float *angles,
                               const float *lights_world, const float *radii);

/* M:63326-63499 (brush-model volumes); light in model space, scale = 32 * radius */
int bq2o_brush_volume(int n_polys, const int32_t *numverts, const float *verts, const int32_t *neighbours,
                      const uint8_t *lit, const uint8_t *fence, const float light[3], float scale,
                      float *vcache, uint32_t *icache, int *vb_out);

/* precompute (SURVEY 8f "next" rows 1-2) */
uint8_t bq2o_normal2index(const float v[3]);                                  /* M:25206-25223 */
void bq2o_md2_neighbours(const int16_t *tris /*dtriangle_t[T]*/, int num_tris, int32_t *out); /* M:51643-51653 + 61782-61824 */
void bq2o_md3_neighbours(const uint16_t *indexes, int num_tris, int32_t *out);               /* M:68646-68696 */
void bq2o_vecs_for_tris(const float *v0, const float *v1, const float *v2,
                        const float *st0, const float *st1, const float *st2,
                        float tangent[3], float binormal[3]);                  /* M:61719-61753 */
/* M:51661-51746: smooth per-vertex tangent/binormal bytes for every frame.  st = fstvert_t[num_st]
 * as the loader leaves them (M:51322-51329). */
void bq2o_md2_tangents_smooth(const void *blob, const float *st, float smooth_cosine,
                              uint8_t *tangents, uint8_t *binormals);
/* M:51747-51804: flat per-triangle normal/tangent/binormal bytes. */
void bq2o_md2_tangents_flat(const void *blob, const float *st,
                            uint8_t *normals, uint8_t *tangents, uint8_t *binormals);
/* M:50956-51015 minus the lat/lng normal decode (normals supplied by the caller):
 * per-frame per-vertex tangent/binormal bytes written into the maliasvertex_t rows. */
void bq2o_md3_tangents(void *vertexes, int num_frames, int num_verts, const float *st /*[V][2]*/,
                       const uint16_t *indexes, int num_tris);

#ifdef __cplusplus
}
#endif
#endif
