/*
 * bq2_synth.h -- procedural MD2 / MD3 models for parity tests and benchmarks.
 *
 * Quake II pak data is not available offline, so inputs are generated: closed genus-0 "blob"
 * meshes in the exact MD2 in-memory layout the reference holds after load (defs.h:3682-3726) and
 * MD3 meshes in the reference's in-memory layout (maliasvertex_t 16 B, defs.h:4401-4407).
 * Deterministic: the same arguments give the same bytes on every machine with this libm.
 * Part of the host library (plain C); the oracle and the CUDA path consume the same bytes.
 */
#ifndef BQ2_SYNTH_H
#define BQ2_SYNTH_H
#include <stddef.h>
#include <stdint.h>
#include "bq2_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif

/* sphere(S, R): S slices, R rings -> V = S*(R-1)+2, T = 2*S*(R-1).  S=16,R=17: V=258,T=512. */
int32_t bq2_synth_sphere_verts(int32_t slices, int32_t rings);
int32_t bq2_synth_sphere_tris(int32_t slices, int32_t rings);

/* Bytes of the MD2 image (header + st + tris + frames). */
size_t  bq2_synth_md2_bytes(int32_t slices, int32_t rings, int32_t num_frames);
/* Fill `blob` (bq2_synth_md2_bytes big).  Per frame f: scale = 0.25*(1 +- 0.05u), translate =
 * -32*(1 +- 0.05u); vertex byte = clamp(128 + rho*dir), rho = 100 + 20*sin(3*phi + 0.1f + phase)
 * *sin(theta); lightnormalindex = nearest anorm (first strict maximum, as Normal2Index M:25206).
 * `phase_seed` makes models distinct.  Returns 0, or -1 on bad arguments. */
int     bq2_synth_md2(int32_t slices, int32_t rings, int32_t num_frames, uint32_t phase_seed, void *blob);
/* Float st coordinates as the reference loader leaves them: (s-0.5)/skinwidth (M:51322-51329). */
void    bq2_synth_md2_st(const void *blob, float *st /*[num_st][2]*/);

/* One MD3 mesh: a sphere centred at `centre`, radius ~24 units, points quantised to 1/64 as the
 * MD3 loader de-quantises them (M:51047).  vertexes = maliasvertex_t[num_frames*V] (normal byte
 * filled, tangent/binormal bytes left 0), indexes = u16[3*T], st = float[V][2]. */
int     bq2_synth_md3_mesh(int32_t slices, int32_t rings, int32_t num_frames, uint32_t phase_seed,
                           const float centre[3], void *vertexes, uint16_t *indexes, float *st);

/* Nearest-anorm lookup used by the generators (restates Normal2Index M:25206-25223). */
uint8_t bq2_host_normal2index(const float v[3]);
const float *bq2_host_anorms(void);    /* [162][4] x,y,z,0 */

/* Scene generator (SURVEY 8d "Entities"/"Lights"): LCG x <- 1664525x + 1013904223. */
/* Writes entity_t-style world entities and lights_per_ent world-space light origins per entity. */
void bq2_synth_scene(int32_t n_ents, int32_t lights_per_ent, int32_t num_frames, int32_t n_models,
                     uint32_t seed, uint32_t step,
                     bq2_world_entity *ents, float *lights_world /*[n_ents*lights_per_ent][3]*/);

#ifdef __cplusplus
}
#endif
#endif
