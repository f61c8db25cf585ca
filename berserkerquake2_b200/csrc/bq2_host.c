/*
 * bq2_host.c -- the part of the path that stays on the host, in C.
 *
 * The reference does its trigonometry with glibc's sincosf (SinCos M:52-60) and CUDA's sincosf
 * differs in the last ulps, so everything that touches angles is done here with the same libm and
 * the same operation order as the reference lines cited, and shipped to the GPU as plain numbers
 * (move/frontv/backv per entity, model-space light per pair).  Compile WITHOUT -march / fast-math /
 * FP contraction (csrc/Makefile does) so each operation rounds once, as in the reference build.
 */
#define _GNU_SOURCE
#include "bq2_gpu.h"
#include <math.h>
#include <string.h>

#define DOT(x, y) ((x)[0]*(y)[0] + (x)[1]*(y)[1] + (x)[2]*(y)[2])   /* M:2253 */

/* AngleVectors, M:37566-37620 (PITCH = 0, YAW = 1, ROLL = 2) */
void bq2_host_angle_vectors(const float angles[3], float forward[3], float right[3], float up[3])
{
    if (angles[0] || angles[1] || angles[2]) {
        float angle, sp, sy, cp, cy, sr, cr;
        angle = angles[1] * (M_PI * 2 / 360);
        sincosf(angle, &sy, &cy);
        angle = angles[0] * (M_PI * 2 / 360);
        sincosf(angle, &sp, &cp);
        if (forward) { forward[0] = cp * cy; forward[1] = cp * sy; forward[2] = -sp; }
        if (right || up) {
            angle = angles[2] * (M_PI * 2 / 360);
            sincosf(angle, &sr, &cr);
            if (right) {
                right[0] = (-sr * sp * cy + cr * sy);
                right[1] = (-sr * sp * sy - cr * cy);
                right[2] = -sr * cp;
            }
            if (up) {
                up[0] = (cr * sp * cy + sr * sy);
                up[1] = (cr * sp * sy - sr * cy);
                up[2] = cr * cp;
            }
        }
    } else {
        if (forward) { forward[0] = 1; forward[1] = 0; forward[2] = 0; }
        if (right)   { right[0] = 0; right[1] = -1; right[2] = 0; }
        if (up)      { up[0] = 0; up[1] = 0; up[2] = 1; }
    }
}

/* delta back to the previous origin, rotated into model space: M:63534-63544 == M:63825-63841 */
static void move_delta(const float origin[3], const float oldorigin[3], const float angles[3], float move[3])
{
    move[0] = oldorigin[0] - origin[0];
    move[1] = oldorigin[1] - origin[1];
    move[2] = oldorigin[2] - origin[2];
    if (angles[0] || angles[1] || angles[2]) {
        float temp[3], f[3], r[3], u[3];
        temp[0] = move[0]; temp[1] = move[1]; temp[2] = move[2];
        bq2_host_angle_vectors(angles, f, r, u);
        move[0] = DOT(temp, f);
        move[1] = -DOT(temp, r);
        move[2] = DOT(temp, u);
    }
}

typedef struct {
    int32_t ident, version, skinwidth, skinheight, framesize;
    int32_t num_skins, num_xyz, num_st, num_tris, num_glcmds, num_frames;
    int32_t ofs_skins, ofs_st, ofs_tris, ofs_frames, ofs_glcmds, ofs_end;
} md2_header;    /* dmdl_t, defs.h:3682-3706 */

static void shell_flags(uint32_t rf_flags, bq2_entity *out)
{
    /* M:63557-63563: POWERSUIT_SCALE 1.0, POWERSUIT_SCALE_WEAPON 0.5 (defs.h:2991-2992) */
    if (rf_flags & (BQ2_RF_SHELL_RED | BQ2_RF_SHELL_GREEN | BQ2_RF_SHELL_BLUE)) {
        out->flags |= BQ2_ENT_SHELL;
        out->shell_scale = (rf_flags & BQ2_RF_WEAPONMODEL) ? 0.5 : 1.0;
    }
}

/* M:63526-63553 */
void bq2_host_entity_md2(const void *dmdl_blob, int32_t frame, int32_t oldframe, float backlerp,
                         const float origin[3], const float oldorigin[3], const float angles[3],
                         uint32_t rf_flags, bq2_entity *out)
{
    const md2_header *h = (const md2_header *)dmdl_blob;
    const float *fr  = (const float *)((const uint8_t *)dmdl_blob + h->ofs_frames + (size_t)frame * h->framesize);
    const float *ofr = (const float *)((const uint8_t *)dmdl_blob + h->ofs_frames + (size_t)oldframe * h->framesize);
    bq2_host_entity_md2_hdr(fr, ofr, frame, oldframe, backlerp, origin, oldorigin, angles, rf_flags, out);
}

/* same, from the two 24-byte {scale[3], translate[3]} frame headers */
void bq2_host_entity_md2_hdr(const float *fr, const float *ofr, int32_t frame, int32_t oldframe,
                             float backlerp, const float origin[3], const float oldorigin[3],
                             const float angles[3], uint32_t rf_flags, bq2_entity *out)
{
    float frontlerp = 1.0 - backlerp;                 /* M:63531: double subtract, float store */
    float move[3];
    int i;
    move_delta(origin, oldorigin, angles, move);
    for (i = 0; i < 3; i++) move[i] = move[i] + ofr[3 + i];                /* M:63546 */
    out->frame = frame; out->oldframe = oldframe; out->flags = 0; out->shell_scale = 0;
    out->backlerp = backlerp; out->frontlerp = frontlerp;
    for (i = 0; i < 3; i++) {                                               /* M:63548-63553 */
        out->move[i]   = backlerp * move[i] + frontlerp * fr[3 + i];
        out->frontv[i] = frontlerp * fr[i];
        out->backv[i]  = backlerp * ofr[i];
    }
    shell_flags(rf_flags, out);
}

/* M:63825-63856 */
void bq2_host_entity_md3(const float frame_translate[3], const float oldframe_translate[3],
                         int32_t frame, int32_t oldframe, float backlerp,
                         const float origin[3], const float oldorigin[3], const float angles[3],
                         bq2_entity *out)
{
    float move[3], frontlerp;
    int j;
    move_delta(origin, oldorigin, angles, move);
    frontlerp = 1.0 - backlerp;                                             /* M:63850 */
    for (j = 0; j < 3; j++) move[j] = move[j] + oldframe_translate[j];     /* M:63854 */
    for (j = 0; j < 3; j++) move[j] = backlerp * move[j] + frontlerp * frame_translate[j];
    out->frame = frame; out->oldframe = oldframe; out->flags = 0; out->shell_scale = 0;
    out->backlerp = backlerp; out->frontlerp = frontlerp;
    for (j = 0; j < 3; j++) { out->move[j] = move[j]; out->frontv[j] = frontlerp; out->backv[j] = backlerp; }
}

/* M:63733-63745 (light) == M:66689-66715 (light and r_origin) */
void bq2_host_point_to_model(const float world[3], const float origin[3], const float angles[3], float out[3])
{
    float t[3];
    t[0] = world[0] - origin[0]; t[1] = world[1] - origin[1]; t[2] = world[2] - origin[2];
    if (angles[0] || angles[1] || angles[2]) {
        float f[3], r[3], u[3];
        bq2_host_angle_vectors(angles, f, r, u);
        out[0] = DOT(t, f);
        out[1] = -DOT(t, r);
        out[2] = DOT(t, u);
    } else {
        out[0] = t[0]; out[1] = t[1]; out[2] = t[2];
    }
}

/* M:63722-63727 */
int bq2_host_casts_shadow(uint32_t rf_flags, int r_mirror, float r_playershadow, float r_shadows)
{
    if (!r_mirror && (rf_flags & BQ2_RF_VIEWERMODEL))
        if (!(r_playershadow && r_shadows > 1))
            return 0;
    if (rf_flags & (BQ2_RF_SHELL_GREEN | BQ2_RF_SHELL_RED | BQ2_RF_SHELL_BLUE | BQ2_RF_FULLBRIGHT |
                    BQ2_RF_BEAM | BQ2_RF_TRANSLUCENT | BQ2_RF_WEAPONMODEL))
        return 0;
    return 1;
}

/* entity-major sharding: contiguous, sizes differ by at most one */
void bq2_shard_entities(int32_t n_ents, int32_t rank, int32_t world, int32_t *first, int32_t *count)
{
    int32_t base, rem;
    if (world < 1) world = 1;
    if (rank < 0) rank = 0;
    if (rank >= world) rank = world - 1;
    base = n_ents / world; rem = n_ents % world;
    *first = rank * base + (rank < rem ? rank : rem);
    *count = base + (rank < rem ? 1 : 0);
}
