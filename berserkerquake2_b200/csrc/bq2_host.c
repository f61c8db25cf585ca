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
#include <stddef.h>

#define DOT(x, y) ((x)[0]*(y)[0] + (x)[1]*(y)[1] + (x)[2]*(y)[2])   /* M:2253 */

/* sincosf of +-0 is (+-0, 1) exactly; most entities turn about one axis only, so the libm call is skipped there */
static inline void sincos_angle(float angle, float *s, float *c)
{
    if (angle == 0.0f) { *s = angle; *c = 1.0f; }
    else sincosf(angle, s, c);
}

/* AngleVectors, M:37566-37620 (PITCH = 0, YAW = 1, ROLL = 2) */
void bq2_host_angle_vectors(const float angles[3], float forward[3], float right[3], float up[3])
{
    if (angles[0] || angles[1] || angles[2]) {
        float angle, sp, sy, cp, cy, sr, cr;
        angle = angles[1] * (M_PI * 2 / 360);
        sincos_angle(angle, &sy, &cy);
        angle = angles[0] * (M_PI * 2 / 360);
        sincos_angle(angle, &sp, &cp);
        if (forward) { forward[0] = cp * cy; forward[1] = cp * sy; forward[2] = -sp; }
        if (right || up) {
            angle = angles[2] * (M_PI * 2 / 360);
            sincos_angle(angle, &sr, &cr);
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

static void entity_md2_finish(const float *fr, const float *ofr, int32_t frame, int32_t oldframe, float backlerp,
                              float move[3], uint32_t rf_flags, bq2_entity *out);
static void entity_md3_finish(const float frame_translate[3], const float oldframe_translate[3], int32_t frame,
                              int32_t oldframe, float backlerp, float move[3], bq2_entity *out);

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
    float move[3];
    move_delta(origin, oldorigin, angles, move);
    entity_md2_finish(fr, ofr, frame, oldframe, backlerp, move, rf_flags, out);
}

static void entity_md2_finish(const float *fr, const float *ofr, int32_t frame, int32_t oldframe, float backlerp,
                              float move[3], uint32_t rf_flags, bq2_entity *out)
{
    float frontlerp = 1.0 - backlerp;                 /* M:63531: double subtract, float store */
    int i;
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
    float move[3];
    move_delta(origin, oldorigin, angles, move);
    entity_md3_finish(frame_translate, oldframe_translate, frame, oldframe, backlerp, move, out);
}

static void entity_md3_finish(const float frame_translate[3], const float oldframe_translate[3], int32_t frame,
                              int32_t oldframe, float backlerp, float move[3], bq2_entity *out)
{
    float frontlerp;
    int j;
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

/* The same transform with the entity's AngleVectors evaluated once and reused for every light that hits the
 * entity (the reference re-evaluates them per light; same inputs, same libm, same results).
 * basis = forward[3], right[3], up[3]; returns 0 when all three angles are zero (identity, M:63737). */
int bq2_host_entity_basis(const float angles[3], float basis[9])
{
    if (!(angles[0] || angles[1] || angles[2])) return 0;
    bq2_host_angle_vectors(angles, basis, basis + 3, basis + 6);
    return 1;
}
void bq2_host_point_to_model_basis(const float world[3], const float origin[3], const float *basis, float out[3])
{
    float t[3];
    t[0] = world[0] - origin[0]; t[1] = world[1] - origin[1]; t[2] = world[2] - origin[2];
    if (basis) {
        out[0] = DOT(t, basis);
        out[1] = -DOT(t, basis + 3);
        out[2] = DOT(t, basis + 6);
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

/* ---- the caller loops' switches (M:64041-64133 + M:63722-63727), reduced to three masks so that the per-entity loop
 *      below tests an entity with two ANDs ------------------------------------------------------------------------ */
void bq2_render_opts_default(bq2_render_opts *o)
{
    o->r_mirror = 0; o->r_mirror_shadows = 2.0f; o->r_mirror_models = 1.0f; o->r_playershadow = 0.0f;
    o->r_shadows = 2.0f; o->r_drawentities = 1.0f; o->pass = BQ2_PASS_ALL;
}
/* off: nothing alias casts under these options; ent_mask: entity flags that reject; an alias model is rejected when
   ((model flags ^ model_xor) & model_mask) != 0; brush: the loop draws brush models in this pass */
void bq2_host_filter_masks(const bq2_render_opts *opts, int *off, uint32_t *ent_mask, uint32_t *model_mask, uint32_t *model_xor, int *brush)
{
    bq2_render_opts d;
    const bq2_render_opts *o = opts;
    int alias_on, loop_on, nomdl;
    if (!o) { bq2_render_opts_default(&d); o = &d; }
    /* M:64046-64060: the loop runs at all */
    if (o->r_mirror) { loop_on = o->r_drawentities != 0.0f && o->r_mirror_shadows != 0.0f; nomdl = !(o->r_mirror_models != 0.0f); }
    else             { loop_on = o->r_drawentities != 0.0f && o->r_shadows != 0.0f; nomdl = 0; }
    /* M:64072-64079: alias volumes only with r_shadows (r_mirror_shadows in the mirror) > 1, and not with nomdl */
    alias_on = loop_on && !nomdl && (o->r_mirror ? o->r_mirror_shadows > 1.0f : o->r_shadows > 1.0f);
    *off = !alias_on;
    *brush = loop_on && (o->pass == BQ2_PASS_ALL || o->pass == BQ2_PASS_SELF);                    /* M:64068-64070 */
    /* M:64089 (entity side) + M:63725-63726 */
    *ent_mask = BQ2_RF_NOCASTSHADOW | BQ2_RF_DISTORT | BQ2_RF_SHELL_GREEN | BQ2_RF_SHELL_RED | BQ2_RF_SHELL_BLUE |
                BQ2_RF_FULLBRIGHT | BQ2_RF_BEAM | BQ2_RF_TRANSLUCENT | BQ2_RF_WEAPONMODEL;
    /* M:63722-63724: the player's own model casts only in the mirror, or with r_playershadow and r_shadows > 1 */
    if (!o->r_mirror && !(o->r_playershadow != 0.0f && o->r_shadows > 1.0f)) *ent_mask |= BQ2_RF_VIEWERMODEL;
    /* M:64089 (model side) and the self-shadow split M:64081-64087 */
    *model_mask = BQ2_RF_NOCASTSHADOW | BQ2_RF_DISTORT; *model_xor = 0u;
    if (o->pass == BQ2_PASS_SELF) *model_mask |= BQ2_RF_NOSELFSHADOW;
    else if (o->pass == BQ2_PASS_NOSELF) { *model_mask |= BQ2_RF_NOSELFSHADOW; *model_xor = BQ2_RF_NOSELFSHADOW; }
}
int bq2_host_entity_casts(uint32_t ent_rf_flags, uint32_t model_rf_flags, int model_kind, const bq2_render_opts *opts)
{
    int off, brush; uint32_t em, mm, mx;
    bq2_host_filter_masks(opts, &off, &em, &mm, &mx, &brush);
    if (model_kind == 1) return brush;
    return !(off || (ent_rf_flags & em) || ((model_rf_flags ^ mx) & mm));
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

/* ---- the per-entity half of the world path (bq2_world_submit), one pass, plain C ---------------------- */
#include "bq2_internal.h"
/* the flag filter of the shadow pass (bq2_host_filter_masks above: M:63722-63727, 64041-64089) arrives as masks */

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

/* the caller's entity and the first 44 bytes of its device row are the same fields */
typedef char bq2_went_layout_check[(offsetof(bq2_dev_went, rec) == offsetof(bq2_world_entity, angles) &&
                                    offsetof(bq2_dev_went, origin) == offsetof(bq2_world_entity, origin) &&
                                    offsetof(bq2_dev_went, oldorigin) == offsetof(bq2_world_entity, oldorigin) &&
                                    sizeof(bq2_dev_went) == 64 && sizeof(bq2_model_row) == 64) ? 1 : -1];

void bq2_model_row_derive(bq2_model_row *r)
{
    const int md2 = !(r->mflags & BQ2_MESH_MD3);
    memset(r->cmax, 0, sizeof r->cmax); r->pad[0] = r->pad[1] = 0;
    r->vpad = (uint32_t)(r->V + 3) & ~3u;
    r->team = !(r->V <= BQ2_WARP_MAX_V && r->T <= BQ2_WARP_MAX_T);
    r->cmax[2 * r->team] = (int16_t)r->V; r->cmax[2 * r->team + 1] = (int16_t)r->T; r->cmax[4 + r->team] = (int16_t)md2;
}

int bq2_host_world_entities(const bq2_world_entity *ents, int32_t n, const bq2_model_row *models, int32_t n_models,
                            const bq2_render_opts *opts,
                            bq2_dev_went *rows, float *bases, int32_t *rec_of_ent, uint32_t *ent_vert_off, bq2_world_stats *st)
{
    int f_off, f_brush; uint32_t f_em, f_mm, f_mx;
    /* the running numbers live in locals (the stores through rows / rec_of_ent could alias *st); the six per-class
       maxima are one 8 x int16 vector (bq2_model_row.cmax), what is a function of the largest mesh (row lengths) is
       derived after the loop */
    int32_t n_warp = 0, n_team = 0, n_rot = 0;
    uint64_t voff = 0, team_rows = 0;          /* stored as 32-bit row numbers; checked after the loop */
    uint64_t total_verts = 0;
    uint32_t and_flags = 0xffffffffu;
    int16_t m[8];
    int32_t e;
#if defined(__SSE2__)
    __m128i mx = _mm_setzero_si128();
#else
    int k;
    memset(m, 0, sizeof m);
#endif
    memset(st, 0, sizeof *st);
    bq2_host_filter_masks(opts, &f_off, &f_em, &f_mm, &f_mx, &f_brush);
    if (f_off) f_em = 0xffffffffu, f_mm = 0xffffffffu, f_mx = 0xffffffffu;      /* nothing casts: every entity is rejected below */
    for (e = 0; e < n; e++) {
        const bq2_world_entity *W = &ents[e];
        const bq2_model_row *mo;
        bq2_dev_went *x = &rows[e];
        uint32_t nrm_off = 0;
        int32_t rec;
        if ((uint32_t)W->model >= (uint32_t)n_models) return -1;
        mo = &models[W->model];
        if (mo->n_mesh != 1) return -3;
        and_flags &= mo->mflags;
        if ((uint32_t)W->frame >= (uint32_t)mo->F || (uint32_t)W->oldframe >= (uint32_t)mo->F) return -2;
#if defined(__SSE2__)
        mx = _mm_max_epi16(mx, _mm_loadu_si128((const __m128i *)mo->cmax));
#else
        for (k = 0; k < 6; k++) if (mo->cmax[k] > m[k]) m[k] = mo->cmax[k];
#endif
        if (!mo->team) rec = n_warp++;                              /* warp-kernel records grow up from 0 ... */
        else {                                                      /* ... team-kernel records down from the end */
            rec = n - 1 - n_team++;
            nrm_off = (uint32_t)team_rows; team_rows += (uint32_t)mo->Tp;
        }
        rec_of_ent[e] = rec;
        /* the caller's entity, without its angles (the device gets them as a basis): the first 44 bytes of both structs */
        memcpy(x, W, offsetof(bq2_world_entity, angles));
        /* AngleVectors once per entity: the reference evaluates them again for the move delta and for every light
           (same inputs, same libm, same values) */
        if (W->angles[0] || W->angles[1] || W->angles[2]) {
            /* the row holds the ANGLES until bq2_host_bases_from_angles turns it into forward / right / up: the trigonometry
               (libm, ~20 ns per sincosf) can then run on the ctx's submit worker instead of the caller's thread */
            float *b = bases + 9 * (size_t)n_rot;
            b[0] = W->angles[0]; b[1] = W->angles[1]; b[2] = W->angles[2];
            x->basis = n_rot++;
        } else
            x->basis = -1;
        x->rec = rec;
        x->pair_rec = (f_off | (W->rf_flags & f_em) | (((uint32_t)mo->pad[0] ^ f_mx) & f_mm) | (uint32_t)mo->pad[1]) ? -1 : rec;   /* M:63722-63727, 64041-64089; pad[1]: the model's one mesh casts nothing */
        x->vert_off = (uint32_t)voff; x->nrm_off = nrm_off;
        ent_vert_off[e] = (uint32_t)voff;
        voff += mo->vpad; total_verts += (uint64_t)mo->V;
    }
    if ((voff | team_rows) >> 32) return -4;
    ent_vert_off[n] = (uint32_t)voff;
#if defined(__SSE2__)
    _mm_storeu_si128((__m128i *)m, mx);
#endif
    st->n_warp = n_warp; st->n_team = n_team;
    st->warp_max_V = m[0]; st->warp_max_T = m[1]; st->team_max_V = m[2]; st->team_max_T = m[3];
    st->warp_md2 = m[4]; st->team_md2 = m[5];
    {
        const int32_t mV = m[0] > m[2] ? m[0] : m[2], mT = m[1] > m[3] ? m[1] : m[3];
        st->max_vpad = (uint32_t)(mV + 3) & ~3u; st->max_bitw = (uint32_t)(mT + 31) >> 5; st->max_T = mT;
    }
    st->lerped_verts_padded = (uint32_t)voff; st->team_normal_rows = (uint32_t)team_rows; st->n_rotated = n_rot;
    st->total_verts = total_verts; st->and_mflags = and_flags;
    return 0;
}

/* sizes of the three structs the entity loop exchanges with its caller, for harnesses that mirror them (tests/test_host_world.py) */
void bq2_host_world_struct_sizes(int32_t out[3])
{ out[0] = (int32_t)sizeof(bq2_model_row); out[1] = (int32_t)sizeof(bq2_dev_went); out[2] = (int32_t)sizeof(bq2_world_stats); }

/* rows of bases[] written by the loop above: angles -> forward, right, up (AngleVectors M:37566, the engine's libm) */
void bq2_host_bases_from_angles(float *bases, int32_t n_rot)
{
    int32_t r;
    for (r = 0; r < n_rot; r++) {
        float *b = bases + 9 * (size_t)r;
        const float ang[3] = { b[0], b[1], b[2] };
        bq2_host_angle_vectors(ang, b, b + 3, b + 6);
    }
}

/* ---- load-time neighbour tables, host C ------------------------------------- */
/* Both reference builders scan every triangle for every edge (O(T^2): M:61782-61824, M:68646-68676).
 * These produce the same tables from one sorted edge list; every order-dependent rule of the reference
 * (first match wins, a second match poisons the edge, MD2 writes the back-pointer of the match, MD3 counts
 * each triangle once with the forward test taking precedence) is kept. */
#include <stdlib.h>

typedef struct { uint32_t key, slot; } edge_rec;        /* slot = 3*tri + edge */
static int edge_cmp(const void *a, const void *b)
{
    const edge_rec *x = (const edge_rec *)a, *y = (const edge_rec *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->slot < y->slot ? -1 : (x->slot > y->slot);
}
static size_t edge_lower_bound(const edge_rec *e, size_t n, uint32_t key)
{
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t mid = (lo + hi) >> 1; if (e[mid].key < key) lo = mid + 1; else hi = mid; }
    return lo;
}

/* MD2: findneighbour (M:61782-61824) driven by the loader loop M:51643-51653.  dtriangles = dtriangle_t[T]
 * (index_xyz[3], index_st[3]).  Returns 0, or -1 when out of memory. */
int bq2_host_md2_neighbours(const int16_t *dtriangles, int32_t num_tris, int32_t *out)
{
    size_t n = 3 * (size_t)num_tris, s;
    edge_rec *e = (edge_rec *)malloc(sizeof(edge_rec) * (n ? n : 1));
    if (!e) return -1;
    for (s = 0; s < n; s++) {
        size_t t = s / 3, j = s % 3;
        uint32_t a = (uint16_t)dtriangles[6 * t + j], b = (uint16_t)dtriangles[6 * t + (j + 1) % 3];
        e[s].key = a < b ? (a << 16 | b) : (b << 16 | a);      /* the match is undirected */
        e[s].slot = (uint32_t)s;
        out[s] = -1;
    }
    qsort(e, n, sizeof *e, edge_cmp);
    for (s = 0; s < n; s++) {
        size_t t = s / 3, j = s % 3, k;
        uint32_t a, b, key;
        int32_t found = -1, foundj = 0, dup = 0;
        if (out[s] != -1) continue;                             /* "none found yet", M:51652 */
        a = (uint16_t)dtriangles[6 * t + j]; b = (uint16_t)dtriangles[6 * t + (j + 1) % 3];
        key = a < b ? (a << 16 | b) : (b << 16 | a);
        for (k = edge_lower_bound(e, n, key); k < n && e[k].key == key; k++) {
            uint32_t ot = e[k].slot / 3;
            if (ot == t) continue;                              /* i == index is skipped whole */
            if (found == -1) { found = (int32_t)ot; foundj = (int32_t)(e[k].slot % 3); }
            else { dup = 1; break; }
        }
        if (!dup && found != -1) { out[3 * found + foundj] = (int32_t)t; out[s] = found; }
    }
    free(e);
    return 0;
}

/* MD3: R_BuildTriangleNeighbors / R_FindTriangleWithEdge (M:68646-68696). */
int bq2_host_md3_neighbours(const uint16_t *indexes, int32_t num_tris, int32_t *out)
{
    size_t n = 3 * (size_t)num_tris, s;
    edge_rec *e = (edge_rec *)malloc(sizeof(edge_rec) * (n ? n : 1));
    if (!e) return -1;
    for (s = 0; s < n; s++) {
        size_t t = s / 3, j = s % 3;
        e[s].key = (uint32_t)indexes[3 * t + j] << 16 | indexes[3 * t + (j + 1) % 3];   /* directed */
        e[s].slot = (uint32_t)s;
    }
    qsort(e, n, sizeof *e, edge_cmp);
    for (s = 0; s < n; s++) {
        size_t t = s / 3, j = s % 3, k;
        /* edge j of the triangle runs idx[j] -> idx[j+1]; the query is its reverse (M:68691-68693) */
        uint32_t start = indexes[3 * t + (j + 1) % 3], end = indexes[3 * t + j];
        uint32_t fkey = start << 16 | end, bkey = end << 16 | start, last = 0xffffffffu;
        int32_t match = -1, count = 0;
        for (k = edge_lower_bound(e, n, fkey); k < n && e[k].key == fkey; k++) {
            uint32_t ot = e[k].slot / 3;
            if (ot == last) continue;                           /* one triangle counts once */
            last = ot;
            if (ot != t) match = (int32_t)ot;                   /* ascending: the last one wins */
            count++;
        }
        last = 0xffffffffu;
        for (k = edge_lower_bound(e, n, bkey); k < n && e[k].key == bkey; k++) {
            uint32_t ot = e[k].slot / 3;
            const uint16_t *ix = indexes + 3 * ot;
            if (ot == last) continue;
            last = ot;
            if ((ix[0] == start && ix[1] == end) || (ix[1] == start && ix[2] == end) || (ix[2] == start && ix[0] == end))
                continue;                                       /* already counted by the forward branch */
            count++;
        }
        if (count > 2) match = -1;                              /* edge shared by three triangles: a seam */
        out[s] = match;
    }
    free(e);
    return 0;
}
