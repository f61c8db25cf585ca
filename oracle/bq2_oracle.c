/*
 * bq2_oracle.c -- CPU restatement of the reference's alias lerp / shadow-volume arithmetic.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT (see bq2_oracle.h).  Parity status: PINNED against the
 * reference's own object code and the golden vectors it produced (tests/test_oracle_vs_reference.py,
 * tests/test_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).  No -march, so
 * like the reference build (Makefile:3) every float operation is a single correctly rounded SSE op.
 *
 * "M:" = reference main.c line numbers.  Operation ORDER below is deliberate: it is what makes the
 * results bit-identical to the reference, and is called out where it matters.
 */
#define _GNU_SOURCE
#include "bq2_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

/* ---- anorms (data.h:2234-2239) ----------------------------------------------------------------- */
static const uint32_t anorm_bits[162 * 4] = {
#include "../include/bq2_anorms.inc"
};
static float anorm_tab[162][3];
static int   anorm_ready;
static void anorm_init(void)
{
    if (anorm_ready) return;
    for (int i = 0; i < 162; i++)
        for (int k = 0; k < 3; k++)
            memcpy(&anorm_tab[i][k], &anorm_bits[i * 4 + k], 4);
    anorm_ready = 1;
}
const float *bq2o_anorms(void) { anorm_init(); return &anorm_tab[0][0]; }

/* ---- MD2 in-memory image (defs.h:3682-3726) ---------------------------------------------------- */
typedef struct {
    int32_t ident, version, skinwidth, skinheight, framesize;
    int32_t num_skins, num_xyz, num_st, num_tris, num_glcmds, num_frames;
    int32_t ofs_skins, ofs_st, ofs_tris, ofs_frames, ofs_glcmds, ofs_end;
} o_md2_header;
typedef struct { int16_t index_xyz[3], index_st[3]; } o_md2_tri;
typedef struct { uint8_t v[3], lightnormalindex; } o_md2_vert;
typedef struct { float scale[3], translate[3]; char name[16]; o_md2_vert verts[1]; } o_md2_frame;
/* maliasvertex_t, defs.h:4401-4407 */
typedef struct { float point[3]; uint8_t normal, tangent, binormal, pad; } o_md3_vert;

static const o_md2_frame *md2_frame(const void *blob, int f)
{
    const o_md2_header *h = (const o_md2_header *)blob;
    return (const o_md2_frame *)((const uint8_t *)blob + h->ofs_frames + (size_t)f * h->framesize);
}

/* ---- helpers: M:2253, M:38107-38119, M:38129-38144, M:38610-38615 ---------------------------- */
static inline float dot3(const float *x, const float *y) { return x[0]*y[0] + x[1]*y[1] + x[2]*y[2]; }
static inline float vlen(const float *v)
{   /* length = 0; length += v[i]*v[i] (i=0..2); Sqrt(length) */
    float length = 0;
    for (int i = 0; i < 3; i++) length += v[i] * v[i];
    return sqrtf(length);
}
static inline float vnormalize(float *v)
{
    float length = sqrtf(dot3(v, v));
    if (length) { float il = 1 / length; v[0] *= il; v[1] *= il; v[2] *= il; }
    return length;
}
static inline void cross3(const float *a, const float *b, float *c)
{
    c[0] = a[1]*b[2] - a[2]*b[1];
    c[1] = a[2]*b[0] - a[0]*b[2];
    c[2] = a[0]*b[1] - a[1]*b[0];
}

/* ---- AngleVectors M:37566-37620 ----------------------------------------------------------------- */
void bq2o_angle_vectors(const float angles[3], float fwd[3], float right[3], float up[3])
{
    if (angles[0] || angles[1] || angles[2]) {
        float sp, sy, cp, cy, sr, cr, angle;
        angle = angles[1] * (3.14159265358979323846 * 2 / 360);   /* YAW; double multiply, float store */
        sincosf(angle, &sy, &cy);
        angle = angles[0] * (3.14159265358979323846 * 2 / 360);   /* PITCH */
        sincosf(angle, &sp, &cp);
        if (fwd) { fwd[0] = cp * cy; fwd[1] = cp * sy; fwd[2] = -sp; }
        if (right || up) {
            angle = angles[2] * (3.14159265358979323846 * 2 / 360);   /* ROLL */
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
        if (fwd)   { fwd[0] = 1; fwd[1] = 0;  fwd[2] = 0; }
        if (right) { right[0] = 0; right[1] = -1; right[2] = 0; }
        if (up)    { up[0] = 0; up[1] = 0; up[2] = 1; }
    }
}

/* M:63733-63745 */
void bq2o_point_to_model(const float world[3], const float origin[3], const float angles[3], float out[3])
{
    float t[3] = { world[0] - origin[0], world[1] - origin[1], world[2] - origin[2] };
    if (angles[0] || angles[1] || angles[2]) {
        float f[3], r[3], u[3];
        bq2o_angle_vectors(angles, f, r, u);
        out[0] = dot3(t, f);
        out[1] = -dot3(t, r);
        out[2] = dot3(t, u);
    } else {
        out[0] = t[0]; out[1] = t[1]; out[2] = t[2];
    }
}

/* M:63526-63553 */
void bq2o_md2_lerp_consts(const void *blob, int frame, int oldframe, float backlerp,
                          const float origin[3], const float oldorigin[3], const float angles[3],
                          float move[3], float frontv[3], float backv[3], float *frontlerp_out)
{
    const o_md2_frame *fr = md2_frame(blob, frame), *ofr = md2_frame(blob, oldframe);
    float frontlerp = 1.0 - backlerp;                         /* double subtract, float store */
    move[0] = oldorigin[0] - origin[0];
    move[1] = oldorigin[1] - origin[1];
    move[2] = oldorigin[2] - origin[2];
    if (angles[0] || angles[1] || angles[2]) {
        float temp[3] = { move[0], move[1], move[2] }, v0[3], v1[3], v2[3];
        bq2o_angle_vectors(angles, v0, v1, v2);
        move[0] = dot3(temp, v0);
        move[1] = -dot3(temp, v1);
        move[2] = dot3(temp, v2);
    }
    for (int i = 0; i < 3; i++) move[i] = move[i] + ofr->translate[i];
    for (int i = 0; i < 3; i++) {
        move[i]   = backlerp * move[i] + frontlerp * fr->translate[i];
        frontv[i] = frontlerp * fr->scale[i];
        backv[i]  = backlerp * ofr->scale[i];
    }
    if (frontlerp_out) *frontlerp_out = frontlerp;
}

/* M:63555-63582: lerp = ((move + ov*backv) + v*frontv) [+ normal*scale] */
void bq2o_md2_lerp(const void *blob, int frame, int oldframe,
                   const float move[3], const float frontv[3], const float backv[3],
                   int shell, float shell_scale, float *lerped)
{
    const o_md2_header *h = (const o_md2_header *)blob;
    const o_md2_vert *v = md2_frame(blob, frame)->verts, *ov = md2_frame(blob, oldframe)->verts;
    anorm_init();
    for (int i = 0; i < h->num_xyz; i++) {
        float *lerp = lerped + 3 * i;
        if (shell) {
            const float *normal = anorm_tab[v[i].lightnormalindex];   /* current frame only, M:63567 */
            for (int k = 0; k < 3; k++)
                lerp[k] = move[k] + ov[i].v[k] * backv[k] + v[i].v[k] * frontv[k] + normal[k] * shell_scale;
        } else {
            for (int k = 0; k < 3; k++)
                lerp[k] = move[k] + ov[i].v[k] * backv[k] + v[i].v[k] * frontv[k];
        }
    }
}

/* M:63825-63856 */
void bq2o_md3_lerp_consts(const float ft[3], const float oft[3], float backlerp,
                          const float origin[3], const float oldorigin[3], const float angles[3],
                          float move[3], float *frontlerp_out)
{
    float frontlerp;
    move[0] = oldorigin[0] - origin[0];
    move[1] = oldorigin[1] - origin[1];
    move[2] = oldorigin[2] - origin[2];
    if (angles[0] || angles[1] || angles[2]) {
        float temp[3] = { move[0], move[1], move[2] }, f[3], r[3], u[3];
        bq2o_angle_vectors(angles, f, r, u);
        move[0] = dot3(temp, f);
        move[1] = -dot3(temp, r);
        move[2] = dot3(temp, u);
    }
    frontlerp = 1.0 - backlerp;
    for (int j = 0; j < 3; j++) move[j] = move[j] + oft[j];
    for (int j = 0; j < 3; j++) move[j] = backlerp * move[j] + frontlerp * ft[j];
    if (frontlerp_out) *frontlerp_out = frontlerp;
}

/* extrusion of one point: M:63606-63610 == M:63893-63897 */
static inline void extrude_point(const float *p, const float *light, float scale, float *e)
{
    float v1[3] = { p[0] - light[0], p[1] - light[1], p[2] - light[2] };
    float sca = scale / vlen(v1);
    e[0] = v1[0] * sca + p[0];
    e[1] = v1[1] * sca + p[1];
    e[2] = v1[2] * sca + p[2];
}

/* M:63885-63898 */
void bq2o_md3_lerp_extrude(const void *verts, const void *oldverts, int num_verts,
                           const float move[3], float backlerp, float frontlerp,
                           const float light[3], float scale, float *lerped, float *extruded)
{
    const o_md3_vert *v = (const o_md3_vert *)verts, *ov = (const o_md3_vert *)oldverts;
    for (int j = 0; j < num_verts; j++) {
        float *p = lerped + 3 * j;
        for (int k = 0; k < 3; k++)
            p[k] = move[k] + ov[j].point[k] * backlerp + v[j].point[k] * frontlerp;
        if (extruded) extrude_point(p, light, scale, extruded + 3 * j);
    }
}

/* M:63599-63630 / M:63901-63916 */
void bq2o_facing(const float *lerped, const int32_t *idx, int num_tris,
                 const float light[3], float scale,
                 float *extruded, float *trinormals, uint8_t *viss)
{
    for (int i = 0; i < num_tris; i++) {
        const float *tri[3];
        float v1[3], v2[3], normal[3], tn[3];
        for (int j = 0; j < 3; j++) {
            int ix = idx[3 * i + j];
            tri[j] = lerped + 3 * ix;
            if (extruded) extrude_point(tri[j], light, scale, extruded + 3 * ix);
        }
        for (int k = 0; k < 3; k++) { v1[k] = tri[0][k] - tri[1][k]; v2[k] = tri[2][k] - tri[1][k]; }
        cross3(v2, v1, normal);
        {   /* VectorScale(normal, 1/VectorLength(normal), trinormal): one divide, three multiplies */
            float inv = 1 / vlen(normal);
            tn[0] = normal[0] * inv; tn[1] = normal[1] * inv; tn[2] = normal[2] * inv;
        }
        if (trinormals) { trinormals[3*i] = tn[0]; trinormals[3*i+1] = tn[1]; trinormals[3*i+2] = tn[2]; }
        /* NaN (degenerate triangle) makes the <= false, hence viss = 1: keep this compare direction */
        if (dot3(tn, light) <= dot3(tri[0], tn)) viss[i] = 0; else viss[i] = 1;
    }
}

/* M:63642-63709 / M:63919-63991 */
int bq2o_shadow_volume(const float *lerped, const float *extruded, int num_verts,
                       const int32_t *idx, const int32_t *neighbours, const uint8_t *viss,
                       int num_tris, float *tris_verts, uint32_t *indices)
{
    int counter = 0;
    uint32_t V = (uint32_t)num_verts;
#define EMIT(ix, ext) do { \
        if (tris_verts) { const float *s_ = ((ext) ? extruded : lerped) + 3 * (ix); \
            tris_verts[3*counter] = s_[0]; tris_verts[3*counter+1] = s_[1]; tris_verts[3*counter+2] = s_[2]; } \
        if (indices) indices[counter] = (uint32_t)(ix) + ((ext) ? V : 0u); \
        counter++; } while (0)
    for (int i = 0; i < num_tris; i++) {
        if (!viss[i]) continue;
        for (int j = 0; j < 3; j++) {
            int nb = neighbours[3 * i + j];
            int shadow = (nb == -1) ? 1 : !viss[nb];
            if (!shadow) continue;
            int i0 = idx[3 * i + j], i1 = idx[3 * i + (j + 1) % 3];
            EMIT(i0, 0); EMIT(i0, 1); EMIT(i1, 1);
            EMIT(i1, 1); EMIT(i1, 0); EMIT(i0, 0);
        }
    }
    for (int i = 0; i < num_tris; i++) {
        if (!viss[i]) continue;
        for (int j = 0; j < 3; j++)  EMIT(idx[3 * i + j], 0);
        for (int j = 2; j >= 0; j--) EMIT(idx[3 * i + j], 1);
    }
#undef EMIT
    return counter;
}

/* CPU-baseline loop ("port" kind, used only when the compiled reference is unavailable): the restated
 * R_DrawAliasShadowVolume per world-space pair, writing the same arrays the reference writes. */
long long bq2o_md2_world_bench(void *const *blobs, void *const *neighbours, int n_pairs,
                               const int32_t *frames, const int32_t *oldframes, const float *backlerps,
                               const float *origins, const float *oldorigins, const float *angles,
                               const float *lights_world, const float *radii)
{
    static float lerped[3 * 4096], extruded[3 * 4096], trin[3 * 8192], tris_verts[36 * 8192];
    static uint8_t viss[8192];
    static int32_t idx[3 * 4096];
    long long total = 0;
    for (int p = 0; p < n_pairs; p++) {
        const o_md2_header *h = (const o_md2_header *)blobs[p];
        const o_md2_tri *tris = (const o_md2_tri *)((const uint8_t *)blobs[p] + h->ofs_tris);
        float move[3], frontv[3], backv[3], fl, light[3];
        bq2o_md2_lerp_consts(blobs[p], frames[p], oldframes[p], backlerps[p], origins + 3 * p, oldorigins + 3 * p,
                             angles + 3 * p, move, frontv, backv, &fl);
        bq2o_point_to_model(lights_world + 3 * p, origins + 3 * p, angles + 3 * p, light);
        bq2o_md2_lerp(blobs[p], frames[p], oldframes[p], move, frontv, backv, 0, 0.0f, lerped);
        for (int t = 0; t < h->num_tris; t++) for (int k = 0; k < 3; k++) idx[3 * t + k] = tris[t].index_xyz[k];
        bq2o_facing(lerped, idx, h->num_tris, light, 32 * radii[p], extruded, trin, viss);
        total += bq2o_shadow_volume(lerped, extruded, h->num_xyz, idx, (const int32_t *)neighbours[p], viss,
                                    h->num_tris, tris_verts, NULL);
    }
    return total;
}

/* R_DrawBrushModelVolumes M:63326-63499 on flat arrays: returns ib; *vb_out = vertex pairs.
 * vcache = [2*vb][3]: vertex then its extrusion; icache = indices into it. */
int bq2o_brush_volume(int n_polys, const int32_t *numverts, const float *verts, const int32_t *neighbours,
                      const uint8_t *lit, const uint8_t *fence, const float light[3], float scale,
                      float *vcache, uint32_t *icache, int *vb_out)
{
    int vb = 0, ib = 0, first = 0, surf_base = 0;
    for (int p = 0; p < n_polys; first += numverts[p], p++) {          /* vertex buffer, M:63384-63403 */
        if (!lit[p] || fence[p]) continue;
        for (int j = 0; j < numverts[p]; j++) {
            const float *v = verts + 3 * (first + j);
            float *o = vcache + 6 * vb;
            o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
            extrude_point(v, light, scale, o + 3);
            vb++;
        }
    }
    first = 0;
    for (int p = 0; p < n_polys; first += numverts[p], p++) {          /* index buffer, M:63406-63460 */
        int n = numverts[p];
        if (!lit[p] || fence[p]) continue;
        for (int j = 0; j < n; j++) {
            int nb = neighbours[first + j];
            int shadow = nb < 0 ? 1 : (lit[nb] != lit[p]);               /* neighbour stamped by another light pass */
            if (shadow) {
                int jj = (j + 1) % n;
                icache[ib++] = j * 2 + 0 + surf_base;  icache[ib++] = j * 2 + 1 + surf_base;  icache[ib++] = jj * 2 + 1 + surf_base;
                icache[ib++] = j * 2 + 0 + surf_base;  icache[ib++] = jj * 2 + 1 + surf_base; icache[ib++] = jj * 2 + 0 + surf_base;
            }
        }
        for (int j = 0; j < n - 2; j++) { icache[ib++] = 0 + surf_base; icache[ib++] = (j + 1) * 2 + 0 + surf_base; icache[ib++] = (j + 2) * 2 + 0 + surf_base; }
        for (int j = 0; j < n - 2; j++) { icache[ib++] = 1 + surf_base; icache[ib++] = (j + 2) * 2 + 1 + surf_base; icache[ib++] = (j + 1) * 2 + 1 + surf_base; }
        surf_base += n * 2;
    }
    *vb_out = vb;
    return ib;
}

/* a10.  MD2 writes old*backlerp + cur*frontlerp (M:65220), MD3 cur*frontlerp + old*backlerp
 * (M:65714); IEEE addition commutes, the two are bit-identical. */
void bq2o_ntb(const uint8_t *n_cur, const uint8_t *n_old, int n_stride,
              const uint8_t *t_cur, const uint8_t *t_old, int t_stride,
              const uint8_t *b_cur, const uint8_t *b_old, int b_stride,
              int rows, float backlerp, float frontlerp,
              float *normals, float *tangents, float *binormals)
{
    anorm_init();
    for (int i = 0; i < rows; i++) {
        for (int k = 0; k < 3; k++) {
            if (normals)   normals[3*i+k]   = anorm_tab[n_old[i*n_stride]][k] * backlerp + anorm_tab[n_cur[i*n_stride]][k] * frontlerp;
            if (tangents)  tangents[3*i+k]  = anorm_tab[t_old[i*t_stride]][k] * backlerp + anorm_tab[t_cur[i*t_stride]][k] * frontlerp;
            if (binormals) binormals[3*i+k] = anorm_tab[b_old[i*b_stride]][k] * backlerp + anorm_tab[b_cur[i*b_stride]][k] * frontlerp;
        }
    }
}

/* a11/a12: M:64952-65003 */
void bq2o_tslight(const float *lerped, const float *normals, const float *tangents,
                  const float *binormals, int rows, const float light[3], const float view[3],
                  float *lightp, float *tsh)
{
    for (int i = 0; i < rows; i++) {
        const float *p = lerped + 3*i, *n = normals + 3*i, *t = tangents + 3*i, *b = binormals + 3*i;
        float ld[3] = { light[0] - p[0], light[1] - p[1], light[2] - p[2] };
        float lp[3] = { dot3(ld, t), dot3(ld, b), dot3(ld, n) };
        lightp[3*i] = lp[0]; lightp[3*i+1] = lp[1]; lightp[3*i+2] = lp[2];
        if (tsh && view) {
            float H[3] = { view[0] - p[0], view[1] - p[1], view[2] - p[2] };
            float h[3] = { dot3(H, t), dot3(H, b), dot3(H, n) };
            tsh[3*i] = lp[0] + h[0]; tsh[3*i+1] = lp[1] + h[1]; tsh[3*i+2] = lp[2] + h[2];
        }
    }
}

/* ---- precompute --------------------------------------------------------------------------------- */
/* M:25206-25223: first strict maximum wins; bestd starts at 0 so an all-nonpositive vector gives 0 */
uint8_t bq2o_normal2index(const float v[3])
{
    int best = 0; float bestd = 0;
    anorm_init();
    for (int i = 0; i < 162; i++) {
        float d = dot3(v, anorm_tab[i]);
        if (d > bestd) { bestd = d; best = i; }
    }
    return (uint8_t)best;
}

/* M:61782-61824 */
static int md2_findneighbour(int index, int edgei, int numtris, int32_t *nb, const o_md2_tri *tris)
{
    int found = -1, foundj = 0, dup = 0;
    const o_md2_tri *cur = &tris[index];
    for (int i = 0; i < numtris; i++) {
        if (i == index) continue;
        for (int j = 0; j < 3; j++) {
            if ((cur->index_xyz[edgei] == tris[i].index_xyz[j] &&
                 cur->index_xyz[(edgei + 1) % 3] == tris[i].index_xyz[(j + 1) % 3]) ||
                (cur->index_xyz[edgei] == tris[i].index_xyz[(j + 1) % 3] &&
                 cur->index_xyz[(edgei + 1) % 3] == tris[i].index_xyz[j])) {
                if (found == -1) { found = i; foundj = j; }
                else dup = 1;
            }
        }
    }
    if (!dup && found != -1) { nb[3 * found + foundj] = index; return found; }
    return -1;
}
/* M:51643-51653 */
void bq2o_md2_neighbours(const int16_t *tris_raw, int num_tris, int32_t *out)
{
    const o_md2_tri *tris = (const o_md2_tri *)tris_raw;
    for (int i = 0; i < 3 * num_tris; i++) out[i] = -1;
    for (int i = 0; i < num_tris; i++)
        for (int j = 0; j < 3; j++)
            if (out[3 * i + j] == -1)
                out[3 * i + j] = md2_findneighbour(i, j, num_tris, out, tris);
}

/* M:68646-68676 */
static int md3_find_edge(const uint16_t *indexes, int numtris, uint16_t start, uint16_t end, int ignore)
{
    int match = -1, count = 0;
    for (int i = 0; i < numtris; i++, indexes += 3) {
        if ((indexes[0] == start && indexes[1] == end) || (indexes[1] == start && indexes[2] == end) ||
            (indexes[2] == start && indexes[0] == end)) {
            if (i != ignore) match = i;
            count++;
        } else if ((indexes[1] == start && indexes[0] == end) || (indexes[2] == start && indexes[1] == end) ||
                   (indexes[0] == start && indexes[2] == end)) {
            count++;
        }
    }
    if (count > 2) match = -1;
    return match;
}
/* M:68684-68696 */
void bq2o_md3_neighbours(const uint16_t *indexes, int num_tris, int32_t *out)
{
    const uint16_t *index = indexes;
    for (int i = 0; i < num_tris; i++, index += 3) {
        out[3*i+0] = md3_find_edge(indexes, num_tris, index[1], index[0], i);
        out[3*i+1] = md3_find_edge(indexes, num_tris, index[2], index[1], i);
        out[3*i+2] = md3_find_edge(indexes, num_tris, index[0], index[2], i);
    }
}

/* M:61719-61753 */
void bq2o_vecs_for_tris(const float *v0, const float *v1, const float *v2,
                        const float *st0, const float *st1, const float *st2,
                        float Tangent[3], float Binormal[3])
{
    float vec1[3], vec2[3], planes[3][3], tmp;
    for (int i = 0; i < 3; i++) {
        vec1[0] = v1[i] - v0[i]; vec1[1] = st1[0] - st0[0]; vec1[2] = st1[1] - st0[1];
        vec2[0] = v2[i] - v0[i]; vec2[1] = st2[0] - st0[0]; vec2[2] = st2[1] - st0[1];
        vnormalize(vec1); vnormalize(vec2);
        cross3(vec1, vec2, planes[i]);
    }
    for (int i = 0; i < 3; i++) {
        tmp = 1.0 / planes[i][0];                    /* double divide, float store (M:61746) */
        Tangent[i]  = -planes[i][1] * tmp;
        Binormal[i] = -planes[i][2] * tmp;
    }
    vnormalize(Tangent); vnormalize(Binormal);
}

/* M:68755-68795 (MD3 flavour divides directly instead of multiplying by a reciprocal) */
static void md3_tangent_for_tri(const uint16_t *index, const o_md3_vert *vertices, const float *st,
                                float Tangent[3], float Binormal[3])
{
    const float *v0 = vertices[index[0]].point, *v1 = vertices[index[1]].point, *v2 = vertices[index[2]].point;
    const float *st0 = st + 2 * index[0], *st1 = st + 2 * index[1], *st2 = st + 2 * index[2];
    float vec1[3], vec2[3], planes[3][3];
    for (int i = 0; i < 3; i++) {
        vec1[0] = v1[i] - v0[i]; vec1[1] = st1[0] - st0[0]; vec1[2] = st1[1] - st0[1];
        vec2[0] = v2[i] - v0[i]; vec2[1] = st2[0] - st0[0]; vec2[2] = st2[1] - st0[1];
        vnormalize(vec1); vnormalize(vec2);
        cross3(vec1, vec2, planes[i]);
    }
    for (int i = 0; i < 3; i++) {
        Tangent[i]  = -planes[i][1] / planes[i][0];
        Binormal[i] = -planes[i][2] / planes[i][0];
    }
    vnormalize(Tangent); vnormalize(Binormal);
}

/* M:51661-51746 */
void bq2o_md2_tangents_smooth(const void *blob, const float *st, float smooth_cosine,
                              uint8_t *tangents, uint8_t *binormals)
{
    const o_md2_header *h = (const o_md2_header *)blob;
    const o_md2_tri *tris = (const o_md2_tri *)((const uint8_t *)blob + h->ofs_tris);
    int V = h->num_xyz, T = h->num_tris;
    float (*tan_)[3] = malloc(sizeof(float[3]) * V), (*bin_)[3] = malloc(sizeof(float[3]) * V);
    anorm_init();
    for (int i = 0; i < h->num_frames; i++) {
        const o_md2_vert *verts = md2_frame(blob, i)->verts;
        memset(tan_, 0, sizeof(float[3]) * V); memset(bin_, 0, sizeof(float[3]) * V);
        for (int j = 0; j < T; j++) {
            float vv[3][3], tangent[3], binormal[3];
            for (int c = 0; c < 3; c++) for (int k = 0; k < 3; k++)
                vv[c][k] = (float)verts[tris[j].index_xyz[c]].v[k];
            bq2o_vecs_for_tris(vv[0], vv[1], vv[2], st + 2 * tris[j].index_st[0],
                               st + 2 * tris[j].index_st[1], st + 2 * tris[j].index_st[2], tangent, binormal);
            for (int k = 0; k < 3; k++) {
                int l = tris[j].index_xyz[k];
                for (int c = 0; c < 3; c++) { tan_[l][c] = tan_[l][c] + tangent[c]; bin_[l][c] = bin_[l][c] + binormal[c]; }
            }
        }
        /* coincident-vertex merge, M:51713-51728 */
        for (int j = 0; j < V; j++)
            for (int k = j + 1; k < V; k++)
                if (verts[j].v[0] == verts[k].v[0] && verts[j].v[1] == verts[k].v[1] && verts[j].v[2] == verts[k].v[2]) {
                    const float *jn = anorm_tab[verts[j].lightnormalindex], *kn = anorm_tab[verts[k].lightnormalindex];
                    if (dot3(jn, kn) >= smooth_cosine) {
                        for (int c = 0; c < 3; c++) { tan_[j][c] = tan_[j][c] + tan_[k][c]; tan_[k][c] = tan_[j][c]; }
                        for (int c = 0; c < 3; c++) { bin_[j][c] = bin_[j][c] + bin_[k][c]; bin_[k][c] = bin_[j][c]; }
                    }
                }
        for (int j = 0; j < V; j++) {
            vnormalize(tan_[j]); vnormalize(bin_[j]);
            tangents[i * V + j]  = bq2o_normal2index(tan_[j]);
            binormals[i * V + j] = bq2o_normal2index(bin_[j]);
        }
    }
    free(tan_); free(bin_);
}

/* M:51747-51804 */
void bq2o_md2_tangents_flat(const void *blob, const float *st,
                            uint8_t *normals, uint8_t *tangents, uint8_t *binormals)
{
    const o_md2_header *h = (const o_md2_header *)blob;
    const o_md2_tri *tris = (const o_md2_tri *)((const uint8_t *)blob + h->ofs_tris);
    int T = h->num_tris;
    for (int i = 0; i < h->num_frames; i++) {
        const o_md2_vert *verts = md2_frame(blob, i)->verts;
        for (int j = 0; j < T; j++) {
            float tri[3][3], v1[3], v2[3], normal[3], tan[3], bin[3];
            for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++)
                tri[k][l] = verts[tris[j].index_xyz[k]].v[l];
            for (int k = 0; k < 3; k++) { v1[k] = tri[0][k] - tri[1][k]; v2[k] = tri[2][k] - tri[1][k]; }
            cross3(v2, v1, normal);
            {   /* VectorScale(normal, -1.0/VectorLength(normal), normal): scale is a DOUBLE -> float arg */
                float s = -1.0 / vlen(normal);
                normal[0] *= s; normal[1] *= s; normal[2] *= s;
            }
            normals[i * T + j] = bq2o_normal2index(normal);
            bq2o_vecs_for_tris(tri[0], tri[1], tri[2], st + 2 * tris[j].index_st[0],
                               st + 2 * tris[j].index_st[1], st + 2 * tris[j].index_st[2], tan, bin);
            tangents[i * T + j]  = bq2o_normal2index(tan);
            binormals[i * T + j] = bq2o_normal2index(bin);
        }
    }
}

/* M:50956-51015 (tangent/binormal part) */
void bq2o_md3_tangents(void *vertexes, int num_frames, int num_verts, const float *st,
                       const uint16_t *indexes, int num_tris)
{
    float (*tan_)[3] = malloc(sizeof(float[3]) * num_verts), (*bin_)[3] = malloc(sizeof(float[3]) * num_verts);
    for (int l = 0; l < num_frames; l++) {
        o_md3_vert *pv = (o_md3_vert *)vertexes + (size_t)l * num_verts;
        memset(tan_, 0, sizeof(float[3]) * num_verts); memset(bin_, 0, sizeof(float[3]) * num_verts);
        for (int j = 0; j < num_tris; j++) {
            float tangent[3], binormal[3];
            md3_tangent_for_tri(indexes + 3 * j, pv, st, tangent, binormal);
            for (int k = 0; k < 3; k++) {
                int ll = indexes[3 * j + k];
                for (int c = 0; c < 3; c++) { tan_[ll][c] = tan_[ll][c] + tangent[c]; bin_[ll][c] = bin_[ll][c] + binormal[c]; }
            }
        }
        for (int j = 0; j < num_verts; j++) {
            vnormalize(tan_[j]); vnormalize(bin_[j]);
            pv[j].tangent  = bq2o_normal2index(tan_[j]);
            pv[j].binormal = bq2o_normal2index(bin_[j]);
        }
    }
    free(tan_); free(bin_);
}
