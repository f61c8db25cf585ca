/*
 * bq2_synth.c -- procedural MD2 / MD3 inputs (see include/bq2_synth.h).
 * Plain C host code.  Layouts follow the reference's in-memory structs: dmdl_t / daliasframe_t /
 * dtrivertx_t / dtriangle_t (defs.h:3682-3726) and maliasvertex_t (defs.h:4401-4407).
 */
#include "bq2_synth.h"
#include <math.h>
#include <stdio.h>
#include <string.h>

#define PI_D 3.14159265358979323846

static const uint32_t anorm_bits[162 * 4] = {
#include "bq2_anorms.inc"
};
const float *bq2_host_anorms(void) { return (const float *)(const void *)anorm_bits; }

/* Normal2Index, M:25206-25223: first strict maximum over the 162 anorms, starting from 0. */
uint8_t bq2_host_normal2index(const float v[3])
{
    const float *a = bq2_host_anorms();
    int best = 0; float bestd = 0;
    for (int i = 0; i < 162; i++) {
        float d = v[0] * a[4*i] + v[1] * a[4*i+1] + v[2] * a[4*i+2];
        if (d > bestd) { bestd = d; best = i; }
    }
    return (uint8_t)best;
}

typedef struct {
    int32_t ident, version, skinwidth, skinheight, framesize;
    int32_t num_skins, num_xyz, num_st, num_tris, num_glcmds, num_frames;
    int32_t ofs_skins, ofs_st, ofs_tris, ofs_frames, ofs_glcmds, ofs_end;
} md2_header;

int32_t bq2_synth_sphere_verts(int32_t S, int32_t R) { return S * (R - 1) + 2; }
int32_t bq2_synth_sphere_tris(int32_t S, int32_t R)  { return 2 * S * (R - 1); }

size_t bq2_synth_md2_bytes(int32_t S, int32_t R, int32_t F)
{
    size_t V = (size_t)bq2_synth_sphere_verts(S, R), T = (size_t)bq2_synth_sphere_tris(S, R);
    return sizeof(md2_header) + V * 4 /*st*/ + T * 12 + (size_t)F * (40 + 4 * V);
}

static uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
static float unit_pm(uint32_t h) { return (float)((int32_t)(hash32(h) & 0xffff) - 32768) / 32768.0f; }

/* sphere topology shared by MD2 and MD3: CCW seen from outside */
static void sphere_tri(int32_t S, int32_t R, int32_t t, int32_t out[3])
{
    int32_t V = S * (R - 1) + 2, south = V - 1;
    if (t < S) {                                   /* north fan */
        int32_t s = t;
        out[0] = 0; out[1] = 1 + s; out[2] = 1 + (s + 1) % S;
    } else if (t < S + 2 * S * (R - 2)) {          /* bands */
        int32_t q = (t - S) / 2, half = (t - S) & 1, r = q / S, s = q % S;
        int32_t a = 1 + r * S + s, b = 1 + r * S + (s + 1) % S;
        int32_t c = 1 + (r + 1) * S + s, d = 1 + (r + 1) * S + (s + 1) % S;
        if (!half) { out[0] = a; out[1] = c; out[2] = d; }
        else       { out[0] = a; out[1] = d; out[2] = b; }
    } else {                                       /* south fan */
        int32_t s = t - (S + 2 * S * (R - 2)), base = 1 + (R - 2) * S;
        out[0] = south; out[1] = base + (s + 1) % S; out[2] = base + s;
    }
}
static void sphere_dir(int32_t S, int32_t R, int32_t v, double *theta, double *phi)
{
    int32_t V = S * (R - 1) + 2;
    if (v == 0)          { *theta = 0.0; *phi = 0.0; }
    else if (v == V - 1) { *theta = PI_D; *phi = 0.0; }
    else { int32_t r = (v - 1) / S + 1, s = (v - 1) % S; *theta = PI_D * r / R; *phi = 2.0 * PI_D * s / S; }
}

int bq2_synth_md2(int32_t S, int32_t R, int32_t F, uint32_t phase_seed, void *blob)
{
    if (S < 3 || R < 2 || F < 1 || !blob) return -1;
    int32_t V = bq2_synth_sphere_verts(S, R), T = bq2_synth_sphere_tris(S, R);
    if (V > 2048 || T > 4096) return -1;
    uint8_t *base = (uint8_t *)blob;
    md2_header *h = (md2_header *)blob;
    memset(h, 0, sizeof *h);
    h->ident = ('2' << 24) + ('P' << 16) + ('D' << 8) + 'I';
    h->version = 8;
    h->skinwidth = 256; h->skinheight = 256;
    h->framesize = 40 + 4 * V;
    h->num_xyz = V; h->num_st = V; h->num_tris = T; h->num_frames = F;
    h->ofs_skins = (int32_t)sizeof *h;
    h->ofs_st = (int32_t)sizeof *h;
    h->ofs_tris = h->ofs_st + 4 * V;
    h->ofs_frames = h->ofs_tris + 12 * T;
    h->ofs_glcmds = h->ofs_frames + F * h->framesize;
    h->ofs_end = h->ofs_glcmds;

    int16_t *st = (int16_t *)(base + h->ofs_st);
    for (int32_t v = 0; v < V; v++) {
        double th, ph; sphere_dir(S, R, v, &th, &ph);
        st[2*v]   = (int16_t)(ph / (2.0 * PI_D) * 255.0 + 0.5);
        st[2*v+1] = (int16_t)(th / PI_D * 255.0 + 0.5);
    }
    int16_t *tri = (int16_t *)(base + h->ofs_tris);
    for (int32_t t = 0; t < T; t++) {
        int32_t ix[3]; sphere_tri(S, R, t, ix);
        for (int k = 0; k < 3; k++) { tri[6*t+k] = (int16_t)ix[k]; tri[6*t+3+k] = (int16_t)ix[k]; }
    }
    double phase = (double)(hash32(phase_seed * 2654435761u + 17u) & 0xffff) / 65536.0 * 2.0 * PI_D;
    uint8_t lni[2048];                              /* the direction, hence the anorm, is per vertex */
    for (int32_t v = 0; v < V; v++) {
        double th, ph; sphere_dir(S, R, v, &th, &ph);
        float df[3] = { (float)(sin(th) * cos(ph)), (float)(sin(th) * sin(ph)), (float)cos(th) };
        lni[v] = bq2_host_normal2index(df);
    }
    for (int32_t f = 0; f < F; f++) {
        uint8_t *fr = base + h->ofs_frames + (size_t)f * h->framesize;
        float *sc = (float *)fr, *tr = sc + 3;
        for (int k = 0; k < 3; k++) {
            sc[k] = 0.25f * (1.0f + 0.05f * unit_pm(phase_seed * 977u + (uint32_t)f * 6u + (uint32_t)k));
            tr[k] = -32.0f * (1.0f + 0.05f * unit_pm(phase_seed * 977u + (uint32_t)f * 6u + 3u + (uint32_t)k));
        }
        memset(fr + 24, 0, 16);
        snprintf((char *)fr + 24, 16, "syn%03d", (int)(f % 1000));
        uint8_t *vb = fr + 40;
        for (int32_t v = 0; v < V; v++) {
            double th, ph; sphere_dir(S, R, v, &th, &ph);
            double st_ = sin(th), d[3] = { st_ * cos(ph), st_ * sin(ph), cos(th) };
            double rho = 100.0 + 20.0 * sin(3.0 * ph + 0.1 * f + phase) * st_;
            for (int k = 0; k < 3; k++) {
                double x = floor(128.0 + rho * d[k] + 0.5);
                vb[4*v+k] = (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
            }
            vb[4*v+3] = lni[v];
        }
    }
    return 0;
}

/* fstvert_t as the loader computes them, M:51322-51329 */
void bq2_synth_md2_st(const void *blob, float *st)
{
    const md2_header *h = (const md2_header *)blob;
    const int16_t *in = (const int16_t *)((const uint8_t *)blob + h->ofs_st);
    float iw = 1.0 / h->skinwidth, ih = 1.0 / h->skinheight;
    for (int32_t i = 0; i < h->num_st; i++) {
        float s = in[2*i], t = in[2*i+1];
        st[2*i]   = (s - 0.5) * iw;
        st[2*i+1] = (t - 0.5) * ih;
    }
}

typedef struct { float point[3]; uint8_t normal, tangent, binormal, pad; } md3_vert;

int bq2_synth_md3_mesh(int32_t S, int32_t R, int32_t F, uint32_t phase_seed, const float centre[3],
                       void *vertexes, uint16_t *indexes, float *st)
{
    if (S < 3 || R < 2 || F < 1 || !vertexes || !indexes) return -1;
    int32_t V = bq2_synth_sphere_verts(S, R), T = bq2_synth_sphere_tris(S, R);
    if (V > 4096 || T > 8192) return -1;
    md3_vert *out = (md3_vert *)vertexes;
    for (int32_t t = 0; t < T; t++) {
        int32_t ix[3]; sphere_tri(S, R, t, ix);
        for (int k = 0; k < 3; k++) indexes[3*t+k] = (uint16_t)ix[k];
    }
    double phase = (double)(hash32(phase_seed * 2654435761u + 29u) & 0xffff) / 65536.0 * 2.0 * PI_D;
    uint8_t lni[4096];
    for (int32_t v = 0; v < V; v++) {
        double th, ph; sphere_dir(S, R, v, &th, &ph);
        float df[3] = { (float)(sin(th) * cos(ph)), (float)(sin(th) * sin(ph)), (float)cos(th) };
        lni[v] = bq2_host_normal2index(df);
    }
    for (int32_t f = 0; f < F; f++)
        for (int32_t v = 0; v < V; v++) {
            double th, ph; sphere_dir(S, R, v, &th, &ph);
            double st_ = sin(th), d[3] = { st_ * cos(ph), st_ * sin(ph), cos(th) };
            double rho = 24.0 + 5.0 * sin(3.0 * ph + 0.1 * f + phase) * st_;
            md3_vert *o = out + (size_t)f * V + v;
            for (int k = 0; k < 3; k++) {
                /* MD3 stores shorts at 1/64 unit: point = (float)short * (1/64)  (M:51047) */
                int16_t q = (int16_t)floor((centre[k] + rho * d[k]) * 64.0 + 0.5);
                o->point[k] = (float)q * (1.0f / 64.0f);
            }
            o->normal = lni[v]; o->tangent = 0; o->binormal = 0; o->pad = 0;
            if (st && f == 0) { st[2*v] = (float)(ph / (2.0 * PI_D)); st[2*v+1] = (float)(th / PI_D); }
        }
    return 0;
}

void bq2_synth_scene(int32_t n_ents, int32_t L, int32_t num_frames, int32_t n_models,
                     uint32_t seed, uint32_t step, bq2_world_entity *ents, float *lights)
{
    uint32_t x = seed;
    int32_t side = 1; while ((int64_t)side * side < n_ents) side++;
    for (int32_t e = 0; e < n_ents; e++) {
        bq2_world_entity *E = &ents[e];
        x = 1664525u * x + 1013904223u;
        E->model = n_models > 0 ? e % n_models : 0;
        E->frame = (int32_t)(((x >> 8) + step) % (uint32_t)num_frames);
        E->oldframe = (E->frame + num_frames - 1) % num_frames;
        E->backlerp = (float)((x >> 4) & 1023) / 1024.0f;
        E->rf_flags = 0;
        E->origin[0] = 64.0f * (float)(e % side); E->origin[1] = 64.0f * (float)(e / side); E->origin[2] = 24.0f;
        x = 1664525u * x + 1013904223u;
        for (int k = 0; k < 3; k++)
            E->oldorigin[k] = E->origin[k] + ((float)((x >> (8 * k)) & 255) - 128.0f) / 64.0f;
        x = 1664525u * x + 1013904223u;
        E->angles[0] = E->angles[1] = E->angles[2] = 0.0f;
        if ((x >> 16) % 10u == 0) E->angles[1] = (float)((x >> 4) & 4095) * (360.0f / 4096.0f);
        for (int32_t l = 0; l < L; l++) {
            float *Lw = lights + 3 * ((size_t)e * L + l);
            x = 1664525u * x + 1013904223u;
            double r  = 150.0 + 250.0 * (double)((x >> 8) & 0xffff) / 65536.0;
            x = 1664525u * x + 1013904223u;
            double az = 2.0 * PI_D * (double)((x >> 8) & 0xffff) / 65536.0;
            x = 1664525u * x + 1013904223u;
            double el = (PI_D / 2.0) * ((double)((x >> 8) & 0xffff) / 65536.0 * 1.6 - 0.6);   /* mostly above */
            Lw[0] = E->origin[0] + (float)(r * cos(el) * cos(az));
            Lw[1] = E->origin[1] + (float)(r * cos(el) * sin(az));
            Lw[2] = E->origin[2] + (float)(r * sin(el));
        }
    }
}
