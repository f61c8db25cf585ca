/*
 * bq2_cache.c -- the data formats either side of the path, host C:
 *   - an MD2 file image -> the in-memory image the engine (and bq2_upload_md2) works on
 *     (Mod_LoadAliasModel M:51216-51345: sanity checks, load scale, optional winding inversion, float st);
 *   - the engine's precompute cache files, so a model's neighbour table and tangent-space bytes can be
 *     exchanged with an existing installation instead of being recomputed:
 *       MD2  <gamedir>/cache/<model>        write M:51811-51838, read + validation M:51501-51606
 *       MD3  <gamedir>/cache/<model>_<mesh> write M:51017-51034, read M:50919-50947
 *   - Com_BlockChecksum (M:40589-40602): the four words of an MD4 digest (RFC 1320) XOR-ed together.
 * Everything works on memory buffers; the caller does the file I/O.
 */
#include "bq2_gpu.h"
#include <string.h>
#include <stddef.h>

/* ---- MD4 (RFC 1320), folded to 32 bits as Com_BlockChecksum does --------------------------------------- */
#define ROL(x, n) (((x) << (n)) | ((x) >> (32 - (n))))
static void md4_block(uint32_t st[4], const uint8_t *p)
{
    uint32_t x[16], a = st[0], b = st[1], c = st[2], d = st[3];
    static const int r1[4] = {3, 7, 11, 19}, r2[4] = {3, 5, 9, 13}, r3[4] = {3, 9, 11, 15};
    static const int o2[16] = {0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15};
    static const int o3[16] = {0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15};
    int i;
    for (i = 0; i < 16; i++)
        x[i] = (uint32_t)p[4*i] | (uint32_t)p[4*i+1] << 8 | (uint32_t)p[4*i+2] << 16 | (uint32_t)p[4*i+3] << 24;
    for (i = 0; i < 16; i++) {          /* round 1: F(b,c,d) = (b & c) | (~b & d) */
        uint32_t t = a + ((b & c) | (~b & d)) + x[i];
        a = d; d = c; c = b; b = ROL(t, r1[i & 3]);
    }
    for (i = 0; i < 16; i++) {          /* round 2: G = majority, + 0x5A827999 */
        uint32_t t = a + ((b & c) | (b & d) | (c & d)) + x[o2[i]] + 0x5A827999u;
        a = d; d = c; c = b; b = ROL(t, r2[i & 3]);
    }
    for (i = 0; i < 16; i++) {          /* round 3: H = xor, + 0x6ED9EBA1 */
        uint32_t t = a + (b ^ c ^ d) + x[o3[i]] + 0x6ED9EBA1u;
        a = d; d = c; c = b; b = ROL(t, r3[i & 3]);
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d;
}

uint32_t bq2_block_checksum(const void *buffer, size_t length)
{
    uint32_t st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    const uint8_t *p = (const uint8_t *)buffer;
    uint8_t tail[128];
    size_t n = length, rem, pad;
    uint64_t bits = (uint64_t)length * 8;
    int i;
    while (n >= 64) { md4_block(st, p); p += 64; n -= 64; }
    rem = n;
    memset(tail, 0, sizeof tail);
    if (rem) memcpy(tail, p, rem);
    tail[rem] = 0x80;
    pad = rem < 56 ? 64 : 128;
    for (i = 0; i < 8; i++) tail[pad - 8 + i] = (uint8_t)(bits >> (8 * i));
    md4_block(st, tail);
    if (pad == 128) md4_block(st, tail + 64);
    return st[0] ^ st[1] ^ st[2] ^ st[3];
}

/* ---- MD2 file -> in-memory image (M:51216-51345) ------------------------------------------------------- */
typedef struct {
    int32_t ident, version, skinwidth, skinheight, framesize;
    int32_t num_skins, num_xyz, num_st, num_tris, num_glcmds, num_frames;
    int32_t ofs_skins, ofs_st, ofs_tris, ofs_frames, ofs_glcmds, ofs_end;
} md2_header;

extern uint8_t bq2_host_normal2index(const float v[3]);          /* bq2_synth.c: Normal2Index M:25206-25223 */
extern const float *bq2_host_anorms(void);                      /* bq2_synth.c: the 162 x 4 anorms table */

/* a file's offsets need not be aligned: multi-byte fields are read and written through memcpy */
static float   rd_f32(const uint8_t *p) { float v; memcpy(&v, p, 4); return v; }
static int16_t rd_i16(const uint8_t *p) { int16_t v; memcpy(&v, p, 2); return v; }
static void    wr_f32(uint8_t *p, float v) { memcpy(p, &v, 4); }
static void    wr_i16(uint8_t *p, int16_t v) { memcpy(p, &v, 2); }

bq2_status bq2_host_md2_from_file(const void *file, size_t file_bytes, float scale, int invert,
                                  void *blob_out, size_t blob_capacity, float *st_out)
{
    md2_header hdr;
    const md2_header *h = &hdr;
    const uint8_t *in = (const uint8_t *)file;
    uint8_t *out = (uint8_t *)blob_out;
    int i, j;
    if (!file || !blob_out || file_bytes < sizeof *h) return BQ2_E_INVALID;
    memcpy(&hdr, file, sizeof hdr);
    if (scale <= 0.001) return BQ2_E_INVALID;                                   /* "invalid scale" */
    if (h->version != 8) return BQ2_E_INVALID;                                  /* ALIAS_VERSION */
    if (h->num_xyz <= 0 || h->num_st <= 0 || h->num_tris <= 0 || h->num_frames <= 0) return BQ2_E_INVALID;
    if (h->num_xyz > BQ2_MD2_MAX_VERTS || h->num_tris > BQ2_MD2_MAX_TRIANGLES) return BQ2_E_LIMIT;
    if (h->ofs_end < (int32_t)sizeof *h || (size_t)h->ofs_end > file_bytes || (size_t)h->ofs_end > blob_capacity) return BQ2_E_INVALID;
    if (h->framesize < 40 + 4 * h->num_xyz || h->ofs_tris < 0 || h->ofs_frames < 0 || h->ofs_st < 0 ||
        (size_t)h->ofs_tris + 12u * (size_t)h->num_tris > (size_t)h->ofs_end ||
        (size_t)h->ofs_st + 4u * (size_t)h->num_st > (size_t)h->ofs_end ||
        (size_t)h->ofs_frames + (size_t)h->num_frames * (size_t)h->framesize > (size_t)h->ofs_end)
        return BQ2_E_INVALID;
    memcpy(out, in, (size_t)h->ofs_end);                    /* little-endian host: LittleLong/LittleShort are identities */
    {
        const uint8_t *pin = in + h->ofs_tris;
        uint8_t *pout = out + h->ofs_tris;
        if (invert)                                         /* invert the index sequence, M:51274-51281 */
            for (i = 0; i < h->num_tris; i++)
                for (j = 0; j < 3; j++) {
                    wr_i16(pout + 2 * (6*i + j), rd_i16(pin + 2 * (6*i + 2 - j)));
                    wr_i16(pout + 2 * (6*i + 3 + j), rd_i16(pin + 2 * (6*i + 3 + 2 - j)));
                }
    }
    for (i = 0; i < h->num_frames; i++) {
        const uint8_t *fi = in + h->ofs_frames + (size_t)i * h->framesize;
        uint8_t *fo = out + h->ofs_frames + (size_t)i * h->framesize;
        for (j = 0; j < 6; j++) {                           /* scale[3], translate[3]; y mirrored when inverted */
            float v = rd_f32(fi + 4 * j) * scale;
            wr_f32(fo + 4 * j, (invert && (j == 1 || j == 4)) ? -v : v);
        }
    }
    if (invert) {                                           /* Mod_InvertNormal M:61850-61871, run by the loader at M:51657 */
        const float *an = bq2_host_anorms();
        uint8_t map[256];
        for (i = 0; i < 256; i++) map[i] = (uint8_t)i;      /* indices past the table are not touched (the engine would read out of bounds) */
        for (i = 0; i < 162; i++) {
            float n[3] = { an[4*i], -an[4*i + 1], an[4*i + 2] };
            map[i] = bq2_host_normal2index(n);
        }
        for (i = 0; i < h->num_frames; i++) {
            uint8_t *v = out + h->ofs_frames + (size_t)i * h->framesize + 40;
            for (j = 0; j < h->num_xyz; j++) v[4*j + 3] = map[v[4*j + 3]];
        }
    }
    if (st_out) {                                           /* fstvert_t, M:51322-51329 */
        const uint8_t *pst = in + h->ofs_st;
        float iw = 1.0 / h->skinwidth, ih = 1.0 / h->skinheight;
        for (i = 0; i < h->num_st; i++) {
            float s = rd_i16(pst + 4*i), t = rd_i16(pst + 4*i + 2);
            st_out[2*i] = (s - 0.5) * iw;
            st_out[2*i + 1] = (t - 0.5) * ih;
        }
    }
    return BQ2_OK;
}

/* ---- MD2 cache file ------------------------------------------------------------------------------------ */
/* layout (M:51811-51838): u8 smooth; u32 (unsigned)(smooth_cosine * 0x7fffffff); u32 crc; neighbours_t[T];
 * u32 crc; binormals[rows*F]; u32 crc; tangents[rows*F]; and when !smooth: u32 crc; normals[rows*F] */
size_t bq2_md2_cache_bytes(int smooth, int32_t num_tris, int32_t rows, int32_t num_frames)
{
    size_t cx = (size_t)rows * (size_t)num_frames;
    return 1 + 4 + 4 + 12 * (size_t)num_tris + (smooth ? 2 : 3) * (4 + cx);
}
static uint32_t cache_angle(float smooth_cosine) { return (unsigned)(smooth_cosine * 0x7fffffff); }
static uint8_t *put(uint8_t *p, const void *src, size_t n) { memcpy(p, src, n); return p + n; }

bq2_status bq2_md2_cache_write(void *out, size_t capacity, int smooth, float smooth_cosine,
                               const int32_t *neighbours, int32_t num_tris, int32_t rows, int32_t num_frames,
                               const uint8_t *binormals, const uint8_t *tangents, const uint8_t *normals)
{
    uint8_t *p = (uint8_t *)out, sm = smooth ? 1 : 0;
    size_t cx = (size_t)rows * (size_t)num_frames, ax = 12 * (size_t)num_tris;
    uint32_t w;
    if (!out || !neighbours || !binormals || !tangents || (!smooth && !normals) || num_tris <= 0 || rows <= 0 || num_frames <= 0)
        return BQ2_E_INVALID;
    if (capacity < bq2_md2_cache_bytes(smooth, num_tris, rows, num_frames)) return BQ2_E_LIMIT;
    p = put(p, &sm, 1);
    w = cache_angle(smooth_cosine); p = put(p, &w, 4);
    w = bq2_block_checksum(neighbours, ax); p = put(p, &w, 4); p = put(p, neighbours, ax);
    w = bq2_block_checksum(binormals, cx); p = put(p, &w, 4); p = put(p, binormals, cx);
    w = bq2_block_checksum(tangents, cx); p = put(p, &w, 4); p = put(p, tangents, cx);
    if (!smooth) { w = bq2_block_checksum(normals, cx); p = put(p, &w, 4); p = put(p, normals, cx); }
    return BQ2_OK;
}

/* Returns BQ2_OK; BQ2_E_INVALID for "invalid cache" (smooth flag or angle differs, truncated, neighbour checksum);
 * BQ2_E_OVERFLOW is never used here; a tangent-space checksum mismatch ("invalid checksum!") is BQ2_E_LIMIT + 0:
 * reported as BQ2_E_INVALID too, with *why = 2 (1 = invalid cache). */
bq2_status bq2_md2_cache_read(const void *in, size_t bytes, int smooth, float smooth_cosine,
                              int32_t num_tris, int32_t rows, int32_t num_frames,
                              int32_t *neighbours, uint8_t *binormals, uint8_t *tangents, uint8_t *normals, int32_t *why)
{
    const uint8_t *p = (const uint8_t *)in;
    size_t cx = (size_t)rows * (size_t)num_frames, ax = 12 * (size_t)num_tris;
    uint32_t ang, cs_tris, cs_b, cs_t, cs_n = 0;
    if (why) *why = 1;
    if (!in || !neighbours || !binormals || !tangents || (!smooth && !normals)) return BQ2_E_INVALID;
    if (bytes < 1 || (p[0] != 0) != (smooth != 0)) return BQ2_E_INVALID;            /* M:51503 */
    if (bytes < bq2_md2_cache_bytes(smooth, num_tris, rows, num_frames)) return BQ2_E_INVALID;
    memcpy(&ang, p + 1, 4);
    if (ang != cache_angle(smooth_cosine)) return BQ2_E_INVALID;                     /* M:51508 */
    p += 5;
    memcpy(&cs_tris, p, 4); p += 4; memcpy(neighbours, p, ax); p += ax;
    memcpy(&cs_b, p, 4); p += 4; memcpy(binormals, p, cx); p += cx;
    memcpy(&cs_t, p, 4); p += 4; memcpy(tangents, p, cx); p += cx;
    if (!smooth) { memcpy(&cs_n, p, 4); p += 4; memcpy(normals, p, cx); p += cx; }
    if (bq2_block_checksum(neighbours, ax) != cs_tris) return BQ2_E_INVALID;          /* M:51580-51586 */
    if (why) *why = 2;
    if (bq2_block_checksum(binormals, cx) != cs_b || bq2_block_checksum(tangents, cx) != cs_t ||
        (!smooth && bq2_block_checksum(normals, cx) != cs_n)) return BQ2_E_INVALID;  /* M:51590-51604 */
    if (why) *why = 0;
    return BQ2_OK;
}

/* ---- MD3 per-mesh cache file (M:51017-51034 / M:50919-50947) ------------------------------------------- */
/* u32 file checksum (Com_BlockChecksum of the .md3 file), then {normal, tangent, binormal} per vertex per frame.
 * vertexes = maliasvertex_t[num_frames * num_verts] (16 bytes each, bytes 12..14). */
size_t bq2_md3_cache_bytes(int32_t num_frames, int32_t num_verts) { return 4 + 3 * (size_t)num_frames * (size_t)num_verts; }

bq2_status bq2_md3_cache_write(void *out, size_t capacity, uint32_t file_checksum, const void *vertexes,
                               int32_t num_frames, int32_t num_verts)
{
    uint8_t *p = (uint8_t *)out;
    const uint8_t *v = (const uint8_t *)vertexes;
    size_t n = (size_t)num_frames * (size_t)num_verts, i;
    if (!out || !vertexes || num_frames <= 0 || num_verts <= 0) return BQ2_E_INVALID;
    if (capacity < bq2_md3_cache_bytes(num_frames, num_verts)) return BQ2_E_LIMIT;
    memcpy(p, &file_checksum, 4); p += 4;
    for (i = 0; i < n; i++) { p[0] = v[16*i + 12]; p[1] = v[16*i + 13]; p[2] = v[16*i + 14]; p += 3; }
    return BQ2_OK;
}

bq2_status bq2_md3_cache_read(const void *in, size_t bytes, uint32_t file_checksum, void *vertexes,
                              int32_t num_frames, int32_t num_verts)
{
    const uint8_t *p = (const uint8_t *)in;
    uint8_t *v = (uint8_t *)vertexes;
    size_t n = (size_t)num_frames * (size_t)num_verts, i;
    uint32_t cs;
    if (!in || !vertexes || bytes < 4) return BQ2_E_INVALID;
    memcpy(&cs, p, 4);
    if (cs != file_checksum) return BQ2_E_INVALID;                                    /* "invalid cache", M:50921-50923 */
    if (bytes < bq2_md3_cache_bytes(num_frames, num_verts)) return BQ2_E_INVALID;     /* short read, M:50929-50937 */
    p += 4;
    for (i = 0; i < n; i++) { v[16*i + 12] = p[0]; v[16*i + 13] = p[1]; v[16*i + 14] = p[2]; p += 3; }
    return BQ2_OK;
}

/* ---- MD3 file -> the engine's in-memory meshes (Mod_LoadAliasMD3Model M:50584-51060) ---------------------- */
/* on-disk structs, defs.h:4298-4391 (little-endian, packed naturally) */
typedef struct { int32_t id, version; char filename[64]; int32_t flags, num_frames, num_tags, num_meshes, num_skins,
                 ofs_frames, ofs_tags, ofs_meshes, ofs_end; } md3_file_header;                       /* dmd3_t */
typedef struct { float mins[3], maxs[3], translate[3], radius; char creator[16]; } md3_file_frame;    /* dmd3frame_t */
typedef struct { char id[4]; char name[64]; int32_t flags, num_frames, num_skins, num_verts, num_tris,
                 ofs_tris, ofs_skins, ofs_tcs, ofs_verts, meshsize; } md3_file_mesh;                  /* dmd3mesh_t */
typedef struct { int16_t point[3]; int16_t norm; } md3_file_vertex;                                   /* dmd3vertex_t */

#define _GNU_SOURCE
#include <math.h>
extern void sincosf(float, float *, float *);

/* header of mesh m (a copy: the offsets of a file need not be aligned) and its offset; NULL when the chain of
   meshsize fields leaves the file */
static const md3_file_mesh *md3_mesh_at(const uint8_t *file, size_t bytes, const md3_file_header *h, int m,
                                        md3_file_mesh *copy, size_t *off_out)
{
    size_t off;
    int i;
    if (h->ofs_meshes < 0) return NULL;
    off = (size_t)h->ofs_meshes;
    for (i = 0;; i++) {
        if (off > bytes || bytes - off < sizeof(md3_file_mesh)) return NULL;
        memcpy(copy, file + off, sizeof *copy);
        if (i == m) { if (off_out) *off_out = off; return copy; }
        if (copy->meshsize <= 0) return NULL;
        off += (size_t)copy->meshsize;
    }
}

/* Header checks of the loader (M:50620-50660, 50722-50756) and the mesh dimensions.  num_verts / num_tris =
 * int32[num_meshes] (may be NULL). */
bq2_status bq2_md3_file_info(const void *file, size_t bytes, int32_t *num_frames, int32_t *num_meshes,
                             int32_t *num_verts, int32_t *num_tris)
{
    md3_file_header hdr;
    const md3_file_header *h = &hdr;
    md3_file_mesh mesh_copy;
    int m;
    if (!file || bytes < sizeof *h) return BQ2_E_INVALID;
    memcpy(&hdr, file, sizeof hdr);
    if (h->version != 15) return BQ2_E_INVALID;                                   /* MD3_ALIAS_VERSION */
    if (h->num_frames <= 0 || h->num_meshes <= 0) return BQ2_E_INVALID;
    if (h->num_frames > 1024 || h->num_meshes > BQ2_MD3_MAX_MESHES) return BQ2_E_LIMIT;
    if (h->ofs_frames < 0 || (size_t)h->ofs_frames + (size_t)h->num_frames * sizeof(md3_file_frame) > bytes) return BQ2_E_INVALID;
    for (m = 0; m < h->num_meshes; m++) {
        size_t off;
        const md3_file_mesh *pm = md3_mesh_at((const uint8_t *)file, bytes, h, m, &mesh_copy, &off);
        if (!pm || memcmp(pm->id, "IDP3", 4)) return BQ2_E_INVALID;               /* "wrong id" */
        if (pm->num_skins <= 0 || pm->num_tris <= 0 || pm->num_verts <= 0) return BQ2_E_INVALID;
        if (pm->num_skins > 32 || pm->num_tris > BQ2_MD3_MAX_TRIANGLES || pm->num_verts > BQ2_MD3_MAX_VERTS) return BQ2_E_LIMIT;
        if (pm->ofs_tris < 0 || pm->ofs_tcs < 0 || pm->ofs_verts < 0 ||
            off + (size_t)pm->ofs_tris + 12u * (size_t)pm->num_tris > bytes ||
            off + (size_t)pm->ofs_tcs + 8u * (size_t)pm->num_verts > bytes ||
            off + (size_t)pm->ofs_verts + 8u * (size_t)pm->num_verts * (size_t)h->num_frames > bytes) return BQ2_E_INVALID;
        if (num_verts) num_verts[m] = pm->num_verts;
        if (num_tris) num_tris[m] = pm->num_tris;
    }
    if (num_frames) *num_frames = h->num_frames;
    if (num_meshes) *num_meshes = h->num_meshes;
    return BQ2_OK;
}

/* One mesh: vertexes = maliasvertex_t[num_frames * num_verts] with point de-quantised (short * (1/64 * scale),
 * y mirrored when invert) and the normal byte from the lat/lng pair (M:50963-50989); tangent / binormal bytes are
 * left 0 for bq2_gpu_md3_tangents or bq2_md3_cache_read.  indexes = u16[3 * num_tris] (winding reversed when
 * invert), st = float[num_verts][2].  frame_translate = float[num_frames][3] (may be NULL; same for every mesh). */
bq2_status bq2_host_md3_mesh_from_file(const void *file, size_t bytes, int32_t mesh, float scale, int invert,
                                       float *frame_translate, void *vertexes, uint16_t *indexes, float *st)
{
    md3_file_header hdr;
    const md3_file_header *h = &hdr;
    const uint8_t *base = (const uint8_t *)file;
    md3_file_mesh mesh_copy;
    const md3_file_mesh *pm;
    size_t off = 0;
    float invscale[3];
    int i, j, l, y;
    bq2_status s = bq2_md3_file_info(file, bytes, NULL, NULL, NULL, NULL);
    if (s != BQ2_OK) return s;
    memcpy(&hdr, file, sizeof hdr);
    if (scale <= 0.001 || mesh < 0 || mesh >= h->num_meshes || !vertexes || !indexes) return BQ2_E_INVALID;
    pm = md3_mesh_at(base, bytes, h, mesh, &mesh_copy, &off);
    if (!pm) return BQ2_E_INVALID;
    invscale[0] = (1.0 / 64) * scale; invscale[1] = invert ? -(1.0 / 64) * scale : (1.0 / 64) * scale; invscale[2] = (1.0 / 64) * scale;
    if (frame_translate) {
        const uint8_t *pf = base + h->ofs_frames + offsetof(md3_file_frame, translate);
        for (i = 0; i < h->num_frames; i++) {
            for (j = 0; j < 3; j++) frame_translate[3*i + j] = rd_f32(pf + sizeof(md3_file_frame) * (size_t)i + 4 * j) * scale;   /* M:50672 */
            if (invert) frame_translate[3*i + 1] = -frame_translate[3*i + 1];
        }
    }
    {
        const uint8_t *ptri = base + off + pm->ofs_tris;
        for (j = 0; j < pm->num_tris; j++, ptri += 12) {
            uint32_t pin[3];
            memcpy(pin, ptri, 12);
            if (pin[0] >= (uint32_t)pm->num_verts || pin[1] >= (uint32_t)pm->num_verts || pin[2] >= (uint32_t)pm->num_verts) return BQ2_E_INVALID;
            indexes[3*j + 0] = (uint16_t)(invert ? pin[2] : pin[0]);
            indexes[3*j + 1] = (uint16_t)pin[1];
            indexes[3*j + 2] = (uint16_t)(invert ? pin[0] : pin[2]);
        }
    }
    if (st) memcpy(st, base + off + pm->ofs_tcs, 8u * (size_t)pm->num_verts);
    {
        const uint8_t *pvb = base + off + pm->ofs_verts;
        uint8_t *out = (uint8_t *)vertexes;
        for (l = 0; l < h->num_frames; l++)
            for (j = 0; j < pm->num_verts; j++, pvb += sizeof(md3_file_vertex), out += 16) {
                float lat, lng, norma[3], slat, clat, slng, clng;
                const int16_t norm = rd_i16(pvb + 6);
                for (y = 0; y < 3; y++) wr_f32(out + 4 * y, (float)rd_i16(pvb + 2 * y) * invscale[y]);
                lat = (norm >> 8) & 0xff;
                lng = (norm & 0xff);
                lat *= M_PI / 128;
                lng *= M_PI / 128;
                sincosf(lat, &slat, &clat);
                sincosf(lng, &slng, &clng);
                norma[0] = clat * slng;
                norma[1] = invert ? -slat * slng : slat * slng;
                norma[2] = clng;
                out[12] = bq2_host_normal2index(norma); out[13] = 0; out[14] = 0; out[15] = 0;
            }
    }
    return BQ2_OK;
}
