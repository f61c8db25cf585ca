/*
 * bq2_dropin.c -- the reference's own entry points, re-implemented on libbq2gpu.so.
 *
 * This file is what a maintainer of Berserker@Quake2 adds to the engine.  It is compiled
 * AGAINST THE ENGINE'S HEADERS ("defs.h" comes from the engine tree on the include path -- it is not part of
 * this repository) and defines, with the engine's exact signatures and global-variable protocol,
 *
 *     void R_CalcAliasFrameLerp        (dmdl_t *paliashdr);     replaces main.c:63510-63583
 *     void R_CalcAliasFrameLerpShadow  (dmdl_t *paliashdr);     replaces main.c:63586-63631
 *     void R_CalcAliasFrameShadowVolume(dmdl_t *paliashdr);     replaces main.c:63634-63710
 *
 * Inputs are read from the engine globals the originals read (currententity, currentmodel,
 * currentshadowlight -- already in model space, main.c:63733-63745), outputs are written to the globals the
 * originals write (tempVertexArray, extrudedVerts, viss, TrisVerts, TrisCounter; data.h:2731-2736, 2827).
 * R_DrawAliasShadowVolume (main.c:63713) calls the last two by name, so with these definitions linked in
 * front of the engine's it drives the GPU with no source change; its GL code is untouched.
 * R_DrawMD3AliasShadowVolume (main.c:63800) is monolithic: bq2_dropin_md3_mesh() below is the body of its
 * per-mesh compute block (main.c:63885-63991) and INTEGRATION.md shows the five-line patch that calls it.
 *
 * Each call is a batch of one.  That is the compatibility path, not the fast one: the engine's light x entity
 * loops (main.c:72113-72280, 64062) should collect pairs and call bq2_frame once per frame; INTEGRATION.md
 * shows that loop too.  There is no CPU fallback: if the device or a launch fails the shim reports through
 * Com_Error-style bq2_dropin_last_error() and leaves TrisCounter at 0 (no volume is drawn).
 */
#include "defs.h"                 /* the ENGINE's header: entity_t, model_t, shadowlight_t, dmdl_t, ... */
#include <string.h>
#include <stdlib.h>
#include "bq2_gpu.h"

extern entity_t       *currententity;                   /* data.h:624  */
extern model_t        *currentmodel;                    /* data.h:2626 */
extern shadowlight_t  *currentshadowlight;              /* data.h:2709 */
extern vec3_t          tempVertexArray[MD3_MAX_VERTS];  /* data.h:2827 */
extern vec3_t          extrudedVerts[MD3_MAX_VERTS];    /* data.h:2731 */
extern byte            viss[MD3_MAX_TRIANGLES];
extern int             TrisCounter;
extern float           TrisVerts[36 * MD3_MAX_TRIANGLES];

static bq2_ctx *g_ctx;
static char     g_err[256];
const char *bq2_dropin_last_error(void) { return g_err; }
int bq2_dropin_kernel_launches(void) { return g_ctx ? (int)bq2_kernel_launches(g_ctx) : 0; }

static int fail(const char *what)
{
    const char *e = bq2_last_error(g_ctx);
    strncpy(g_err, what, sizeof g_err - 1);
    if (e && *e) { strncat(g_err, ": ", sizeof g_err - strlen(g_err) - 1); strncat(g_err, e, sizeof g_err - strlen(g_err) - 1); }
    return -1;
}

static int ensure_ctx(void)
{
    if (g_ctx) return 0;
    if (bq2_create(0, &g_ctx) != BQ2_OK) { g_ctx = NULL; return fail("bq2_create"); }
    return 0;
}

/* engine model (keyed by its extradata pointer + neighbour table) -> resident model id */
#define MAX_RESIDENT 512
static struct { const void *key, *key2; int32_t id; } g_models[MAX_RESIDENT];
static int g_num_models;

static int32_t resident_md2(dmdl_t *h)
{
    const void *nb = currentmodel ? (const void *)currentmodel->triangles : NULL;
    for (int i = 0; i < g_num_models; i++)
        if (g_models[i].key == (const void *)h && g_models[i].key2 == nb) return g_models[i].id;
    if (g_num_models == MAX_RESIDENT) { fail("too many resident models"); return -1; }
    int32_t id = -1;
    uint32_t flags = (currentmodel && (currentmodel->flags & RF_SMOOTHVECS)) ? BQ2_MODEL_SMOOTHVECS : 0;
    /* ofs_end is the size of the in-memory image (Mod_LoadAliasModel copies the whole file, main.c:51216) */
    if (bq2_upload_md2(g_ctx, h, (size_t)h->ofs_end, (const int32_t *)nb,
                       currentmodel ? currentmodel->tangents : NULL, currentmodel ? currentmodel->binormals : NULL,
                       currentmodel ? currentmodel->normals : NULL, flags, &id) != BQ2_OK) { fail("bq2_upload_md2"); return -1; }
    g_models[g_num_models].key = h; g_models[g_num_models].key2 = nb; g_models[g_num_models].id = id; g_num_models++;
    return id;
}

static int         g_have_pair;          /* the last frame carried a (entity, light) pair */
static int32_t     g_V, g_T;

static int run_md2(dmdl_t *h, int with_light)
{
    bq2_entity ent; bq2_pair pair; bq2_frame_desc fd; bq2_frame_out out;
    g_have_pair = 0;
    if (ensure_ctx()) return -1;
    int32_t id = resident_md2(h);
    if (id < 0) return -1;
    /* the engine's own host arithmetic for move / frontv / backv, main.c:63526-63553 */
    bq2_host_entity_md2(h, currententity->frame, currententity->oldframe, currententity->backlerp,
                        currententity->origin, currententity->oldorigin, currententity->angles,
                        (uint32_t)currententity->flags, &ent);
    ent.model = id;
    memset(&fd, 0, sizeof fd);
    fd.ents = &ent; fd.n_ents = 1; fd.want = 0;
    if (with_light) {
        memset(&pair, 0, sizeof pair);
        pair.entity = 0;
        pair.light[0] = currentshadowlight->origin[0];          /* already model space (main.c:63733-63745) */
        pair.light[1] = currentshadowlight->origin[1];
        pair.light[2] = currentshadowlight->origin[2];
        pair.scale = 32 * currentshadowlight->radius;            /* main.c:63597 */
        fd.pairs = &pair; fd.n_pairs = 1; fd.want = BQ2_WANT_SHADOW;
        fd.index_stride = 24u * (uint32_t)h->num_tris;           /* worst case: every edge a silhouette */
    }
    if (bq2_frame(g_ctx, &fd, &out) != BQ2_OK) return fail("bq2_frame");
    g_V = h->num_xyz; g_T = h->num_tris;
    if (bq2_download(g_ctx, BQ2_A_LERPED, 0, 3u * (uint64_t)g_V, tempVertexArray) != BQ2_OK) return fail("download lerped");
    if (with_light) {
        uint32_t bits[(MD3_MAX_TRIANGLES + 31) / 32];
        if (bq2_download(g_ctx, BQ2_A_EXTRUDED, 0, 3u * (uint64_t)g_V, extrudedVerts) != BQ2_OK) return fail("download extruded");
        if (bq2_download(g_ctx, BQ2_A_FACING_BITS, 0, (uint64_t)((g_T + 31) / 32), bits) != BQ2_OK) return fail("download facing");
        for (int t = 0; t < g_T; t++) viss[t] = (byte)((bits[t >> 5] >> (t & 31)) & 1u);
        g_have_pair = 1;
    }
    return 0;
}

void R_CalcAliasFrameLerp(dmdl_t *paliashdr)
{
    run_md2(paliashdr, 0);
}

void R_CalcAliasFrameLerpShadow(dmdl_t *paliashdr)
{
    run_md2(paliashdr, 1);
}

void R_CalcAliasFrameShadowVolume(dmdl_t *paliashdr)
{
    int32_t n = 0;
    (void)paliashdr;
    TrisCounter = 0;
    if (!g_have_pair) return;
    /* indices -> the engine's non-indexed stream: TrisVerts[k] = (lerped ++ extruded)[index[k]] */
    if (bq2_expand_trisverts(g_ctx, 0, TrisVerts, 12 * MD3_MAX_TRIANGLES, &n) != BQ2_OK) { fail("bq2_expand_trisverts"); return; }
    TrisCounter = n;
}

/* ---- MD3 -------------------------------------------------------------------------------------------- */
static struct { const void *key; int32_t id; } g_md3[MAX_RESIDENT];
static int g_num_md3;

static int32_t resident_md3(maliasmodel_t *mm)
{
    for (int i = 0; i < g_num_md3; i++) if (g_md3[i].key == (const void *)mm) return g_md3[i].id;
    if (g_num_md3 == MAX_RESIDENT || mm->num_meshes > BQ2_MD3_MAX_MESHES) { fail("md3: limits"); return -1; }
    bq2_md3_mesh meshes[BQ2_MD3_MAX_MESHES];
    float *ft = (float *)malloc(sizeof(float) * 3 * (size_t)mm->num_frames);
    if (!ft) return -1;
    for (int f = 0; f < mm->num_frames; f++) for (int k = 0; k < 3; k++) ft[3 * f + k] = mm->frames[f].translate[k];
    for (int m = 0; m < mm->num_meshes; m++) {
        meshes[m].num_verts = mm->meshes[m].num_verts; meshes[m].num_tris = mm->meshes[m].num_tris;
        meshes[m].vertexes = mm->meshes[m].vertexes; meshes[m].indexes = mm->meshes[m].indexes;
        meshes[m].neighbours = (const int32_t *)mm->meshes[m].triangles;
        meshes[m].casts_shadow = 1;                    /* the skin filter stays in the caller (main.c:63862-63880) */
    }
    int32_t id = -1;
    bq2_status s = bq2_upload_md3(g_ctx, mm->num_frames, ft, meshes, mm->num_meshes, &id);
    free(ft);
    if (s != BQ2_OK) { fail("bq2_upload_md3"); return -1; }
    g_md3[g_num_md3].key = mm; g_md3[g_num_md3].id = id; g_num_md3++;
    return id;
}

/* Body of the per-mesh compute block of R_DrawMD3AliasShadowVolume (main.c:63885-63991): fills
 * tempVertexArray, extrudedVerts, viss, TrisVerts and TrisCounter for mesh `nmesh` of the current entity's
 * model against currentshadowlight (model space).  The first call for an entity/light computes every mesh
 * in one launch; the following calls for the same entity/light only copy out. */
int bq2_dropin_md3_mesh(maliasmodel_t *mm, int nmesh)
{
    static struct { const void *mm; int frame, oldframe; float backlerp, org[3], oorg[3], ang[3], lp[3], radius; } last, now;
    bq2_entity ent; bq2_pair pair; bq2_frame_desc fd; bq2_frame_out out;
    TrisCounter = 0;
    if (ensure_ctx()) return -1;
    int32_t id = resident_md3(mm);
    if (id < 0) return -1;
    memset(&now, 0, sizeof now);
    now.mm = mm; now.frame = currententity->frame; now.oldframe = currententity->oldframe;
    now.backlerp = currententity->backlerp; now.radius = currentshadowlight->radius;
    for (int k = 0; k < 3; k++) {
        now.org[k] = currententity->origin[k]; now.oorg[k] = currententity->oldorigin[k];
        now.ang[k] = currententity->angles[k]; now.lp[k] = currentshadowlight->origin[k];
    }
    if (memcmp(&last, &now, sizeof now)) {
        bq2_host_entity_md3(mm->frames[currententity->frame].translate, mm->frames[currententity->oldframe].translate,
                            currententity->frame, currententity->oldframe, currententity->backlerp,
                            currententity->origin, currententity->oldorigin, currententity->angles, &ent);
        ent.model = id;
        memset(&pair, 0, sizeof pair);
        for (int k = 0; k < 3; k++) pair.light[k] = currentshadowlight->origin[k];
        pair.scale = 32 * currentshadowlight->radius;
        memset(&fd, 0, sizeof fd);
        fd.ents = &ent; fd.n_ents = 1; fd.pairs = &pair; fd.n_pairs = 1; fd.want = BQ2_WANT_SHADOW;
        for (int m = 0; m < mm->num_meshes; m++)
            if (24u * (uint32_t)mm->meshes[m].num_tris > fd.index_stride) fd.index_stride = 24u * (uint32_t)mm->meshes[m].num_tris;
        if (bq2_frame(g_ctx, &fd, &out) != BQ2_OK) return fail("bq2_frame(md3)");
        last = now;
    } else if (bq2_frame_wait(g_ctx, &out) != BQ2_OK) return fail("bq2_frame_wait(md3)");
    {
        int V = mm->meshes[nmesh].num_verts, T = mm->meshes[nmesh].num_tris; int32_t n = 0;
        uint32_t bits[(MD3_MAX_TRIANGLES + 31) / 32];
        if (bq2_download(g_ctx, BQ2_A_LERPED, 3u * (uint64_t)out.ent_vert_offset[nmesh], 3u * (uint64_t)V, tempVertexArray) != BQ2_OK ||
            bq2_download(g_ctx, BQ2_A_EXTRUDED, 3u * (uint64_t)out.pair_vert_offset[nmesh], 3u * (uint64_t)V, extrudedVerts) != BQ2_OK ||
            bq2_download(g_ctx, BQ2_A_FACING_BITS, out.pair_bits_offset[nmesh], (uint64_t)((T + 31) / 32), bits) != BQ2_OK)
            return fail("download(md3)");
        for (int t = 0; t < T; t++) viss[t] = (byte)((bits[t >> 5] >> (t & 31)) & 1u);
        if (bq2_expand_trisverts(g_ctx, nmesh, TrisVerts, 12 * MD3_MAX_TRIANGLES, &n) != BQ2_OK) return fail("expand(md3)");
        TrisCounter = n;
    }
    return 0;
}
