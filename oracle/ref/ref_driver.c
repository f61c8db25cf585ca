/*
 * ref_driver.c -- thin C driver linked in FRONT of the unmodified reference objects
 * (main.o, unpak.o) to expose the reference's own hot-path functions to tests through ctypes.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Compiled only by oracle/ref/build_ref.sh into
 * oracle/_ref/libbq2ref.so; includes the reference's defs.h FROM the reference tree (never
 * copied).  Every bq2ref_* function sets the reference's file-scope globals exactly as the
 * engine would, calls the reference function, and copies the result globals out.
 * No arithmetic of the path is restated here.
 */
#include "defs.h"          /* reference types: entity_t, model_t, shadowlight_t, dmdl_t, ... */
#include <string.h>
#include <math.h>

/* ---- reference globals / functions we touch (data.h) --------------------------------------- */
extern entity_t       *currententity;                   /* data.h:624  */
extern model_t        *currentmodel;                    /* data.h:2626 */
extern shadowlight_t  *currentshadowlight;              /* data.h:2709 */
extern vec3_t          r_origin;                        /* data.h:766  */
extern vec3_t          tempVertexArray[MD3_MAX_VERTS];  /* data.h:2827 */
extern vec3_t          extrudedVerts[MD3_MAX_VERTS];    /* data.h:2731 */
extern vec3_t          trinormals[MD3_MAX_TRIANGLES];
extern byte            viss[MD3_MAX_TRIANGLES];
extern int             TrisCounter;
extern float           TrisVerts[36 * MD3_MAX_TRIANGLES];
extern float           smooth_cosine;                   /* data.h:2927 */
extern bool            r_mirror;
extern float         (*Sqrt)(float);                    /* data.h:2964 */
extern cvar_t         *r_faststencil, *r_shadows, *r_playershadow;
extern glconfig_t      gl_config;
extern glstate_t       gl_state;
extern image_t        *r_notexture;
extern float           r_avertexnormals[162][3];
extern PFNGLACTIVETEXTUREARBPROC        pglActiveTextureARB;
extern PFNGLCLIENTACTIVETEXTUREARBPROC  pglClientActiveTextureARB;

void  BuildSqrtTable(void);
void  R_CalcAliasFrameLerp(dmdl_t *);
void  R_CalcAliasFrameLerpShadow(dmdl_t *);
void  R_CalcAliasFrameShadowVolume(dmdl_t *);
void  R_DrawAliasShadowVolume(void);
void  R_DrawMD3AliasShadowVolume(void);
void  GL_DrawAliasFrameLerpLightARB(dmdl_t *, int);
void  GL_DrawAliasFrameLerpLightARB_vp(dmdl_t *, int);
void  GL_DrawMD3AliasFrameLerpLightARB(maliasmesh_t *, int, int);
void  GL_DrawMD3AliasFrameLerpLightARB_vp(maliasmesh_t *, int, maliasmodel_t *, int);
int   findneighbour(int, int, int, neighbours_t *, dtriangle_t *);
void  R_BuildTriangleNeighbors(neighbours_t *, WORD *, int);
void  VecsForTris(float *, float *, float *, float *, float *, float *, vec3_t, vec3_t);
void  TangentForTrimd3(WORD *, maliasvertex_t *, maliascoord_t *, vec3_t, vec3_t);
byte  Normal2Index(const vec3_t);
float VectorNormalize(vec3_t);
void  AngleVectors(vec3_t, vec3_t, vec3_t, vec3_t);

#define SHADER_ARB6_VALUE 6     /* data.h:2143 (enum shadertype) */

/* ---- GL capture: the light-lerp family hands its stack arrays to glTexCoordPointer --------- */
#define CAP_TMUS 8
static int          cap_tmu;
static const float *cap_ptr[CAP_TMUS];
static int          cap_size[CAP_TMUS];
static const float *cap_vertex_ptr;
static float       *cap_out[CAP_TMUS];          /* destination for each TMU's array, or NULL  */
static float       *cap_vertex_out;
static int          cap_count;                  /* vertices of the last glDrawArrays          */
static int          cap_mesh_hook;              /* MD3: per-mesh readback inside glDrawArrays */

static void APIENTRY cap_active(GLenum t)        { (void)t; }
static void APIENTRY cap_client_active(GLenum t) { cap_tmu = (int)(t - GL_TEXTURE0); }
void glTexCoordPointer(GLint size, GLenum type, GLsizei stride, const GLvoid *p)
{ (void)type; (void)stride; if (cap_tmu >= 0 && cap_tmu < CAP_TMUS) { cap_ptr[cap_tmu] = p; cap_size[cap_tmu] = size; } }
void glVertexPointer(GLint size, GLenum type, GLsizei stride, const GLvoid *p)
{ (void)size; (void)type; (void)stride; cap_vertex_ptr = p; }

/* MD3 per-mesh results (the reference overwrites its globals for every mesh, M:63919) */
typedef struct {
    float *lerped, *extruded, *tris_verts; uint8_t *viss; int *tris_counter;
    const int *vert_off, *tri_off, *tv_off; int n; int drawn; int draws;
} md3_sink_t;
static md3_sink_t md3_sink;
static int md3_mesh_nverts[MD3_MAX_MESHES], md3_mesh_ntris[MD3_MAX_MESHES], md3_mesh_order[MD3_MAX_MESHES], md3_nmesh_cast;

void glDrawArrays(GLenum mode, GLint first, GLsizei count)
{
    (void)mode; (void)first;
    cap_count = count;
    for (int t = 0; t < CAP_TMUS; t++)
        if (cap_out[t] && cap_ptr[t]) memcpy(cap_out[t], cap_ptr[t], sizeof(float) * cap_size[t] * count);
    if (cap_vertex_out && cap_vertex_ptr) memcpy(cap_vertex_out, cap_vertex_ptr, sizeof(float) * 3 * count);
    /* with r_faststencil 0 every casting mesh is drawn twice (incr pass, decr pass; M:64016-64029):
       capture on the second draw of each mesh */
    if (cap_mesh_hook && (++md3_sink.draws & 1) == 0 && md3_sink.drawn < md3_nmesh_cast) {
        int m = md3_mesh_order[md3_sink.drawn++];
        int V = md3_mesh_nverts[m], T = md3_mesh_ntris[m];
        memcpy(md3_sink.lerped + 3 * md3_sink.vert_off[m], tempVertexArray, sizeof(float) * 3 * V);
        memcpy(md3_sink.extruded + 3 * md3_sink.vert_off[m], extrudedVerts, sizeof(float) * 3 * V);
        memcpy(md3_sink.viss + md3_sink.tri_off[m], viss, T);
        memcpy(md3_sink.tris_verts + 3 * md3_sink.tv_off[m], TrisVerts, sizeof(float) * 3 * TrisCounter);
        md3_sink.tris_counter[m] = TrisCounter;
    }
}

static int cap_elems_count = -1;
static int cap_elems_nverts;                    /* MD3 light family draws indexed: copy this many rows */
void glDrawElements(GLenum mode, GLsizei count, GLenum type, const GLvoid *indices)
{
    (void)mode; (void)type; (void)indices;
    cap_elems_count = count;
    for (int t = 0; t < CAP_TMUS; t++)
        if (cap_out[t] && cap_ptr[t]) memcpy(cap_out[t], cap_ptr[t], sizeof(float) * cap_size[t] * cap_elems_nverts);
}
/* R_RotateForEntity (M:26670) is called right after the light went to model space (M:63747):
   snapshot what the compute functions will see. */
static float snap_light[3], snap_view[3];
/* R_DrawEntsShadowVolumes run whole (bq2ref_ents_shadow_volumes below): each of the three volume builders calls
   R_RotateForEntity exactly once, after its own early-outs -- the entity that got this far is recorded here */
static entity_t *sink_ents; static int sink_n, *sink_order, sink_count;
void glTranslatef(GLfloat x, GLfloat y, GLfloat z)
{
    (void)x; (void)y; (void)z;
    if (currentshadowlight) for (int k = 0; k < 3; k++) snap_light[k] = currentshadowlight->origin[k];
    for (int k = 0; k < 3; k++) snap_view[k] = r_origin[k];
    if (sink_order && currententity >= sink_ents && currententity < sink_ents + sink_n)
        sink_order[sink_count++] = (int)(currententity - sink_ents);
}

/* ---- set-up ---------------------------------------------------------------------------------- */
static cvar_t cv_faststencil, cv_shadows, cv_playershadow;
static entity_t      ent;          /* 9 KB each: static, zeroed */
static model_t       mdl;
static shadowlight_t light;
static image_t       skin_img;
static int           inited;

int bq2ref_init(void)
{
    if (inited) return 0;
    BuildSqrtTable();
    Sqrt = sqrtf;                                   /* the reference's default (M:79665-79669, 97006-97010) */
    smooth_cosine = cos(45.0 * M_PI / 180.0);       /* r_smoothangle 45 (M:4427-4435) */
    memset(&cv_faststencil, 0, sizeof cv_faststencil);
    memset(&cv_shadows, 0, sizeof cv_shadows);      cv_shadows.value = 2;       /* r_shadows "2" */
    memset(&cv_playershadow, 0, sizeof cv_playershadow);
    r_faststencil = &cv_faststencil; r_shadows = &cv_shadows; r_playershadow = &cv_playershadow;
    r_mirror = false;
    memset(&gl_config, 0, sizeof gl_config);        /* vbo = 0: the non-VBO branches compute everything */
    pglActiveTextureARB = cap_active;
    pglClientActiveTextureARB = cap_client_active;
    inited = 1;
    return 0;
}
float bq2ref_smooth_cosine(void) { bq2ref_init(); return smooth_cosine; }
const float *bq2ref_anorms(void) { return &r_avertexnormals[0][0]; }

static void set_entity(int frame, int oldframe, float backlerp, const float *origin,
                       const float *oldorigin, const float *angles, int flags)
{
    memset(&ent, 0, sizeof ent);
    ent.model = &mdl; ent.frame = frame; ent.oldframe = oldframe; ent.backlerp = backlerp; ent.flags = flags;
    for (int k = 0; k < 3; k++) { ent.origin[k] = origin[k]; ent.oldorigin[k] = oldorigin[k]; ent.angles[k] = angles[k]; }
    currententity = &ent;
}
static void set_light(const float *org, float radius)
{
    memset(&light, 0, sizeof light);
    for (int k = 0; k < 3; k++) light.origin[k] = org[k];
    light.radius = radius;
    currentshadowlight = &light;
}

/* ---- MD2 ------------------------------------------------------------------------------------- */
/* R_CalcAliasFrameLerp alone (M:63510). flags may carry RF_SHELL_* / RF_WEAPONMODEL. */
int bq2ref_md2_lerp(void *blob, int frame, int oldframe, float backlerp, const float *origin,
                    const float *oldorigin, const float *angles, int flags, float *lerped)
{
    bq2ref_init();
    dmdl_t *h = (dmdl_t *)blob;
    memset(&mdl, 0, sizeof mdl); mdl.type = mod_alias; mdl.extradata = blob; currentmodel = &mdl;
    set_entity(frame, oldframe, backlerp, origin, oldorigin, angles, flags);
    R_CalcAliasFrameLerp(h);
    memcpy(lerped, tempVertexArray, sizeof(float) * 3 * h->num_xyz);
    return h->num_xyz;
}

/* The whole R_DrawAliasShadowVolume (M:63713): world-space light in, every intermediate out.
 * light_local receives currentshadowlight->origin as seen by the compute functions. */
int bq2ref_md2_shadow(void *blob, const int32_t *neighbours, int frame, int oldframe, float backlerp,
                      const float *origin, const float *oldorigin, const float *angles, int flags,
                      const float *light_world, float radius,
                      float *lerped, float *extruded, float *trinorm, uint8_t *viss_out,
                      float *tris_verts, float *light_local)
{
    bq2ref_init();
    dmdl_t *h = (dmdl_t *)blob;
    memset(&mdl, 0, sizeof mdl); mdl.type = mod_alias; mdl.extradata = blob;
    mdl.triangles = (neighbours_t *)neighbours;
    set_entity(frame, oldframe, backlerp, origin, oldorigin, angles, flags);
    set_light(light_world, radius);
    /* poison outputs so "not written" is visible */
    memset(extrudedVerts, 0xff, sizeof(float) * 3 * h->num_xyz);
    TrisCounter = -1;
    R_DrawAliasShadowVolume();
    if (TrisCounter < 0) return -1;                   /* early-out by flags (M:63722-63727) */
    if (lerped)     memcpy(lerped, tempVertexArray, sizeof(float) * 3 * h->num_xyz);
    if (extruded)   memcpy(extruded, extrudedVerts, sizeof(float) * 3 * h->num_xyz);
    if (trinorm)    memcpy(trinorm, trinormals, sizeof(float) * 3 * h->num_tris);
    if (viss_out)   memcpy(viss_out, viss, h->num_tris);
    if (tris_verts) memcpy(tris_verts, TrisVerts, sizeof(float) * 3 * TrisCounter);
    if (light_local) for (int k = 0; k < 3; k++) light_local[k] = snap_light[k];
    return TrisCounter;
}

/* The two compute functions with a light that is ALREADY in model space (M:63586, 63634). */
int bq2ref_md2_shadow_local(void *blob, const int32_t *neighbours, int frame, int oldframe, float backlerp,
                            const float *origin, const float *oldorigin, const float *angles,
                            const float *light_model, float radius,
                            float *lerped, float *extruded, float *trinorm, uint8_t *viss_out, float *tris_verts)
{
    bq2ref_init();
    dmdl_t *h = (dmdl_t *)blob;
    memset(&mdl, 0, sizeof mdl); mdl.type = mod_alias; mdl.extradata = blob;
    mdl.triangles = (neighbours_t *)neighbours; currentmodel = &mdl;
    set_entity(frame, oldframe, backlerp, origin, oldorigin, angles, 0);
    set_light(light_model, radius);
    memset(extrudedVerts, 0xff, sizeof(float) * 3 * h->num_xyz);
    R_CalcAliasFrameLerpShadow(h);
    R_CalcAliasFrameShadowVolume(h);
    if (lerped)     memcpy(lerped, tempVertexArray, sizeof(float) * 3 * h->num_xyz);
    if (extruded)   memcpy(extruded, extrudedVerts, sizeof(float) * 3 * h->num_xyz);
    if (trinorm)    memcpy(trinorm, trinormals, sizeof(float) * 3 * h->num_tris);
    if (viss_out)   memcpy(viss_out, viss, h->num_tris);
    if (tris_verts) memcpy(tris_verts, TrisVerts, sizeof(float) * 3 * TrisCounter);
    return TrisCounter;
}

/* Timing loop for the CPU baseline: n pairs over one model, entity/light lists supplied.
 * Returns the sum of TrisCounter (so the work cannot be optimised away). */
long long bq2ref_md2_bench(void *blob, const int32_t *neighbours, int n_pairs,
                           const int32_t *frames, const int32_t *oldframes, const float *backlerps,
                           const float *lights_model /*[n][3]*/, float radius)
{
    bq2ref_init();
    dmdl_t *h = (dmdl_t *)blob;
    static const float zero[3] = {0, 0, 0};
    long long total = 0;
    memset(&mdl, 0, sizeof mdl); mdl.type = mod_alias; mdl.extradata = blob;
    mdl.triangles = (neighbours_t *)neighbours; currentmodel = &mdl;
    for (int p = 0; p < n_pairs; p++) {
        set_entity(frames[p], oldframes[p], backlerps[p], zero, zero, zero, 0);
        set_light(lights_model + 3 * p, radius);
        R_CalcAliasFrameLerpShadow(h);
        R_CalcAliasFrameShadowVolume(h);
        total += TrisCounter;
    }
    return total;
}

/* CPU-baseline loop over a world-space pair list with one model per pair: the WHOLE of
 * R_DrawAliasShadowVolume (M:63713: flag filter, light -> model space, lerp, facing, silhouette + caps,
 * the (stubbed) GL submit) per pair, exactly what R_DrawEntsShadowVolumes (M:64062) does per marked entity.
 * Returns the sum of TrisCounter. */
long long bq2ref_md2_world_bench(void *const *blobs, void *const *neighbours, int n_pairs,
                                 const int32_t *frames, const int32_t *oldframes, const float *backlerps,
                                 const float *origins, const float *oldorigins, const float *angles,
                                 const float *lights_world, const float *radii)
{
    bq2ref_init();
    long long total = 0;
    for (int p = 0; p < n_pairs; p++) {
        memset(&mdl, 0, sizeof mdl); mdl.type = mod_alias; mdl.extradata = blobs[p];
        mdl.triangles = (neighbours_t *)neighbours[p];
        set_entity(frames[p], oldframes[p], backlerps[p], origins + 3 * p, oldorigins + 3 * p, angles + 3 * p, 0);
        set_light(lights_world + 3 * p, radii[p]);
        TrisCounter = 0;
        R_DrawAliasShadowVolume();
        total += TrisCounter;
    }
    return total;
}

/* Full-size parity: the per-pair 64-bit hash the product computes on the device (csrc/bq2_kernels.cu, bq2_hash_kernel;
 * include/bq2_gpu.h bq2_hash_results), here from the engine's own globals after the WHOLE R_DrawAliasShadowVolume ran
 * for the pair: tempVertexArray, extrudedVerts, viss packed into 32-bit words, TrisCounter, TrisVerts.  The hash is test
 * plumbing (a sum of mixed (section, position, word) triples); no arithmetic of the path is restated.  Every vertex of
 * the model must be referenced by a triangle (the engine extrudes per triangle corner), which holds for the
 * generators' closed meshes. */
static uint64_t hash_word(uint32_t section, uint32_t i, uint32_t w)
{
    uint64_t x = ((uint64_t)section << 56) ^ ((uint64_t)i << 32) ^ (uint64_t)w;
    x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
static uint64_t hash_globals(int V, int T)
{
    uint64_t h = 0;
    const uint32_t *lp = (const uint32_t *)tempVertexArray, *ep = (const uint32_t *)extrudedVerts, *tv = (const uint32_t *)TrisVerts;
    for (int i = 0; i < 3 * V; i++) { h += hash_word(1, (uint32_t)i, lp[i]); h += hash_word(2, (uint32_t)i, ep[i]); }
    for (int c = 0; c < (T + 31) / 32; c++) {
        uint32_t w = 0;
        for (int b = 0; b < 32 && 32 * c + b < T; b++) if (viss[32 * c + b]) w |= 1u << b;
        h += hash_word(3, (uint32_t)c, w);
    }
    h += hash_word(4, 0, (uint32_t)TrisCounter);
    for (int i = 0; i < 3 * TrisCounter; i++) h += hash_word(5, (uint32_t)i, tv[i]);
    return h;
}
long long bq2ref_md2_world_hash(void *const *blobs, void *const *neighbours, int n_pairs,
                                const int32_t *frames, const int32_t *oldframes, const float *backlerps,
                                const float *origins, const float *oldorigins, const float *angles,
                                const float *lights_world, const float *radii, uint64_t *hashes)
{
    bq2ref_init();
    long long total = 0;
    for (int p = 0; p < n_pairs; p++) {
        dmdl_t *h = (dmdl_t *)blobs[p];
        memset(&mdl, 0, sizeof mdl); mdl.type = mod_alias; mdl.extradata = blobs[p];
        mdl.triangles = (neighbours_t *)neighbours[p];
        set_entity(frames[p], oldframes[p], backlerps[p], origins + 3 * p, oldorigins + 3 * p, angles + 3 * p, 0);
        set_light(lights_world + 3 * p, radii[p]);
        TrisCounter = 0;
        R_DrawAliasShadowVolume();
        total += TrisCounter;
        hashes[p] = hash_globals(h->num_xyz, h->num_tris);
    }
    return total;
}

/* ---- the caller loop: R_DrawEntsShadowVolumes M:64041-64133 whole ---------------------- */
extern int        cl_numlightvisedicts;                          /* data.h:2728 */
extern entity_t  *cl_lightvisedicts[];                           /* data.h:2729 */
extern cvar_t    *r_mirror_shadows, *r_mirror_models, *r_drawentities;
void R_DrawEntsShadowVolumes(bool SelfShadow, bool all);

/* n entities of kind 0 (MD2: md2_blob / md2_nb), 1 (MD3: an empty mesh list -- only the early-outs and the transform
 * run) or 2 (brush model without surfaces), each with entity flags and MODEL flags; the cvars and r_mirror as given.
 * Runs the engine's loop and returns how many entities reached their volume builder; order_out = their indices in
 * call order.  Nothing of the loop's logic is restated. */
int bq2ref_ents_shadow_volumes(int n, const int *kinds, const int *ent_flags, const int *model_flags,
                               void *md2_blob, const int32_t *md2_nb,
                               int mirror, float mirror_shadows, float mirror_models, float playershadow, float shadows,
                               float drawentities, int self_shadow, int all, int *order_out)
{
    static cvar_t cv_ms, cv_mm, cv_de;
    static maliasmodel_t mm; static maliasframe_t mframes[1];
    static const float org[3] = {10, 20, 30}, lorg[3] = {100, 50, 80};
    bq2ref_init();
    if (n > 256) return -1;
    entity_t *es = calloc((size_t)n, sizeof *es);
    model_t *ms = calloc((size_t)n, sizeof *ms);
    if (!es || !ms) { free(es); free(ms); return -1; }
    memset(&mm, 0, sizeof mm); memset(mframes, 0, sizeof mframes); mm.num_frames = 1; mm.frames = mframes; mm.num_meshes = 0;
    memset(&cv_ms, 0, sizeof cv_ms); memset(&cv_mm, 0, sizeof cv_mm); memset(&cv_de, 0, sizeof cv_de);
    cv_ms.value = mirror_shadows; cv_mm.value = mirror_models; cv_de.value = drawentities;
    r_mirror_shadows = &cv_ms; r_mirror_models = &cv_mm; r_drawentities = &cv_de;
    const float keep_shadows = cv_shadows.value, keep_ps = cv_playershadow.value;
    cv_shadows.value = shadows; cv_playershadow.value = playershadow; r_mirror = mirror ? true : false;
    for (int i = 0; i < n; i++) {
        model_t *m = &ms[i];
        m->flags = model_flags[i];
        if (kinds[i] == 0) { m->type = mod_alias; m->extradata = md2_blob; m->triangles = (neighbours_t *)md2_nb; }
        else if (kinds[i] == 1) { m->type = mod_alias_md3; m->extradata = &mm; }
        else { m->type = mod_brush; m->nummodelsurfaces = 0; m->firstmodelsurface = 0; }
        es[i].model = m; es[i].flags = ent_flags[i]; es[i].frame = 0; es[i].oldframe = 0; es[i].backlerp = 0.25f;
        for (int k = 0; k < 3; k++) { es[i].origin[k] = org[k] + i; es[i].oldorigin[k] = org[k] + i; }
        cl_lightvisedicts[i] = &es[i];
    }
    cl_numlightvisedicts = n;
    set_light(lorg, 300.0f);
    sink_ents = es; sink_n = n; sink_order = order_out; sink_count = 0;
    R_DrawEntsShadowVolumes(self_shadow ? true : false, all ? true : false);
    sink_order = 0; cl_numlightvisedicts = 0;
    cv_shadows.value = keep_shadows; cv_playershadow.value = keep_ps; r_mirror = false;
    free(es); free(ms);
    return sink_count;
}

/* MD2 light-lerp family.  which: 0 = ..LightARB (LightP on TMU1, tsH on TMU3, expanded verts on
 * TMU2), 1 = ..LightARB_vp (tangent TMU1? see below).  We record every TMU; callers pick. */
int bq2ref_md2_light(void *blob, const uint8_t *tangents, const uint8_t *binormals, const uint8_t *normals,
                     int smooth, int frame, int oldframe, float backlerp,
                     const float *origin, const float *oldorigin, const float *angles,
                     const float *light_model, const float *view_model, int vp,
                     float *tmu_out /*[CAP_TMUS][3*3*T]*/, float *vertex_out, int *sizes /*[CAP_TMUS]*/)
{
    bq2ref_init();
    dmdl_t *h = (dmdl_t *)blob;
    memset(&mdl, 0, sizeof mdl); mdl.type = mod_alias; mdl.extradata = blob;
    mdl.tangents = (byte *)tangents; mdl.binormals = (byte *)binormals; mdl.normals = (byte *)normals;
    mdl.flags = smooth ? RF_SMOOTHVECS : 0; currentmodel = &mdl;
    set_entity(frame, oldframe, backlerp, origin, oldorigin, angles, 0);
    set_light(light_model, 300.0f);
    for (int k = 0; k < 3; k++) r_origin[k] = view_model[k];
    memset(cap_ptr, 0, sizeof cap_ptr); memset(cap_size, 0, sizeof cap_size); cap_vertex_ptr = 0;
    gl_state.currenttmu = -1; cap_tmu = 0;
    cap_out[0] = 0;
    for (int t = 1; t < CAP_TMUS; t++) cap_out[t] = tmu_out + (size_t)t * 9 * h->num_tris;
    cap_vertex_out = vertex_out;
    if (vp) GL_DrawAliasFrameLerpLightARB_vp(h, SHADER_ARB6_VALUE);
    else    GL_DrawAliasFrameLerpLightARB(h, SHADER_ARB6_VALUE);
    for (int t = 0; t < CAP_TMUS; t++) { cap_out[t] = 0; if (sizes) sizes[t] = cap_ptr[t] ? cap_size[t] : 0; }
    cap_vertex_out = 0;
    return cap_count;
}

/* ---- MD3 ------------------------------------------------------------------------------------- */
/* Build a maliasmodel_t around caller arrays and run R_DrawMD3AliasShadowVolume (M:63800) whole.
 * mesh_verts[m] : maliasvertex_t[num_frames*V_m]; results are gathered per mesh from inside the
 * (stubbed) glDrawArrays the function issues after each mesh. */
int bq2ref_md3_shadow(int num_frames, const float *frame_translate, int num_meshes,
                      const int *num_verts, const int *num_tris, void *const *mesh_verts,
                      void *const *mesh_indexes, void *const *mesh_neighbours, const int *casts,
                      int frame, int oldframe, float backlerp,
                      const float *origin, const float *oldorigin, const float *angles, int flags,
                      const float *light_world, float radius,
                      const int *vert_off, const int *tri_off, const int *tv_off,
                      float *lerped, float *extruded, uint8_t *viss_out, float *tris_verts, int *tris_counter)
{
    static maliasmodel_t mm; static maliasmesh_t meshes[MD3_MAX_MESHES]; static maliasframe_t frames[MD3_MAX_FRAMES];
    static image_t noshadow_img;
    bq2ref_init();
    if (num_meshes > MD3_MAX_MESHES || num_frames > MD3_MAX_FRAMES) return -2;
    memset(&mm, 0, sizeof mm); memset(meshes, 0, sizeof meshes);
    memset(&skin_img, 0, sizeof skin_img); skin_img.CastShadow = true;
    memset(&noshadow_img, 0, sizeof noshadow_img); noshadow_img.AlphaTest = true;   /* rejected by M:63878-63880 */
    for (int f = 0; f < num_frames; f++) for (int k = 0; k < 3; k++) frames[f].translate[k] = frame_translate[3*f+k];
    mm.num_frames = num_frames; mm.frames = frames; mm.num_meshes = num_meshes; mm.meshes = meshes;
    md3_nmesh_cast = 0;
    for (int m = 0; m < num_meshes; m++) {
        meshes[m].num_verts = num_verts[m]; meshes[m].num_tris = num_tris[m];
        meshes[m].vertexes = (maliasvertex_t *)mesh_verts[m];
        meshes[m].indexes = (WORD *)mesh_indexes[m];
        meshes[m].triangles = (neighbours_t *)mesh_neighbours[m];
        meshes[m].num_skins = 1;
        meshes[m].img_skins[0] = casts[m] ? &skin_img : &noshadow_img;
        md3_mesh_nverts[m] = num_verts[m]; md3_mesh_ntris[m] = num_tris[m];
        tris_counter[m] = -1;
        if (casts[m]) md3_mesh_order[md3_nmesh_cast++] = m;
    }
    memset(&mdl, 0, sizeof mdl); mdl.type = mod_alias_md3; mdl.extradata = &mm;
    set_entity(frame, oldframe, backlerp, origin, oldorigin, angles, flags);
    set_light(light_world, radius);
    md3_sink.lerped = lerped; md3_sink.extruded = extruded; md3_sink.viss = viss_out;
    md3_sink.tris_verts = tris_verts; md3_sink.tris_counter = tris_counter;
    md3_sink.vert_off = vert_off; md3_sink.tri_off = tri_off; md3_sink.tv_off = tv_off; md3_sink.drawn = 0; md3_sink.draws = 0;
    cap_mesh_hook = 1;
    R_DrawMD3AliasShadowVolume();
    cap_mesh_hook = 0;
    return md3_sink.drawn;
}

#ifdef BQ2_WITH_DROPIN
/* The drop-in shim's MD3 per-mesh body (integration/bq2_dropin.c) under the same globals the reference's
 * monolithic R_DrawMD3AliasShadowVolume sets up: light moved to model space by the reference's own lines
 * (M:63806-63820 are the same statements as M:63733-63745, exercised here through R_DrawAliasShadowVolume's
 * twin below), then one call per casting mesh. */
int  bq2_dropin_md3_mesh(maliasmodel_t *mm, int nmesh);
int bq2ref_dropin_md3(int num_frames, const float *frame_translate, int num_meshes,
                      const int *num_verts, const int *num_tris, void *const *mesh_verts,
                      void *const *mesh_indexes, void *const *mesh_neighbours, const int *casts,
                      int frame, int oldframe, float backlerp,
                      const float *origin, const float *oldorigin, const float *angles,
                      const float *light_world, float radius,
                      const int *vert_off, const int *tri_off, const int *tv_off,
                      float *lerped, float *extruded, uint8_t *viss_out, float *tris_verts, int *tris_counter)
{
    static maliasmodel_t mm; static maliasmesh_t meshes[MD3_MAX_MESHES]; static maliasframe_t frames[MD3_MAX_FRAMES];
    bq2ref_init();
    memset(&mm, 0, sizeof mm); memset(meshes, 0, sizeof meshes);
    for (int f = 0; f < num_frames; f++) for (int k = 0; k < 3; k++) frames[f].translate[k] = frame_translate[3*f+k];
    mm.num_frames = num_frames; mm.frames = frames; mm.num_meshes = num_meshes; mm.meshes = meshes;
    for (int m = 0; m < num_meshes; m++) {
        meshes[m].num_verts = num_verts[m]; meshes[m].num_tris = num_tris[m];
        meshes[m].vertexes = (maliasvertex_t *)mesh_verts[m]; meshes[m].indexes = (WORD *)mesh_indexes[m];
        meshes[m].triangles = (neighbours_t *)mesh_neighbours[m];
    }
    memset(&mdl, 0, sizeof mdl); mdl.type = mod_alias_md3; mdl.extradata = &mm; currentmodel = &mdl;
    set_entity(frame, oldframe, backlerp, origin, oldorigin, angles, 0);
    set_light(light_world, radius);
    {   /* light -> model space with the reference's AngleVectors */
        vec3_t t, f, r, u;
        for (int k = 0; k < 3; k++) light.origin[k] = light.origin[k] - ent.origin[k];
        if (ent.angles[0] || ent.angles[1] || ent.angles[2]) {
            AngleVectors(ent.angles, f, r, u);
            for (int k = 0; k < 3; k++) t[k] = light.origin[k];
            light.origin[0] = t[0]*f[0] + t[1]*f[1] + t[2]*f[2];
            light.origin[1] = -(t[0]*r[0] + t[1]*r[1] + t[2]*r[2]);
            light.origin[2] = t[0]*u[0] + t[1]*u[1] + t[2]*u[2];
        }
    }
    int done = 0;
    for (int m = 0; m < num_meshes; m++) {
        tris_counter[m] = -1;
        if (!casts[m]) continue;                       /* the skin filter M:63862-63880 stays with the caller */
        if (bq2_dropin_md3_mesh(&mm, m) != 0) return -1;
        memcpy(lerped + 3 * vert_off[m], tempVertexArray, sizeof(float) * 3 * num_verts[m]);
        memcpy(extruded + 3 * vert_off[m], extrudedVerts, sizeof(float) * 3 * num_verts[m]);
        memcpy(viss_out + tri_off[m], viss, num_tris[m]);
        memcpy(tris_verts + 3 * tv_off[m], TrisVerts, sizeof(float) * 3 * TrisCounter);
        tris_counter[m] = TrisCounter; done++;
    }
    return done;
}
#endif

/* MD3 light family, one mesh.  The reference lerps the mesh's verts in R_DrawMD3AliasObjectLight
 * (M:66892-66896) before calling these; callers pass the lerped verts they want used.
 * vp=0: TMU1 = LightP, TMU3 = tsH (M:65839).  vp=1: TMU1 = binormal, TMU2 = normal, TMU3 = tangent
 * (M:65671).  tmu_out = [CAP_TMUS][3*V]. */
int bq2ref_md3_light(void *mesh_verts, int num_verts, int frame, int oldframe, float backlerp,
                     const float *lerped_in, const float *light_model, const float *view_model, int vp,
                     float *tmu_out, int *sizes)
{
    static maliasmodel_t mm; static maliasmesh_t mesh;
    static const float zero[3] = {0, 0, 0};
    bq2ref_init();
    memset(&mm, 0, sizeof mm); memset(&mesh, 0, sizeof mesh);
    mesh.num_verts = num_verts; mesh.vertexes = (maliasvertex_t *)mesh_verts; mesh.num_tris = 0;
    mm.num_meshes = 1; mm.meshes = &mesh;
    memset(&mdl, 0, sizeof mdl); mdl.type = mod_alias_md3; mdl.extradata = &mm; currentmodel = &mdl;
    set_entity(frame, oldframe, backlerp, zero, zero, zero, 0);
    set_light(light_model, 300.0f);
    for (int k = 0; k < 3; k++) r_origin[k] = view_model[k];
    memcpy(tempVertexArray, lerped_in, sizeof(float) * 3 * num_verts);
    memset(cap_ptr, 0, sizeof cap_ptr); memset(cap_size, 0, sizeof cap_size);
    gl_state.currenttmu = -1; cap_tmu = 0;
    cap_out[0] = 0;
    for (int t = 1; t < CAP_TMUS; t++) cap_out[t] = tmu_out + (size_t)t * 3 * num_verts;
    cap_elems_nverts = num_verts;
    if (vp) GL_DrawMD3AliasFrameLerpLightARB_vp(&mesh, SHADER_ARB6_VALUE, &mm, 0);
    else    GL_DrawMD3AliasFrameLerpLightARB(&mesh, SHADER_ARB6_VALUE, 0);
    for (int t = 0; t < CAP_TMUS; t++) { cap_out[t] = 0; if (sizes) sizes[t] = cap_ptr[t] ? cap_size[t] : 0; }
    return num_verts;
}

/* ---- brush-model dynamic volumes ------------------------------------------------ */
extern int       r_lightTimestamp;                               /* data.h:2716 */
extern vec3_t    vcache[];                                       /* data.h:2787 */
extern unsigned  icache[];                                       /* data.h:2788 */
void R_DrawBrushModelVolumes(void);                              /* M:63326 */

/* R_DrawBrushModelVolumes whole on a brush model described by flat arrays: polygon p has numverts[p] vertices at
 * verts[first[p]..], neighbours[first[p] + j] = polygon across edge j or -1, lit[p] = the light's marking pass
 * stamped it (lightTimestamp == r_lightTimestamp, done elsewhere in the engine), fence[p] = its texture is a
 * non-casting SURF_FENCE one (M:63387-63389).  Returns ib (the index count handed to glDrawElements, 0 when
 * nothing is drawn); *vb_out = vertex pairs written to vcache. */
int bq2ref_brush_volume(int n_polys, const int *numverts, const float *verts, const int *neighbours,
                        const uint8_t *lit, const uint8_t *fence,
                        const float *origin, const float *angles, const float *light_world, float radius,
                        float *vcache_out, unsigned *icache_out, int *vb_out)
{
    static image_t img_cast, img_fence; static mtexinfo_t ti_cast, ti_fence;
    bq2ref_init();
    glpoly_t **polys = (glpoly_t **)calloc((size_t)n_polys, sizeof *polys);
    msurface_t *surfs = (msurface_t *)calloc((size_t)n_polys, sizeof *surfs);
    memset(&img_cast, 0, sizeof img_cast); img_cast.CastShadow = true;
    memset(&img_fence, 0, sizeof img_fence); img_fence.CastShadow = false;
    memset(&ti_cast, 0, sizeof ti_cast); ti_cast.image = &img_cast;
    memset(&ti_fence, 0, sizeof ti_fence); ti_fence.image = &img_fence; ti_fence.flags = SURF_FENCE;
    r_lightTimestamp = 7;
    int first = 0, vb = 0;
    for (int p = 0; p < n_polys; p++) {
        int n = numverts[p];
        glpoly_t *g = (glpoly_t *)calloc(1, sizeof(glpoly_t) + sizeof(float) * VERTEXSIZE * (size_t)(n > 4 ? n - 4 : 0));
        g->numverts = n; g->lightTimestamp = lit[p] ? r_lightTimestamp : 3;
        g->neighbours = (glpoly_t **)calloc((size_t)n, sizeof(glpoly_t *));
        for (int j = 0; j < n; j++) for (int k = 0; k < 3; k++) g->verts[j][k] = verts[3 * (first + j) + k];
        polys[p] = g;
        surfs[p].polys = g; surfs[p].numedges = n; surfs[p].texinfo = fence[p] ? &ti_fence : &ti_cast;
        if (lit[p] && !fence[p]) vb += n;
        first += n;
    }
    first = 0;
    for (int p = 0; p < n_polys; p++) {
        for (int j = 0; j < numverts[p]; j++) { int nb = neighbours[first + j]; polys[p]->neighbours[j] = nb < 0 ? NULL : polys[nb]; }
        first += numverts[p];
    }
    memset(&mdl, 0, sizeof mdl); mdl.type = mod_brush; mdl.surfaces = surfs; mdl.firstmodelsurface = 0; mdl.nummodelsurfaces = n_polys;
    static const float zero[3] = {0, 0, 0};
    set_entity(0, 0, 0.0f, origin, zero, angles, 0);
    set_light(light_world, radius);
    cap_elems_count = 0;
    R_DrawBrushModelVolumes();
    int ib = cap_elems_count;
    memcpy(vcache_out, vcache, sizeof(float) * 3 * 2 * (size_t)vb);
    memcpy(icache_out, icache, sizeof(unsigned) * (size_t)ib);
    *vb_out = vb;
    for (int p = 0; p < n_polys; p++) { free(polys[p]->neighbours); free(polys[p]); }
    free(polys); free(surfs);
    return ib;
}

/* ---- precompute ------------------------------------------------------------------------------ */
/* M:51643-51653 restated as a CALL sequence around the reference's findneighbour. */
void bq2ref_md2_neighbours(void *blob, int32_t *out)
{
    dmdl_t *h = (dmdl_t *)blob;
    dtriangle_t *tris = (dtriangle_t *)((byte *)h + h->ofs_tris);
    neighbours_t *nb = (neighbours_t *)out;
    for (int i = 0; i < h->num_tris; i++) for (int j = 0; j < 3; j++) nb[i].neighbours[j] = -1;
    for (int i = 0; i < h->num_tris; i++)
        for (int j = 0; j < 3; j++)
            if (nb[i].neighbours[j] == -1)
                nb[i].neighbours[j] = findneighbour(i, j, h->num_tris, nb, tris);
}
void bq2ref_md3_neighbours(uint16_t *indexes, int num_tris, int32_t *out)
{ R_BuildTriangleNeighbors((neighbours_t *)out, (WORD *)indexes, num_tris); }
int  bq2ref_normal2index(const float *v) { vec3_t t = { v[0], v[1], v[2] }; return Normal2Index(t); }
void bq2ref_vecs_for_tris(float *v0, float *v1, float *v2, float *st0, float *st1, float *st2, float *tan, float *bin)
{ VecsForTris(v0, v1, v2, st0, st1, st2, tan, bin); }
void bq2ref_tangent_for_tri_md3(uint16_t *index, void *verts, float *st, float *tan, float *bin)
{ TangentForTrimd3((WORD *)index, (maliasvertex_t *)verts, (maliascoord_t *)st, tan, bin); }
unsigned Com_BlockChecksum(void *buffer, int length);                 /* M:40589 */
unsigned bq2ref_block_checksum(void *buffer, int length) { return Com_BlockChecksum(buffer, length); }
void bq2ref_angle_vectors(float *angles, float *f, float *r, float *u) { AngleVectors(angles, f, r, u); }
float bq2ref_vector_normalize(float *v) { return VectorNormalize(v); }

/* ---- the engine's model loaders, run whole ----------------------------------------------------
 * Mod_LoadAliasModel (M:51217) and Mod_LoadAliasMD3Model (M:50584) with cache = true (no skins are looked up):
 * file image -> in-memory image, neighbour tables, the assembled tangent / binormal / normal loops
 * (M:51661-51804, 50956-51015) and the cache file they write under <gamedir>/cache/.  The caller passes a scratch
 * directory; the loader runs with it as working directory, "./baseq2" as game directory and search path, so a
 * models/.../tris.fx2 placed there ("unsmoothvecs") is honoured as in the engine (M:51363-51375). */
#include <unistd.h>
#include <stdlib.h>
extern searchpath_t *fs_searchpaths;
extern char          fs_gamedir[MAX_OSPATH];
extern cvar_t       *r_simple, *developer;
extern int           modfilelen;
extern struct zlink { struct zlink *prev, *next; } z_chain;   /* the list head of M:2053-2061 (type private to main.c): its two links */
bool  Mod_LoadAliasModel(model_t *, void *, float, bool, int, bool);
void  Mod_LoadAliasMD3Model(model_t *, void *, float, bool, bool);
void *Hunk_Begin(int, char *);
void  Z_Free(void *);
void  Swap_Init(void);

static model_t       ld_mod;
static void         *ld_base;
static searchpath_t  ld_search;
static cvar_t        cv_simple;

static int loader_enter(const char *workdir, const char *modname, char *cwd, size_t cwd_len)
{
    bq2ref_init();
    if (!getcwd(cwd, cwd_len) || chdir(workdir)) return -1;
    memset(&cv_simple, 0, sizeof cv_simple); r_simple = &cv_simple; developer = NULL;
    if (!z_chain.next) z_chain.next = z_chain.prev = &z_chain;      /* Qcommon_Init M:87900 */
    Swap_Init();                                                    /* LittleLong & co. are function pointers, M:2227 */
    strcpy(fs_gamedir, "./" BASEDIRNAME);
    memset(&ld_search, 0, sizeof ld_search); strcpy(ld_search.filename, "./" BASEDIRNAME);
    fs_searchpaths = &ld_search;
    if (ld_base) { free(ld_base); ld_base = NULL; }
    memset(&ld_mod, 0, sizeof ld_mod);
    strncpy(ld_mod.name, modname, sizeof ld_mod.name - 1);
    ld_base = Hunk_Begin(64 << 20, ld_mod.name);            /* as Mod_ForName does before it calls a loader (M:27000-27110) */
    ld_mod.extradata = ld_base;
    return 0;
}

/* returns 0, or -1 when the loader refused; *smooth = RF_SMOOTHVECS after the .fx2 was parsed.  image: ofs_end bytes. */
int bq2ref_load_md2(void *file, int file_len, const char *workdir, const char *modname, float scale, int invert,
                    void *image, int32_t *neighbours, uint8_t *tangents, uint8_t *binormals, uint8_t *normals, int *smooth)
{
    char cwd[4096];
    if (loader_enter(workdir, modname, cwd, sizeof cwd)) return -2;
    modfilelen = file_len;
    bool ok = Mod_LoadAliasModel(&ld_mod, file, scale, invert != 0, 0, true);
    if (chdir(cwd)) return -2;
    if (ld_mod.st) { Z_Free(ld_mod.st); ld_mod.st = NULL; }
    if (!ok) return -1;
    dmdl_t *h = (dmdl_t *)ld_base;
    *smooth = (ld_mod.flags & RF_SMOOTHVECS) != 0;
    size_t rows = (size_t)(*smooth ? h->num_xyz : h->num_tris) * (size_t)h->num_frames;
    memcpy(image, h, (size_t)h->ofs_end);
    memcpy(neighbours, ld_mod.triangles, sizeof(neighbours_t) * (size_t)h->num_tris);
    memcpy(tangents, ld_mod.tangents, rows);
    memcpy(binormals, ld_mod.binormals, rows);
    if (!*smooth) memcpy(normals, ld_mod.normals, rows);
    return 0;
}

/* loads the whole .md3; returns the number of meshes, the model stays loaded for bq2ref_loaded_md3_mesh */
int bq2ref_load_md3(void *file, int file_len, const char *workdir, const char *modname, float scale, int invert,
                    int *num_frames)
{
    char cwd[4096];
    if (loader_enter(workdir, modname, cwd, sizeof cwd)) return -2;
    modfilelen = file_len;                                   /* the cache files carry the checksum of the whole file */
    Mod_LoadAliasMD3Model(&ld_mod, file, scale, invert != 0, true);
    if (chdir(cwd)) return -2;
    maliasmodel_t *mm = (maliasmodel_t *)ld_base;
    *num_frames = mm->num_frames;
    return mm->num_meshes;
}
int bq2ref_loaded_md3_dims(int mesh, int *num_verts, int *num_tris)
{
    maliasmodel_t *mm = (maliasmodel_t *)ld_base;
    if (!mm || mesh < 0 || mesh >= mm->num_meshes) return -1;
    *num_verts = mm->meshes[mesh].num_verts; *num_tris = mm->meshes[mesh].num_tris;
    return 0;
}
int bq2ref_loaded_md3_mesh(int mesh, void *verts, uint16_t *indexes, int32_t *neighbours, float *st, float *frame_translate)
{
    maliasmodel_t *mm = (maliasmodel_t *)ld_base;
    if (!mm || mesh < 0 || mesh >= mm->num_meshes) return -1;
    maliasmesh_t *m = &mm->meshes[mesh];
    memcpy(verts, m->vertexes, sizeof(maliasvertex_t) * (size_t)mm->num_frames * (size_t)m->num_verts);
    memcpy(indexes, m->indexes, sizeof(WORD) * 3 * (size_t)m->num_tris);
    memcpy(neighbours, m->triangles, sizeof(neighbours_t) * (size_t)m->num_tris);
    memcpy(st, m->stcoords, sizeof(maliascoord_t) * (size_t)m->num_verts);
    for (int f = 0; f < mm->num_frames; f++)
        for (int k = 0; k < 3; k++) frame_translate[3 * f + k] = mm->frames[f].translate[k];
    return 0;
}
