/*
 * bq2_api.cu -- host side of the C ABI in include/bq2_gpu.h: context, model residency in HBM,
 * per-frame record packing, launch, read-back.  C++ inside, extern "C" outside.
 *
 * What lives in HBM (bq2_internal.h has the exact layouts): per mesh the vertex frames, the
 * 12-byte/triangle index+neighbour table and the tangent-space byte arrays, allocated from
 * 64 MiB slabs; per frame one packed record buffer (entity slots then pair slots, one H2D copy)
 * and the output arrays, grown geometrically and reused across frames.
 *
 * There is no CPU fallback anywhere in this file: every entry point either runs the CUDA path or
 * returns an error.
 */
#include <cuda_runtime.h>
#include <cuda_profiler_api.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include <algorithm>
#include <atomic>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include "bq2_gpu.h"
#include "bq2_internal.h"

#include <time.h>
#if defined(__linux__)
#include <sched.h>
#endif
namespace {

/* BQ2_TRACE=1: host time per phase of bq2_frame_submit, printed by bq2_destroy (diagnostics only) */
struct Trace {
    bool on = getenv("BQ2_TRACE") != nullptr;
    bool gpu = on && strcmp(getenv("BQ2_TRACE"), "host") != 0;   /* BQ2_TRACE=host: host laps only, frames stay pipelined */
    double acc[8] = {0}; long n = 0; double t0 = 0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; }
    void start() { if (on) t0 = now(); }
    void lap(int i) { if (on) { double t = now(); if (n > 64) acc[i] += t - t0; t0 = t; } }   /* first frames allocate */
};

std::string g_create_error;

struct Slab { unsigned char *base; size_t size, used; };

struct Model {
    bool md3 = false;
    int  F = 0;
    int  first_mesh = 0, n_mesh = 0;
    uint32_t flags = 0;
    uint32_t rf_flags = 0;       /* model_t.flags bits of the caller loop's filter (bq2_model_set_rf_flags) */
    std::vector<int> V, T, casts;
    std::vector<float> hdr;      /* MD2: F x {scale[3], translate[3]} (daliasframe_t header);
                                    MD3: F x translate[3] (maliasframe_t.translate)            */
    const float *d_hdr = nullptr;/* the same on the device (world path: lerp constants are computed there) */
};

template <typename T> struct DevBuf {
    T *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        size_t want = std::max(n, cap + cap / 2);
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    unsigned char *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        size_t want = std::max(n, cap + cap / 2);
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost((void **)&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

/* everything that belongs to ONE frame */
#define BQ2_DIRECT_MIN_BYTES (1u << 20)
#define BQ2_TAIL_STORE_MAX_WORDS 16384u      /* per-pair counts beyond 64 KB go through the copy engine */
/* Up to sixteen frames in flight: a world frame is a chain of four dependent stream operations (copy in, record-building
   kernel, frame kernel, a last kernel that hands the counts to the host) whose END-TO-END LATENCY is ~40 us on a B200
   (profiles/r02_chain_latency.md), so the frame rate of a pipelined caller is latency / frames in flight until the host
   or the SMs become the limit: C2 measured 30.0 / 16.9 / 12.7 / 12.0 us per frame with 2 / 4 / 8 / 16 in flight.
   States are taken on demand: a caller pays for the depth it uses. */
#ifndef BQ2_FRAME_STATES        /* (overridable for experiments: make EXTRA="-DBQ2_FRAME_STATES=5 -DBQ2_MAX_IN_FLIGHT=4") */
#define BQ2_FRAME_STATES   17
#define BQ2_MAX_IN_FLIGHT  16
#endif
static_assert(BQ2_MAX_IN_FLIGHT < BQ2_FRAME_STATES, "the state bq2_frame_wait returned last must not be the one taken next");
struct FrameState {
    cudaStream_t stream = nullptr;
    std::vector<int32_t>  ent_slot_first, pair_slot_first;
    std::vector<uint32_t> ent_vert_off, pair_vert_off, pair_bits_off, ent_tb_off, pair_ts_off;
    std::vector<uint32_t> cursor;                                /* scratch for the counting sort */
    std::vector<int32_t>  slot_mesh_V, slot_mesh_T;              /* per pair slot, for expand    */
    std::vector<int32_t>  pair_slot_entslot;
    std::vector<int32_t>  slot_record;                           /* entity slot -> record position */
    std::vector<uint32_t> world_pvoff, world_pboff;
    PinBuf hrec, hcnt;
    struct Meta { bool world = false, tables = false; int32_t n_ent_slots = 0, n_pair_slots = 0; uint64_t total_verts = 0; uint32_t stride = 0; } meta;
    /* A frame is ~6 stream operations.  When two consecutive frames on the same state have the same shape
       and buffers, the operations are captured once into a CUDA graph and replayed with one call. */
    struct GraphKey { int32_t v[20]; const void *p[24]; };
    struct GraphCache { GraphKey key{}; bool have_key = false; cudaGraphExec_t exec = nullptr; uint64_t kernels = 0; } gcache;
    DevBuf<unsigned char> d_records;
    DevBuf<float> lerped, extruded, normals, tangents, binormals, lightp, tsh, tri_normals;
    DevBuf<uint32_t> facing_bits, indices, index_count;
    DevBuf<unsigned char> d_world;                               /* world path: dpairs | buckets | overflow chain */
    bq2_world_launch wlaunch{};
    bool world_mode = false, world_tables = false;
    /* The last kernel of a world frame hands the counts to the host and clears what the next frame accumulates into
       (totals, pairs per record, chain heads): as long as the frames on this state keep their shape and buffer, no
       clear is needed in front of them.  tail_clean: that holds for (clean_ptr, clean_nP, clean_nE). */
    bool tail_clean = false; const void *clean_ptr = nullptr; int clean_nP = -1, clean_nE = -1;
    uint32_t *hcnt_dev = nullptr; const void *hcnt_dev_of = nullptr;     /* device alias of hcnt (mapped page-locked memory) */
    uint32_t *keep_cnt = nullptr;                                /* pairs per record of the finished frame (replay, hash) */
    /* what the stream operations of a world frame need, so that they can be enqueued by the ctx's submit worker
       (bq2_world_submit returns after the per-entity loop; bq2_frame_wait meets the worker before it meets the GPU) */
    struct WorldJob {
        const void *lights = nullptr, *pair_ent = nullptr, *pair_light = nullptr;   /* caller arrays the worker still reads */
        bool direct = false, stage = false, use_key = false, clean = false;
        size_t off_xf = 0, off_lights = 0, off_pe = 0, off_pl = 0, off_b = 0, basis_bytes = 0, copy_end = 0, tot_word = 0, cnt_word = 0;
        int nE = 0, nL = 0, nP = 0, n_rot = 0; uint32_t want = 0;
        GraphKey key{};
        bq2_status status = BQ2_OK; std::string err;
    } job;
    std::atomic<uint64_t> job_posted{0}, job_done{0};
    bq2_launch launch{};
    /* entity-slot records are grouped by kernel: [0, n_warp) -> warp kernel, [n_warp, n_ent_slots) -> team kernel */
    struct KLaunch { int n = 0, max_V = 0, max_T = 0, any_md2 = 0; } k_warp, k_team;
    int max_V = 0, max_T = 0, any_md2 = 0;
    int32_t n_ent_slots = 0, n_pair_slots = 0, n_ents = 0, n_pairs = 0;
    uint64_t total_verts = 0;
    bq2_frame_out out{};
    void release() {
        if (gcache.exec) cudaGraphExecDestroy(gcache.exec);
        gcache.exec = nullptr;
        d_records.release(); lerped.release(); extruded.release(); normals.release(); tangents.release(); binormals.release();
        lightp.release(); tsh.release(); tri_normals.release(); facing_bits.release(); indices.release(); index_count.release();
        d_world.release(); hrec.release(); hcnt.release();
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
};

} // namespace

struct bq2_resident {                   /* bq2_frame_keep: a frame's device records and launch parameters */
    DevBuf<unsigned char> records;
    int state = 0;
    bq2_launch L{};
    FrameState::KLaunch k_warp, k_team;
    /* bq2_frame_keep_own: the frame's result arrays are its own (a ring of such frames writes a different set of
       output lines on every step instead of overwriting one set that stays in L2) */
    bool own = false;
    DevBuf<float> lerped, extruded, normals, tangents, binormals, lightp, tsh, tri_normals;
    DevBuf<uint32_t> facing_bits, indices, index_count;
    size_t result_bytes = 0;
    void release() {
        records.release(); tri_normals.release(); lerped.release(); extruded.release(); normals.release(); tangents.release(); binormals.release();
        lightp.release(); tsh.release(); facing_bits.release(); indices.release(); index_count.release();
    }
};

struct bq2_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;                               /* = fs[0].stream: uploads, load-time kernels, brush pairs */
    std::string err;
    std::atomic<uint64_t> launches{0};
    Trace trace;
    /* submit worker: one thread per ctx that issues the stream operations of world frames (staging copy of the caller's
       page-locked arrays, graph launch), so that the caller's thread pays for the per-entity loop only */
    std::thread worker;
    std::mutex wmx; std::condition_variable wcv;
    std::atomic<int> wq[BQ2_FRAME_STATES + 1];              /* ring of frame-state numbers */
    std::atomic<uint64_t> wq_head{0}, wq_tail{0};
    std::atomic<bool> wstop{false}, wsleeping{false};
    bool use_worker = getenv("BQ2_NO_SUBMIT_THREAD") == nullptr;
    double last_submit_us = -1e18;

    std::vector<Slab> slabs;
    std::vector<Model> models;
    std::vector<bq2_dev_mesh> meshes;           /* host mirror of the device mesh table */
    DevBuf<bq2_dev_mesh> d_meshes;
    std::vector<bq2_dev_model> dev_models;      /* host mirror of the device model table */
    DevBuf<bq2_dev_model> d_models;
    bool meshes_dirty = false;

    std::vector<bq2_model_row> model_rows;                       /* flat mirror of models for the C entity loop */
    std::vector<float>    ent_basis;                             /* build_records scratch: f, r, u, state */
    std::vector<bq2_entity> tmp_ents; std::vector<bq2_pair> tmp_pairs;     /* host-built fallback of the world path */
    /* Complete frame states (host tables, pinned buffers, device records AND device result arrays, stream, event,
       graph) used as a ring: frames i+1, i+2 are packed, copied and computed while frame i is still on the GPU or
       being consumed.  A submit takes the next state of the ring when a frame is in flight, the same one otherwise;
       with at most BQ2_MAX_IN_FLIGHT (sixteen) in flight the state bq2_frame_wait returned last is never the one taken next. */
    FrameState fs[BQ2_FRAME_STATES];
    int cur = 0;                                                 /* state of the most recently submitted frame */
    uint64_t submit_seq = 0, wait_seq = 0;
    int last_waited = 0;
    bool busy[BQ2_FRAME_STATES] = {};                            /* submitted, not yet waited for */
    int  order[BQ2_FRAME_STATES] = {};                           /* states of the frames in flight, oldest first (ring) */
    bool use_graphs = getenv("BQ2_NO_GRAPH") == nullptr;
    bool pdl_world = getenv("BQ2_PDL_WORLD") != nullptr && atoi(getenv("BQ2_PDL_WORLD")) != 0;
    bool submitted = false;
    struct HostRange { const unsigned char *base; size_t bytes; };
    std::vector<HostRange> host_ranges;                          /* page-locked memory handed out by bq2_host_alloc */
    size_t direct_min_bytes = getenv("BQ2_DIRECT_MIN") ? (size_t)atoll(getenv("BQ2_DIRECT_MIN")) : BQ2_DIRECT_MIN_BYTES;   /* tests set 0 */
    /* brush models (8f.4) */
    struct Brush { int n_polys = 0, n_slots = 0, worst_indices = 0; };
    std::vector<Brush> brushes; std::vector<bq2_brush_dev_h> brush_dev; DevBuf<bq2_brush_dev_h> d_brushes;
    DevBuf<unsigned char> d_brush_in; DevBuf<float> brush_vcache; DevBuf<uint32_t> brush_icache, brush_counts;
    PinBuf h_brush, h_brush_lit;                                 /* page-locked: pair records | counts; staged lit bytes */
    const uint32_t *h_brush_counts = nullptr; uint32_t brush_vstride = 0, brush_istride = 0; int brush_pairs = 0;
};

namespace {

/* kernels issued while a stream is being captured are not launched: they are counted aside and added per graph launch */
thread_local uint64_t *tl_capture_count = nullptr;
#define COUNT_LAUNCH(ctx, n) do { if (tl_capture_count) *tl_capture_count += (uint64_t)(n); else (ctx)->launches += (uint64_t)(n); } while (0)
/* the submit worker (below) reports into its job: ctx->err belongs to the caller's thread */
thread_local std::string *tl_err_sink = nullptr;
bq2_status fail(bq2_ctx *c, bq2_status s, const char *what)
{ if (tl_err_sink) *tl_err_sink = what; else if (c) c->err = what; else g_create_error = what; return s; }
bq2_status fail_cuda(bq2_ctx *c, cudaError_t e, const char *where)
{
    std::string m = std::string(where) + ": " + cudaGetErrorString(e);
    if (tl_err_sink) *tl_err_sink = m; else if (c) c->err = m; else g_create_error = m;
    cudaGetLastError();                      /* a launch error must not be reported again by the next call */
    return BQ2_E_CUDA;
}
#define CK(call, where) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail_cuda(ctx, e_, where); } while (0)

/* bump allocation out of 64 MiB slabs; 256-byte aligned so every array is TMA/128-bit friendly */
/* The C ABI never lets a C++ exception out (std::bad_alloc from a host-side table is the one that can occur):
   every entry point that builds host tables is a function-try-block ending here. */
bq2_status guard_exception(bq2_ctx *ctx) noexcept
{
    bq2_status s = BQ2_E_INVALID;
    const char *what = "internal error (C++ exception)";
    try { throw; }
    catch (const std::bad_alloc &) { s = BQ2_E_NOMEM; what = "out of host memory"; }
    catch (...) {}
    try { if (ctx) ctx->err = what; } catch (...) {}
    return s;
}

/* The state a submit takes: the one in use when nothing is in flight (a caller that never overlaps frames lives in one
   state), else the lowest-numbered state that is neither in flight nor the one bq2_frame_wait returned last (its results
   stay valid until the next wait).  A caller with d frames in flight thus touches d + 1 states, not the whole ring --
   every state keeps result arrays of the largest frame it has seen. */
int pick_state(const bq2_ctx *ctx)
{
    if (ctx->submit_seq == ctx->wait_seq) return ctx->cur;
    for (int k = 0; k < BQ2_FRAME_STATES; k++)
        if (!ctx->busy[k] && k != ctx->last_waited) return k;
    return -1;                                                    /* cannot happen: at most BQ2_MAX_IN_FLIGHT are busy */
}

bq2_status arena_alloc(bq2_ctx *ctx, size_t bytes, unsigned char **out)
{
    const size_t kSlab = 64u << 20;
    bytes = (bytes + 255) & ~(size_t)255;
    if (!ctx->slabs.empty()) {
        Slab &s = ctx->slabs.back();
        if (s.used + bytes <= s.size) { *out = s.base + s.used; s.used += bytes; return BQ2_OK; }
    }
    Slab s; s.size = std::max(kSlab, bytes); s.used = bytes;
    CK(cudaMalloc((void **)&s.base, s.size), "cudaMalloc(model slab)");
    ctx->slabs.push_back(s);
    *out = s.base;
    return BQ2_OK;
}

/* frame headers of a model to the device */
bq2_status upload_hdr(bq2_ctx *ctx, Model &mo)
{
    unsigned char *d = nullptr;
    bq2_status s = arena_alloc(ctx, mo.hdr.size() * sizeof(float), &d);
    if (s != BQ2_OK) return s;
    CK(cudaMemcpyAsync(d, mo.hdr.data(), mo.hdr.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpy(frame headers)");
    CK(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize(frame headers)");
    mo.d_hdr = (const float *)d;
    return BQ2_OK;
}

struct md2_header {
    int32_t ident, version, skinwidth, skinheight, framesize;
    int32_t num_skins, num_xyz, num_st, num_tris, num_glcmds, num_frames;
    int32_t ofs_skins, ofs_st, ofs_tris, ofs_frames, ofs_glcmds, ofs_end;
};

inline uint32_t up16(uint32_t x) { return (x + 15u) & ~15u; }

/* The reference's load-time checks on a dmdl_t image (M:51247-51264) plus the bounds of the two sections this
   library reads; shared by every entry point that takes an MD2 blob.  Returns NULL when the image is fine. */
const char *md2_header_problem(const md2_header *h, size_t blob_bytes, bq2_status *st)
{
    const int V = h->num_xyz, T = h->num_tris, F = h->num_frames;
    *st = BQ2_E_INVALID;
    if (h->version != 8) return "wrong version number";
    if (V <= 0) return "model has no vertices";
    if (T <= 0) return "model has no triangles";
    if (F <= 0) return "model has no frames";
    *st = BQ2_E_LIMIT;
    if (V > BQ2_MD2_MAX_VERTS) return "model has too many vertices";
    if (T > BQ2_MD2_MAX_TRIANGLES) return "model has too many triangles";
    *st = BQ2_E_INVALID;
    if (h->framesize < 40 + 4 * V || h->ofs_tris < 0 || h->ofs_frames < 0 ||
        (size_t)h->ofs_tris + 12u * (size_t)T > blob_bytes ||
        (size_t)h->ofs_frames + (size_t)F * (size_t)h->framesize > blob_bytes)
        return "offsets run past the blob";
    return nullptr;
}

/* index + neighbour table -> Tp records {a, b, c, n0+1} (8 B) followed by Tp records {n1+1, n2+1} (4 B);
   neighbour + 1 so that 0 means "open edge" and no sign extension is needed on the device */
void pack_tri_table(int T, int Tp, const int32_t *idx /*[T][3]*/, const int32_t *nb /*[T][3] or null*/, int16_t *dst)
{
    uint16_t *t4 = (uint16_t *)dst, *n2 = t4 + 4 * (size_t)Tp;
    memset(dst, 0, sizeof(int16_t) * 6 * (size_t)Tp);
    for (int t = 0; t < T; t++) {
        t4[4 * t + 0] = (uint16_t)idx[3 * t]; t4[4 * t + 1] = (uint16_t)idx[3 * t + 1]; t4[4 * t + 2] = (uint16_t)idx[3 * t + 2];
        t4[4 * t + 3] = (uint16_t)(nb ? nb[3 * t] + 1 : 0);
        n2[2 * t + 0] = (uint16_t)(nb ? nb[3 * t + 1] + 1 : 0);
        n2[2 * t + 1] = (uint16_t)(nb ? nb[3 * t + 2] + 1 : 0);
    }
}

/* the frame's kernels: warp-per-pair over the small-mesh records, CTA-per-pair over the rest */
bq2_status launch_kernels(bq2_ctx *ctx, FrameState &F, int overlap = 0)
{
    bq2_launch L = F.launch;
    L.seq = (uint32_t)ctx->launches;
    if (F.k_warp.n) {
        L.n_ent_slots = F.k_warp.n;
        CK((cudaError_t)bq2_launch_frame_warp(&L, F.k_warp.max_V, F.k_warp.max_T, F.k_warp.any_md2, F.stream, overlap),
           "launch(bq2_frame_warp_kernel)");
        COUNT_LAUNCH(ctx, 1);
    }
    if (F.k_team.n) {
        L.ents = F.launch.ents + F.k_warp.n;
        L.rec_base = (uint32_t)F.k_warp.n;
        L.n_ent_slots = F.k_team.n;
        CK((cudaError_t)bq2_launch_frame(&L, F.k_team.max_V, F.k_team.max_T, F.k_team.any_md2, F.stream, overlap),
           "launch(bq2_frame_kernel)");
        COUNT_LAUNCH(ctx, 1);
    }
    return BQ2_OK;
}

/* Enqueue a frame's stream operations on its state's stream.  A frame whose shape and buffers (the key) equal those of
   the previous frame on this state goes through a CUDA graph: the second such frame is captured while it is enqueued,
   later ones are one cudaGraphLaunch.  Anything else -- and every frame once a capture has failed -- is enqueued
   directly.  enqueue() issues the operations and returns a status. */
template <typename Enqueue>
bq2_status enqueue_frame(bq2_ctx *ctx, FrameState &F, const FrameState::GraphKey *key, Enqueue enqueue)
{
    if (key && ctx->use_graphs) {
        FrameState::GraphCache &g = F.gcache;
        const bool same = g.have_key && memcmp(&g.key, key, sizeof *key) == 0;
        if (same && g.exec) {
            CK(cudaGraphLaunch(g.exec, F.stream), "cudaGraphLaunch(frame)");
            ctx->launches += g.kernels;
            return BQ2_OK;
        }
        if (same) {                                      /* second frame of this shape: capture it */
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(F.stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                uint64_t captured = 0;
                tl_capture_count = &captured;
                bq2_status es = enqueue();
                tl_capture_count = nullptr;
                cudaError_t ce = cudaStreamEndCapture(F.stream, &graph);
                if (es == BQ2_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess) {
                    g.kernels = captured;
                    cudaGraphDestroy(graph);
                    CK(cudaGraphLaunch(g.exec, F.stream), "cudaGraphLaunch(frame)");
                    ctx->launches += g.kernels;
                    return BQ2_OK;
                }
                if (graph) cudaGraphDestroy(graph);      /* capture is an optimisation only */
                g.exec = nullptr; ctx->use_graphs = false;
                cudaGetLastError();
            }
        } else {
            if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
            g.key = *key; g.have_key = true;
        }
    }
    return enqueue();
}

/* BQ2_TRACE: GPU time of the four stages of a world frame, direct launches, one frame at a time (diagnostics) */
struct StageTimer {
    cudaEvent_t ev[5]; bool init = false; double acc[4] = {0, 0, 0, 0}; long frames = 0;
    void mark(int i, cudaStream_t s) { if (!init) { for (auto &e : ev) cudaEventCreate(&e); init = true; } cudaEventRecord(ev[i], s); }
    void finish() {
        cudaEventSynchronize(ev[4]);
        for (int i = 0; i < 4; i++) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); acc[i] += ms * 1e3; }
        if (++frames % 100 == 0)
            fprintf(stderr, "bq2 trace (GPU, us/frame): H2D %.1f  clear + record-building kernel %.1f  frame kernel(s) %.1f  counts to host %.1f\n",
                    acc[0] / frames, acc[1] / frames, acc[2] / frames, acc[3] / frames);
    }
};
StageTimer g_stage_timer;

} // namespace

extern "C" {

static void worker_drain(bq2_ctx *ctx);
static bq2_status worker_join_frame(bq2_ctx *ctx, FrameState &F);
static bq2_status worker_post(bq2_ctx *ctx, int k);
static void worker_wake(bq2_ctx *ctx);
static bq2_status world_enqueue(bq2_ctx *ctx, FrameState &F);
static void stage_caller_arrays(FrameState &F);
static void finish_bases(FrameState &F);
static bool host_pinned(const bq2_ctx *ctx, const void *p, size_t bytes);

int bq2_abi_version(void) { return BQ2_ABI_VERSION; }
int bq2_max_frames_in_flight(void) { return BQ2_MAX_IN_FLIGHT; }

bq2_status bq2_create(int device, bq2_ctx **out)
{
    if (!out) return fail(nullptr, BQ2_E_INVALID, "bq2_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        g_create_error = std::string("bq2_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path";
        return BQ2_E_NODEVICE;
    }
    if (device < 0 || device >= n) return fail(nullptr, BQ2_E_INVALID, "bq2_create: device index out of range");
    bq2_ctx *ctx = new (std::nothrow) bq2_ctx();
    if (!ctx) return fail(nullptr, BQ2_E_NOMEM, "bq2_create: out of host memory");
    ctx->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess) { fail_cuda(nullptr, e, "cudaSetDevice"); delete ctx; return BQ2_E_CUDA; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { fail_cuda(nullptr, e, "cudaGetDeviceProperties"); delete ctx; return BQ2_E_CUDA; }
    if (prop.major != 10) {
        g_create_error = "bq2_create: kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor);
        delete ctx; return BQ2_E_NODEVICE;
    }
    for (int k = 0; k < BQ2_FRAME_STATES; k++) {
        if ((e = cudaStreamCreateWithFlags(&ctx->fs[k].stream, cudaStreamNonBlocking)) != cudaSuccess) {
            fail_cuda(nullptr, e, "cudaStreamCreate");
            for (FrameState &f : ctx->fs) f.release();
            delete ctx; return BQ2_E_CUDA;
        }
    }
    ctx->stream = ctx->fs[0].stream;
    if ((e = (cudaError_t)bq2_kernel_prepare()) != cudaSuccess) {
        fail_cuda(nullptr, e, "cudaFuncSetAttribute(max dynamic smem)");
        for (FrameState &f : ctx->fs) f.release();
            delete ctx; return BQ2_E_CUDA;
    }
    *out = ctx;
    return BQ2_OK;
}

void bq2_destroy(bq2_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->worker.joinable()) {
        worker_drain(ctx);
        ctx->wstop.store(true, std::memory_order_seq_cst);
        { std::lock_guard<std::mutex> lk(ctx->wmx); ctx->wcv.notify_all(); }
        ctx->worker.join();
    }
    cudaSetDevice(ctx->device);
    for (FrameState &f : ctx->fs) if (f.stream) cudaStreamSynchronize(f.stream);
    if (ctx->trace.on && ctx->trace.n > 64) {
        static const char *nm[8] = {"validate+slots/class", "class split", "entity records", "pair offsets/memcpy", "pair records",
                                    "buffers", "h2d+launch+d2h calls", ""};
        for (int i = 0; i < 7; i++) fprintf(stderr, "bq2 trace: %-22s %8.2f us/frame\n", nm[i], ctx->trace.acc[i] / (ctx->trace.n - 64));
    }
    for (Slab &s : ctx->slabs) cudaFree(s.base);
    ctx->d_meshes.release(); ctx->d_models.release();
    ctx->d_brushes.release(); ctx->d_brush_in.release(); ctx->brush_vcache.release();
    ctx->brush_icache.release(); ctx->brush_counts.release(); ctx->h_brush.release(); ctx->h_brush_lit.release();
    for (FrameState &f : ctx->fs) f.release();                   /* streams included */
    for (const bq2_ctx::HostRange &r : ctx->host_ranges) cudaFreeHost((void *)r.base);
    delete ctx;
}

const char *bq2_last_error(const bq2_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
void *bq2_stream(bq2_ctx *ctx) { return ctx ? (void *)ctx->fs[ctx->cur].stream : nullptr; }   /* of the last submitted frame */
int bq2_device(const bq2_ctx *ctx) { return ctx ? ctx->device : -1; }
uint64_t bq2_kernel_launches(const bq2_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

static bq2_status gpu_neighbours(bq2_ctx *ctx, const uint16_t *idx, int stride, int32_t T, int md3, int32_t *out);

bq2_status bq2_upload_md2(bq2_ctx *ctx, const void *blob, size_t blob_bytes, const int32_t *neighbours,
                          const uint8_t *tangents, const uint8_t *binormals, const uint8_t *normals,
                          uint32_t model_flags, int32_t *model_id)
try {
    if (!ctx) return BQ2_E_INVALID;
    worker_drain(ctx);
    if (!blob || !model_id || blob_bytes < sizeof(md2_header)) return fail(ctx, BQ2_E_INVALID, "bq2_upload_md2: bad arguments");
    const md2_header *h = (const md2_header *)blob;
    const int V = h->num_xyz, T = h->num_tris, F = h->num_frames;
    {   /* the reference's load-time checks, M:51247-51264 */
        bq2_status hs; const char *why = md2_header_problem(h, blob_bytes, &hs);
        if (why) { static thread_local std::string msg; msg = std::string("bq2_upload_md2: ") + why; return fail(ctx, hs, msg.c_str()); }
    }
    const unsigned char *base = (const unsigned char *)blob;
    const int16_t *tri = (const int16_t *)(base + h->ofs_tris);
    std::vector<int32_t> idx(3 * (size_t)T);
    for (int t = 0; t < T; t++)
        for (int k = 0; k < 3; k++) {
            int v = tri[6 * t + k];
            if (v < 0 || v >= V) return fail(ctx, BQ2_E_INVALID, "bq2_upload_md2: triangle index out of range");
            idx[3 * t + k] = v;
        }
    std::vector<int32_t> built;
    if (!neighbours && (model_flags & BQ2_MODEL_BUILD_NEIGHBOURS)) {         /* findneighbour rules, on the device */
        built.resize(3 * (size_t)T);
        bq2_status ns = gpu_neighbours(ctx, (const uint16_t *)tri, 6, T, 0, built.data());
        if (ns != BQ2_OK) return ns;
        neighbours = built.data();
    }
    if (neighbours)
        for (size_t i = 0; i < 3 * (size_t)T; i++)
            if (neighbours[i] < -1 || neighbours[i] >= T) return fail(ctx, BQ2_E_INVALID, "bq2_upload_md2: neighbour out of range");

    const bool smooth = model_flags & BQ2_MODEL_SMOOTHVECS;
    const bool has_tb = tangents && binormals && (smooth || normals);
    const int Tp = (T + 31) & ~31;
    const uint32_t vstride = up16(4u * (uint32_t)V);
    const uint32_t rows = smooth ? (uint32_t)V : (uint32_t)T, tbstride = up16(rows);
    const size_t off_tri = (size_t)F * vstride;
    const size_t off_tb = off_tri + up16(12u * (uint32_t)Tp);
    const size_t n_tb = has_tb ? (smooth ? 2u : 3u) : 0u;
    const size_t total = off_tb + n_tb * (size_t)F * tbstride;
    std::vector<unsigned char> img(total, 0);
    for (int f = 0; f < F; f++)
        memcpy(&img[(size_t)f * vstride], base + h->ofs_frames + (size_t)f * h->framesize + 40, 4u * (size_t)V);
    pack_tri_table(T, Tp, idx.data(), neighbours, (int16_t *)&img[off_tri]);
    if (has_tb)
        for (int f = 0; f < F; f++) {
            memcpy(&img[off_tb + (size_t)f * tbstride], tangents + (size_t)f * rows, rows);
            memcpy(&img[off_tb + ((size_t)F + f) * tbstride], binormals + (size_t)f * rows, rows);
            if (!smooth) memcpy(&img[off_tb + (2 * (size_t)F + f) * tbstride], normals + (size_t)f * rows, rows);
        }
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    unsigned char *d = nullptr;
    bq2_status s = arena_alloc(ctx, total, &d);
    if (s != BQ2_OK) return s;
    CK(cudaMemcpyAsync(d, img.data(), total, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpy(model)");
    CK(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize(model)");

    bq2_dev_mesh m{};
    m.verts = d; m.tri = (const int16_t *)(d + off_tri);
    m.tb = has_tb ? d + off_tb : nullptr;
    m.bn = has_tb ? d + off_tb + (size_t)F * tbstride : nullptr;
    m.nm = (has_tb && !smooth) ? d + off_tb + 2 * (size_t)F * tbstride : nullptr;
    m.V = V; m.T = T; m.F = F; m.Tp = Tp; m.vstride = vstride; m.tbstride = tbstride;
    m.flags = (smooth ? BQ2_MESH_SMOOTH : 0u) | (has_tb ? BQ2_MESH_HAS_TB : 0u);
    Model mo; mo.md3 = false; mo.F = F; mo.first_mesh = (int)ctx->meshes.size(); mo.n_mesh = 1; mo.flags = model_flags;
    mo.V.push_back(V); mo.T.push_back(T); mo.casts.push_back(1);
    mo.hdr.resize(6 * (size_t)F);
    for (int f = 0; f < F; f++) memcpy(&mo.hdr[6 * (size_t)f], base + h->ofs_frames + (size_t)f * h->framesize, 24);
    if ((s = upload_hdr(ctx, mo)) != BQ2_OK) return s;
    ctx->meshes.push_back(m); ctx->models.push_back(mo); ctx->meshes_dirty = true;
    *model_id = (int32_t)ctx->models.size() - 1;
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

bq2_status bq2_upload_md3(bq2_ctx *ctx, int32_t F, const float *frame_translate, const bq2_md3_mesh *meshes,
                          int32_t num_meshes, int32_t *model_id)
try {
    if (!ctx) return BQ2_E_INVALID;
    worker_drain(ctx);
    if (!meshes || !model_id || !frame_translate || F <= 0 || num_meshes <= 0) return fail(ctx, BQ2_E_INVALID, "bq2_upload_md3: bad arguments");
    if (num_meshes > BQ2_MD3_MAX_MESHES) return fail(ctx, BQ2_E_LIMIT, "bq2_upload_md3: too many meshes");
    if (F > 1024) return fail(ctx, BQ2_E_LIMIT, "bq2_upload_md3: too many frames");
    for (int m = 0; m < num_meshes; m++) {
        const bq2_md3_mesh &me = meshes[m];
        if (me.num_verts <= 0 || me.num_tris <= 0 || !me.vertexes || !me.indexes)
            return fail(ctx, BQ2_E_INVALID, "bq2_upload_md3: empty mesh");
        if (me.num_verts > BQ2_MD3_MAX_VERTS) return fail(ctx, BQ2_E_LIMIT, "bq2_upload_md3: mesh has too many vertices");
        if (me.num_tris > BQ2_MD3_MAX_TRIANGLES) return fail(ctx, BQ2_E_LIMIT, "bq2_upload_md3: mesh has too many triangles");
        for (int i = 0; i < 3 * me.num_tris; i++) {
            if (me.indexes[i] >= me.num_verts) return fail(ctx, BQ2_E_INVALID, "bq2_upload_md3: index out of range");
            if (me.neighbours && (me.neighbours[i] < -1 || me.neighbours[i] >= me.num_tris))
                return fail(ctx, BQ2_E_INVALID, "bq2_upload_md3: neighbour out of range");
        }
    }
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    Model mo; mo.md3 = true; mo.F = F; mo.first_mesh = (int)ctx->meshes.size(); mo.n_mesh = num_meshes; mo.flags = BQ2_MODEL_SMOOTHVECS;
    mo.hdr.assign(frame_translate, frame_translate + 3 * (size_t)F);   /* host-only: folded into bq2_entity.move */
    for (int mi = 0; mi < num_meshes; mi++) {
        const bq2_md3_mesh &me = meshes[mi];
        const int V = me.num_verts, T = me.num_tris, Tp = (T + 31) & ~31;
        const uint32_t vstride = 16u * (uint32_t)V;
        const size_t off_tri = (size_t)F * vstride, total = off_tri + up16(12u * (uint32_t)Tp);
        std::vector<int32_t> idx(3 * (size_t)T);
        for (size_t i = 0; i < idx.size(); i++) idx[i] = me.indexes[i];
        std::vector<int16_t> tab(6 * (size_t)Tp);
        std::vector<int32_t> built;
        const int32_t *nbp = me.neighbours;
        if (!nbp && (me.casts_shadow & 2)) {                                 /* R_BuildTriangleNeighbors rules, on the device */
            built.resize(3 * (size_t)T);
            bq2_status ns = gpu_neighbours(ctx, me.indexes, 3, T, 1, built.data());
            if (ns != BQ2_OK) return ns;
            nbp = built.data();
        }
        pack_tri_table(T, Tp, idx.data(), nbp, tab.data());
        unsigned char *d = nullptr;
        bq2_status s = arena_alloc(ctx, total, &d);
        if (s != BQ2_OK) return s;
        CK(cudaMemcpyAsync(d, me.vertexes, off_tri, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpy(md3 verts)");
        CK(cudaMemcpyAsync(d + off_tri, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpy(md3 tris)");
        CK(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize(md3)");
        bq2_dev_mesh m{};
        m.verts = d; m.tri = (const int16_t *)(d + off_tri);
        m.V = V; m.T = T; m.F = F; m.Tp = Tp; m.vstride = vstride; m.tbstride = 0;
        m.flags = BQ2_MESH_MD3 | BQ2_MESH_SMOOTH | BQ2_MESH_HAS_TB;
        ctx->meshes.push_back(m);
        mo.V.push_back(V); mo.T.push_back(T); mo.casts.push_back((me.casts_shadow & 1) ? 1 : 0);
    }
    { bq2_status hs = upload_hdr(ctx, mo); if (hs != BQ2_OK) return hs; }
    ctx->models.push_back(mo); ctx->meshes_dirty = true;
    *model_id = (int32_t)ctx->models.size() - 1;
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

bq2_status bq2_model_info(const bq2_ctx *ctx, int32_t id, int32_t *is_md3, int32_t *num_frames,
                          int32_t *num_meshes, int32_t *num_verts, int32_t *num_tris)
{
    if (!ctx || id < 0 || id >= (int32_t)ctx->models.size()) return BQ2_E_INVALID;
    const Model &m = ctx->models[id];
    if (is_md3) *is_md3 = m.md3; if (num_frames) *num_frames = m.F; if (num_meshes) *num_meshes = m.n_mesh;
    for (int i = 0; i < m.n_mesh; i++) { if (num_verts) num_verts[i] = m.V[i]; if (num_tris) num_tris[i] = m.T[i]; }
    return BQ2_OK;
}

static bq2_status build_records_impl(bq2_ctx *ctx, const bq2_world_entity *ents, int32_t n_ents,
                             const bq2_world_light *lights, int32_t n_lights,
                             const int32_t *pair_ent, const int32_t *pair_light, int32_t n_pairs,
                             const float *view_world, const bq2_render_opts *opts, bq2_entity *out_ents, bq2_pair *out_pairs,
                             int32_t *n_pairs_out, bool keep_dead);

bq2_status bq2_model_set_rf_flags(bq2_ctx *ctx, int32_t id, uint32_t model_rf_flags)
{
    if (!ctx || id < 0 || id >= (int32_t)ctx->models.size()) return ctx ? fail(ctx, BQ2_E_INVALID, "bq2_model_set_rf_flags: unknown model") : BQ2_E_INVALID;
    ctx->models[id].rf_flags = model_rf_flags;
    if ((size_t)id < ctx->model_rows.size()) ctx->model_rows[id].pad[0] = (int32_t)model_rf_flags;
    return BQ2_OK;
}

bq2_status bq2_build_records_opts(bq2_ctx *ctx, const bq2_world_entity *ents, int32_t n_ents,
                                  const bq2_world_light *lights, int32_t n_lights,
                                  const int32_t *pair_ent, const int32_t *pair_light, int32_t n_pairs,
                                  const float *view_world, const bq2_render_opts *opts,
                                  bq2_entity *out_ents, bq2_pair *out_pairs, int32_t *n_pairs_out)
try {
    return build_records_impl(ctx, ents, n_ents, lights, n_lights, pair_ent, pair_light, n_pairs, view_world, opts,
                              out_ents, out_pairs, n_pairs_out, false);
} catch (...) { return guard_exception(ctx); }

bq2_status bq2_build_records(bq2_ctx *ctx, const bq2_world_entity *ents, int32_t n_ents,
                             const bq2_world_light *lights, int32_t n_lights,
                             const int32_t *pair_ent, const int32_t *pair_light, int32_t n_pairs,
                             const float *view_world, bq2_entity *out_ents, bq2_pair *out_pairs,
                             int32_t *n_pairs_out)
try {
    return build_records_impl(ctx, ents, n_ents, lights, n_lights, pair_ent, pair_light, n_pairs, view_world, nullptr,
                              out_ents, out_pairs, n_pairs_out, false);
} catch (...) { return guard_exception(ctx); }

static bq2_status build_records_impl(bq2_ctx *ctx, const bq2_world_entity *ents, int32_t n_ents,
                             const bq2_world_light *lights, int32_t n_lights,
                             const int32_t *pair_ent, const int32_t *pair_light, int32_t n_pairs,
                             const float *view_world, const bq2_render_opts *opts, bq2_entity *out_ents, bq2_pair *out_pairs,
                             int32_t *n_pairs_out, bool keep_dead)
{
    if (!ctx) return BQ2_E_INVALID;
    int f_off, f_brush; uint32_t f_em, f_mm, f_mx;
    bq2_host_filter_masks(opts, &f_off, &f_em, &f_mm, &f_mx, &f_brush);
    if (n_ents < 0 || n_pairs < 0 || (n_ents && (!ents || !out_ents)) ||
        (n_pairs && (!lights || !pair_ent || !pair_light || !out_pairs)) || !n_pairs_out)
        return fail(ctx, BQ2_E_INVALID, "bq2_build_records: bad arguments");
    const int nModels = (int)ctx->models.size();
    /* per entity: lerp constants, the rotation basis (AngleVectors once, not once per light) and the flag filter */
    ctx->ent_basis.resize(10 * (size_t)n_ents);
    for (int e = 0; e < n_ents; e++) {
        const bq2_world_entity &W = ents[e];
        if (W.model < 0 || W.model >= nModels) return fail(ctx, BQ2_E_INVALID, "bq2_build_records: unknown model");
        const Model &mo = ctx->models[W.model];
        if (W.frame < 0 || W.frame >= mo.F || W.oldframe < 0 || W.oldframe >= mo.F)
            return fail(ctx, BQ2_E_INVALID, "bq2_build_records: frame out of range");
        if (mo.md3)
            bq2_host_entity_md3(&mo.hdr[3 * (size_t)W.frame], &mo.hdr[3 * (size_t)W.oldframe], W.frame, W.oldframe,
                                W.backlerp, W.origin, W.oldorigin, W.angles, &out_ents[e]);
        else
            bq2_host_entity_md2_hdr(&mo.hdr[6 * (size_t)W.frame], &mo.hdr[6 * (size_t)W.oldframe], W.frame, W.oldframe,
                                    W.backlerp, W.origin, W.oldorigin, W.angles, W.rf_flags, &out_ents[e]);
        out_ents[e].model = W.model;
        float *b = &ctx->ent_basis[10 * (size_t)e];
        const bool casts = !(f_off || (W.rf_flags & f_em) || ((mo.rf_flags ^ f_mx) & f_mm));   /* M:63722-63727, 64041-64089 */
        b[9] = !casts ? -1.0f : (float)bq2_host_entity_basis(W.angles, b);
    }
    int kept = 0;
    float view_local[3] = {0.0f, 0.0f, 0.0f};
    for (int k = 0; k < n_pairs; k++) {
        const int e = pair_ent[k], l = pair_light[k];
        if ((unsigned)e >= (unsigned)n_ents || (unsigned)l >= (unsigned)n_lights)
            return fail(ctx, BQ2_E_INVALID, "bq2_build_records: pair out of range");
        const float *b = &ctx->ent_basis[10 * (size_t)e];
        if (b[9] < 0.0f) {                                                          /* M:63722-63727 */
            if (keep_dead) { bq2_pair &D = out_pairs[kept++]; memset(&D, 0, sizeof D); D.entity = ~e; }
            continue;
        }
        const bq2_world_entity &W = ents[e];
        bq2_pair &P = out_pairs[kept++];
        P.entity = e;
        bq2_host_point_to_model_basis(lights[l].origin, W.origin, b[9] > 0.0f ? b : nullptr, P.light);
        P.scale = 32 * lights[l].radius;                                            /* M:63597 */
        if (view_world) bq2_host_point_to_model_basis(view_world, W.origin, b[9] > 0.0f ? b : nullptr, P.view);
        else { P.view[0] = view_local[0]; P.view[1] = view_local[1]; P.view[2] = view_local[2]; }
    }
    *n_pairs_out = kept;
    return BQ2_OK;
}

bq2_status bq2_frame_submit(bq2_ctx *ctx, const bq2_frame_desc *fd)
try {
    if (!ctx) return BQ2_E_INVALID;
    worker_drain(ctx);
    if (!fd || fd->n_ents < 0 || fd->n_pairs < 0 || (fd->n_ents > 0 && !fd->ents) || (fd->n_pairs > 0 && !fd->pairs))
        return fail(ctx, BQ2_E_INVALID, "bq2_frame: bad descriptor");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    if (ctx->submit_seq - ctx->wait_seq >= BQ2_MAX_IN_FLIGHT) return fail(ctx, BQ2_E_INVALID, "bq2_frame_submit: the maximum number of frames is already in flight (bq2_max_frames_in_flight), call bq2_frame_wait first");
    const int slot_k = pick_state(ctx);
    if (slot_k < 0) return fail(ctx, BQ2_E_INVALID, "bq2_frame_submit: no free frame state");
    FrameState &F = ctx->fs[slot_k];
    PinBuf &h_records = F.hrec, &h_counts = F.hcnt;
    const int nE = fd->n_ents, nP = fd->n_pairs;
    const int nModels = (int)ctx->models.size();
    if (fd->want & ~BQ2_WANT_PUBLIC_MASK) return fail(ctx, BQ2_E_INVALID, "bq2_frame: unknown bits in want");
    const uint32_t want = fd->want & ~BQ2_WANT_TABLES;
    Trace &tr = ctx->trace; tr.start(); tr.n++;

    /* ---- entity slots ------------------------------------------------------------------------- */
    F.ent_slot_first.resize((size_t)nE + 1);
    int nES = 0, maxV = 0, maxT = 0, anyMd2 = 0;
    for (int e = 0; e < nE; e++) {
        int mid = fd->ents[e].model;
        if (mid < 0 || mid >= nModels) return fail(ctx, BQ2_E_INVALID, "bq2_frame: entity refers to an unknown model");
        const Model &mo = ctx->models[mid];
        if (fd->ents[e].frame < 0 || fd->ents[e].frame >= mo.F || fd->ents[e].oldframe < 0 || fd->ents[e].oldframe >= mo.F)
            return fail(ctx, BQ2_E_INVALID, "bq2_frame: frame out of range");
        F.ent_slot_first[e] = nES; nES += mo.n_mesh;
    }
    F.ent_slot_first[nE] = nES;
    F.ent_vert_off.resize((size_t)nES + 1); F.ent_tb_off.resize((size_t)nES + 1);
    F.cursor.assign((size_t)nES + 1, 0u);                    /* pair-slot count per entity slot */

    /* ---- pair slots ---------------------------------------------------------------------------- */
    F.pair_slot_first.resize((size_t)nP + 1);
    int nPS = 0;
    for (int p = 0; p < nP; p++) {
        int e = fd->pairs[p].entity;
        if (e < 0) e = ~e;                                       /* a pair that keeps its slot but casts nothing */
        if (e < 0 || e >= nE) return fail(ctx, BQ2_E_INVALID, "bq2_frame: pair refers to an unknown entity");
        F.pair_slot_first[p] = nPS; nPS += ctx->models[fd->ents[e].model].n_mesh;
    }
    F.pair_slot_first[nP] = nPS;
    F.pair_vert_off.resize((size_t)nPS + 1); F.pair_bits_off.resize((size_t)nPS + 1); F.pair_ts_off.resize((size_t)nPS + 1);
    F.slot_mesh_V.resize((size_t)nPS); F.slot_mesh_T.resize((size_t)nPS); F.pair_slot_entslot.resize((size_t)nPS);

    size_t rec_bytes = (size_t)nES * sizeof(bq2_dev_ent) + (size_t)nPS * sizeof(bq2_dev_pair);
    CK(h_records.ensure(std::max<size_t>(rec_bytes, 16)), "cudaMallocHost(records)");
    CK(F.d_records.ensure(std::max<size_t>(rec_bytes, 16)), "cudaMalloc(records)");
    bq2_dev_ent  *he = (bq2_dev_ent *)h_records.p;
    bq2_dev_pair *hp = (bq2_dev_pair *)(h_records.p + (size_t)nES * sizeof(bq2_dev_ent));

    tr.lap(0);
    /* which kernel takes each entity slot: small meshes -> one warp per pair; large meshes and the
       tangent-space outputs -> the whole CTA per pair */
    const bool team_only = false;          /* both kernels write the tangent-space rows */
    F.slot_record.resize((size_t)nES);
    F.k_warp = FrameState::KLaunch(); F.k_team = FrameState::KLaunch();
    for (int e = 0; e < nE; e++) {
        const Model &mo = ctx->models[fd->ents[e].model];
        for (int m = 0; m < mo.n_mesh; m++) {
            const bq2_dev_mesh &me = ctx->meshes[mo.first_mesh + m];
            bool small = !team_only && me.V <= BQ2_WARP_MAX_V && me.T <= BQ2_WARP_MAX_T;
            FrameState::KLaunch &k = small ? F.k_warp : F.k_team;
            F.slot_record[(size_t)F.ent_slot_first[e] + m] = small ? k.n : -1 - k.n;
            k.n++; k.max_V = std::max(k.max_V, me.V); k.max_T = std::max(k.max_T, me.T);
            if (!(me.flags & BQ2_MESH_MD3)) k.any_md2 = 1;
        }
    }
    for (int s = 0; s < nES; s++)
        if (F.slot_record[s] < 0) F.slot_record[s] = F.k_warp.n + (-1 - F.slot_record[s]);

    tr.lap(1);
    /* rows are numbered in 32 bits on the device: the running sums are kept in 64 and checked after the loops */
    uint64_t voff = 0, tboff = 0, team_rows = 0;
    uint64_t total_verts = 0;
    for (int e = 0; e < nE; e++) {
        const bq2_entity &E = fd->ents[e];
        const Model &mo = ctx->models[E.model];
        for (int m = 0; m < mo.n_mesh; m++) {
            int s = F.ent_slot_first[e] + m;
            const bq2_dev_mesh &me = ctx->meshes[mo.first_mesh + m];
            if ((want & BQ2_WANT_NTB) && !(me.flags & BQ2_MESH_HAS_TB))
                return fail(ctx, BQ2_E_INVALID, "bq2_frame: BQ2_WANT_NTB but the model was uploaded without tangent/binormal bytes");
            bq2_dev_ent &d = he[F.slot_record[s]];
            d.cur = me.verts + (size_t)E.frame * me.vstride; d.old = me.verts + (size_t)E.oldframe * me.vstride;
            d.tri = me.tri; d.V = me.V; d.T = me.T; d.Tp = me.Tp; d.mflags = me.flags; d.vstride = me.vstride; d.nrm_off = 0;
            if (F.slot_record[s] >= F.k_warp.n) { d.nrm_off = (uint32_t)team_rows; team_rows += (uint32_t)me.Tp; }
            d.mesh = mo.first_mesh + m; d.frame = E.frame; d.oldframe = E.oldframe; d.flags = E.flags;
            d.backlerp = E.backlerp; d.frontlerp = E.frontlerp;
            for (int k = 0; k < 3; k++) { d.move[k] = E.move[k]; d.frontv[k] = E.frontv[k]; d.backv[k] = E.backv[k]; }
            d.shell_scale = E.shell_scale;
            d.vert_off = (uint32_t)voff; d.tb_off = (uint32_t)tboff; d.pair_begin = 0; d.pair_count = 0;
            F.ent_vert_off[s] = (uint32_t)voff; F.ent_tb_off[s] = (uint32_t)tboff;
            voff += (uint32_t)(me.V + 3) & ~3u;
            tboff += (uint32_t)(((me.flags & BQ2_MESH_SMOOTH) ? me.V : me.T) + 3) & ~3u;
            total_verts += (uint64_t)me.V;
            maxV = std::max(maxV, me.V); maxT = std::max(maxT, me.T);
            if (!(me.flags & BQ2_MESH_MD3)) anyMd2 = 1;
        }
    }
    if ((voff | tboff | team_rows) >> 32) return fail(ctx, BQ2_E_LIMIT, "bq2_frame: more than 2^32 output rows, split the frame");
    F.ent_vert_off[nES] = (uint32_t)voff; F.ent_tb_off[nES] = (uint32_t)tboff;
    const uint32_t ent_verts = (uint32_t)voff, tb_rows = (uint32_t)tboff;

    tr.lap(2);
    /* pair slots: offsets in caller order, then a counting sort by entity slot for the kernel */
    uint64_t pvoff = 0, pboff = 0, ptsoff = 0;
    for (int p = 0; p < nP; p++) {
        int e = fd->pairs[p].entity;
        const bool dead = e < 0;
        if (dead) e = ~e;
        const Model &mo = ctx->models[fd->ents[e].model];
        for (int m = 0; m < mo.n_mesh; m++) {
            int ps = F.pair_slot_first[p] + m, es = F.ent_slot_first[e] + m;
            const bq2_dev_mesh &me = ctx->meshes[mo.first_mesh + m];
            F.pair_vert_off[ps] = (uint32_t)pvoff; F.pair_bits_off[ps] = (uint32_t)pboff; F.pair_ts_off[ps] = (uint32_t)ptsoff;
            /* LightP / tsH rows: per vertex, or per triangle corner for a non-smooth MD2 (M:64943-65017) */
            ptsoff += (uint32_t)(((me.flags & BQ2_MESH_SMOOTH) ? me.V : 3 * me.T) + 3) & ~3u;
            F.slot_mesh_V[ps] = me.V; F.slot_mesh_T[ps] = me.T; F.pair_slot_entslot[ps] = es;
            pvoff += (uint32_t)(me.V + 3) & ~3u;
            pboff += (uint32_t)(me.T + 31) >> 5;
            if (mo.casts[m] && !dead) F.cursor[es]++;
        }
    }
    if ((pvoff | pboff | ptsoff) >> 32) return fail(ctx, BQ2_E_LIMIT, "bq2_frame: more than 2^32 output rows, split the frame");
    F.pair_vert_off[nPS] = (uint32_t)pvoff; F.pair_bits_off[nPS] = (uint32_t)pboff; F.pair_ts_off[nPS] = (uint32_t)ptsoff;
    uint32_t n_dev_pairs = 0;
    {
        uint32_t run = 0;
        for (int s = 0; s < nES; s++) {
            uint32_t c = F.cursor[s];
            bq2_dev_ent &d = he[F.slot_record[s]];
            d.pair_begin = run; d.pair_count = c; F.cursor[s] = run; run += c;
        }
        n_dev_pairs = run;
    }
    tr.lap(3);
    for (int p = 0; p < nP; p++) {
        const bq2_pair &P = fd->pairs[p];
        if (P.entity < 0) continue;
        const Model &mo = ctx->models[fd->ents[P.entity].model];
        for (int m = 0; m < mo.n_mesh; m++) {
            if (!mo.casts[m]) continue;
            int ps = F.pair_slot_first[p] + m, es = F.ent_slot_first[P.entity] + m;
            bq2_dev_pair &d = hp[F.cursor[es]++];
            for (int k = 0; k < 3; k++) { d.light[k] = P.light[k]; d.view[k] = P.view[k]; }
            d.scale = P.scale; d.vert_off = F.pair_vert_off[ps]; d.bits_off = F.pair_bits_off[ps];
            d.slot = (uint32_t)ps; d.pad[0] = F.pair_ts_off[ps]; d.pad[1] = 0;
        }
    }

    tr.lap(4);
    uint32_t stride = fd->index_stride ? fd->index_stride : 12u * (uint32_t)maxT;
    stride = (stride + 3u) & ~3u;

    /* ---- output capacity ------------------------------------------------------------------------ */
    CK(F.lerped.ensure(3 * (size_t)ent_verts + 4), "cudaMalloc(lerped)");
    if (want & BQ2_WANT_SHADOW) {
        CK(F.extruded.ensure(3 * (size_t)pvoff + 4), "cudaMalloc(extruded)");
        CK(F.facing_bits.ensure((size_t)pboff + 1), "cudaMalloc(facing_bits)");
        CK(F.indices.ensure((size_t)nPS * stride + 4), "cudaMalloc(indices)");
    }
    CK(F.index_count.ensure((size_t)nPS + 1), "cudaMalloc(index_count)");
    if (want & BQ2_WANT_NTB) {
        CK(F.normals.ensure(3 * (size_t)tb_rows + 4), "cudaMalloc(normals)");
        CK(F.tangents.ensure(3 * (size_t)tb_rows + 4), "cudaMalloc(tangents)");
        CK(F.binormals.ensure(3 * (size_t)tb_rows + 4), "cudaMalloc(binormals)");
    }
    if (want & BQ2_WANT_TSLIGHT) {
        if (!(want & BQ2_WANT_NTB)) return fail(ctx, BQ2_E_INVALID, "bq2_frame: BQ2_WANT_TSLIGHT needs BQ2_WANT_NTB");
        CK(F.lightp.ensure(3 * (size_t)ptsoff + 4), "cudaMalloc(lightp)");
        CK(F.tsh.ensure(3 * (size_t)ptsoff + 4), "cudaMalloc(tsh)");
    }
    if (team_rows) CK(F.tri_normals.ensure(4 * (size_t)team_rows + 4), "cudaMalloc(triangle normals)");
    CK(h_counts.ensure(sizeof(uint32_t) * ((size_t)nPS + 1)), "cudaMallocHost(counts)");
    if (ctx->meshes_dirty) {
        CK(ctx->d_meshes.ensure(ctx->meshes.size()), "cudaMalloc(mesh table)");
        CK(cudaMemcpyAsync(ctx->d_meshes.p, ctx->meshes.data(), ctx->meshes.size() * sizeof(bq2_dev_mesh),
                           cudaMemcpyHostToDevice, F.stream), "cudaMemcpy(mesh table)");
        CK(cudaStreamSynchronize(F.stream), "sync(mesh table)");   /* source is pageable */
        ctx->meshes_dirty = false;
    }
    if (F.k_team.n && bq2_kernel_smem_bytes(F.k_team.max_V, F.k_team.max_T, F.k_team.any_md2) > 227u * 1024u)
        return fail(ctx, BQ2_E_LIMIT, "bq2_frame: mesh too large for one CTA's shared memory");

    /* ---- go -------------------------------------------------------------------------------------- */
    tr.lap(5);
    F.tail_clean = false;                                        /* this frame's counts overwrite what a world frame left cleared */
    if (rec_bytes)
        CK(cudaMemcpyAsync(F.d_records.p, h_records.p, rec_bytes, cudaMemcpyHostToDevice, F.stream), "cudaMemcpy(records)");
    /* slots of non-casting MD3 meshes are never written by a kernel, and a frame without BQ2_WANT_SHADOW writes no
       count at all: only then is the clear needed */
    if (nPS && (n_dev_pairs != (uint32_t)nPS || !(want & BQ2_WANT_SHADOW)))
        CK(cudaMemsetAsync(F.index_count.p, 0, sizeof(uint32_t) * (size_t)nPS, F.stream), "cudaMemset(index_count)");

    bq2_launch &L = F.launch;
    L.meshes = ctx->d_meshes.p;
    L.ents = (const bq2_dev_ent *)F.d_records.p;
    L.pairs = (const bq2_dev_pair *)(F.d_records.p + (size_t)nES * sizeof(bq2_dev_ent));
    L.n_ent_slots = nES; L.want = want; L.index_stride = stride;
    L.lerped = F.lerped.p; L.extruded = F.extruded.p; L.facing_bits = F.facing_bits.p;
    L.indices = F.indices.p; L.index_count = F.index_count.p;
    L.normals = F.normals.p; L.tangents = F.tangents.p; L.binormals = F.binormals.p;
    L.lightp = F.lightp.p; L.tsh = F.tsh.p; L.totals = nullptr; L.tri_normals = F.tri_normals.p;
    L.bucket_cnt = nullptr; L.bucket = nullptr; L.ohead = nullptr; L.onext = nullptr; L.rec_base = 0;
    F.world_mode = false;
    F.max_V = maxV; F.max_T = maxT; F.any_md2 = anyMd2;
    F.n_ent_slots = nES; F.n_pair_slots = nPS; F.n_ents = nE; F.n_pairs = nP;
    F.total_verts = total_verts;
    { bq2_status ls = launch_kernels(ctx, F); if (ls != BQ2_OK) return ls; }
    if (nPS) CK(cudaMemcpyAsync(h_counts.p, F.index_count.p, sizeof(uint32_t) * (size_t)nPS,
                                cudaMemcpyDeviceToHost, F.stream), "cudaMemcpy(counts)");
    ctx->submitted = true;
    { FrameState::Meta &m = F.meta; m.world = false; m.tables = false; m.n_ent_slots = nES; m.n_pair_slots = nPS;
      m.total_verts = total_verts; m.stride = stride; }
    ctx->order[ctx->submit_seq % BQ2_FRAME_STATES] = slot_k; ctx->busy[slot_k] = true;
    ctx->submit_seq++; ctx->cur = slot_k;
    tr.lap(6);
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

bq2_status bq2_frame_wait(bq2_ctx *ctx, bq2_frame_out *out)
try {
    if (!ctx) return BQ2_E_INVALID;
    if (!ctx->submitted) return fail(ctx, BQ2_E_INVALID, "bq2_frame_wait: nothing submitted");
    int k;
    if (ctx->submit_seq == ctx->wait_seq) {                      /* nothing pending (e.g. after bq2_frame_replay) */
        worker_drain(ctx);
        CK(cudaStreamSynchronize(ctx->fs[ctx->cur].stream), "cudaStreamSynchronize(frame)");
        k = ctx->last_waited;
    } else {                                                     /* the OLDEST frame in flight */
        k = ctx->order[ctx->wait_seq % BQ2_FRAME_STATES];
        ctx->busy[k] = false;
        ctx->wait_seq++; ctx->last_waited = k;                   /* the frame is consumed whatever happens below */
        { bq2_status js = worker_join_frame(ctx, ctx->fs[k]); if (js != BQ2_OK) return js; }   /* the submit worker has issued it */
        CK(cudaStreamSynchronize(ctx->fs[k].stream), "cudaStreamSynchronize(frame)");   /* one frame per state and stream */
    }
    FrameState &F = ctx->fs[k];
    const FrameState::Meta &m = F.meta;
    bq2_frame_out &o = F.out;
    o.n_ent_slots = m.n_ent_slots; o.n_pair_slots = m.n_pair_slots;
    o.ent_slot_first = F.ent_slot_first.data(); o.pair_slot_first = F.pair_slot_first.data();
    o.ent_vert_offset = F.ent_vert_off.data(); o.pair_vert_offset = F.pair_vert_off.data();
    o.pair_bits_offset = F.pair_bits_off.data(); o.ent_tb_offset = F.ent_tb_off.data();
    o.pair_ts_offset = F.pair_ts_off.data();
    o.lerped = F.lerped.p; o.extruded = F.extruded.p; o.facing_bits = F.facing_bits.p;
    o.indices = F.indices.p; o.index_count = F.index_count.p;
    o.normals = F.normals.p; o.tangents = F.tangents.p; o.binormals = F.binormals.p;
    o.lightp = F.lightp.p; o.tsh = F.tsh.p;
    o.index_stride = m.stride;
    o.total_lerped_verts = m.total_verts;
    o.total_pairs = 0; o.total_indices = 0; o.overflow_pairs = 0;
    if (m.world) {                                               /* totals were accumulated by the kernels */
        const unsigned long long *t = (const unsigned long long *)((const uint32_t *)F.hcnt.p + (((size_t)m.n_pair_slots + 1) & ~(size_t)1));
        o.total_indices = t[0]; o.overflow_pairs = (int32_t)t[1];
        o.pair_slot_first = nullptr;                             /* slot p is caller pair p */
        o.pair_vert_offset = m.tables ? F.world_pvoff.data() : nullptr;
        o.pair_ts_offset = o.pair_vert_offset;                   /* LightP / tsH rows of a pair are numbered like its extruded row */
        o.pair_bits_offset = m.tables ? F.world_pboff.data() : nullptr;
        if (t[2]) { o.total_pairs = (uint64_t)m.n_pair_slots; if (out) *out = o; return fail(ctx, BQ2_E_INVALID, "bq2_world: a pair refers to an entity or light out of range (it was skipped)"); }
    } else {
        const uint32_t *cnt = (const uint32_t *)F.hcnt.p;
        for (int i = 0; i < m.n_pair_slots; i++) {
            o.total_indices += cnt[i];
            if (cnt[i] > o.index_stride) o.overflow_pairs++;
        }
    }
    o.total_pairs = (uint64_t)m.n_pair_slots;
    if (out) *out = o;
    if (o.overflow_pairs) return fail(ctx, BQ2_E_OVERFLOW, "bq2_frame: index_stride too small for at least one pair (counts are exact, indices truncated)");
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

bq2_status bq2_frame(bq2_ctx *ctx, const bq2_frame_desc *fd, bq2_frame_out *out)
{
    while (ctx && ctx->submit_seq != ctx->wait_seq) bq2_frame_wait(ctx, nullptr);     /* drain frames in flight */
    bq2_status s = bq2_frame_submit(ctx, fd);
    if (s != BQ2_OK) return s;
    return bq2_frame_wait(ctx, out);
}

static bq2_status gpu_neighbours(bq2_ctx *ctx, const uint16_t *idx, int stride, int32_t T, int md3, int32_t *out)
try {
    if (!ctx) return BQ2_E_INVALID;
    worker_drain(ctx);
    if (!idx || !out || T <= 0) return fail(ctx, BQ2_E_INVALID, "bq2_gpu_neighbours: bad arguments");
    if (T > BQ2_MD3_MAX_TRIANGLES) return fail(ctx, BQ2_E_LIMIT, "bq2_gpu_neighbours: too many triangles");
    int V = 0;
    for (int t = 0; t < T; t++) for (int k = 0; k < 3; k++) V = std::max(V, (int)idx[(size_t)t * stride + k] + 1);
    if (bq2_neighbours_smem_bytes(T, V) > 226u * 1024u) return fail(ctx, BQ2_E_LIMIT, "bq2_gpu_neighbours: mesh too large for one CTA");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t in_bytes = ((size_t)T * stride * 2 + 15) & ~(size_t)15, out_bytes = (size_t)T * 12;
    DevBuf<unsigned char> buf;
    CK(buf.ensure(in_bytes + out_bytes), "cudaMalloc(neighbours)");
    cudaError_t e = cudaMemcpyAsync(buf.p, idx, (size_t)T * stride * 2, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = (cudaError_t)bq2_launch_neighbours((const uint16_t *)buf.p, stride, T, V, md3, (int32_t *)(buf.p + in_bytes), ctx->stream);
    if (e == cudaSuccess) { ctx->launches++; e = cudaMemcpyAsync(out, buf.p + in_bytes, out_bytes, cudaMemcpyDeviceToHost, ctx->stream); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    buf.release();
    if (e != cudaSuccess) return fail_cuda(ctx, e, "bq2_gpu_neighbours");
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

bq2_status bq2_gpu_md2_neighbours(bq2_ctx *ctx, const int16_t *dtriangles, int32_t num_tris, int32_t *out)
{ return gpu_neighbours(ctx, (const uint16_t *)dtriangles, 6, num_tris, 0, out); }
bq2_status bq2_gpu_md3_neighbours(bq2_ctx *ctx, const uint16_t *indexes, int32_t num_tris, int32_t *out)
{ return gpu_neighbours(ctx, indexes, 3, num_tris, 1, out); }

bq2_status bq2_selftest_ratio(bq2_ctx *ctx, int mode, float num, uint64_t count, uint64_t seed, uint64_t *mismatches)
try {
    if (!ctx) return BQ2_E_INVALID;
    worker_drain(ctx);
    if (!mismatches || mode < 0 || mode > 2) return fail(ctx, BQ2_E_INVALID, "bq2_selftest_ratio: bad arguments");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    if (mode == 0) count = 1ull << 32;
    unsigned long long *d = nullptr, h = 0;
    CK(cudaMalloc(&d, sizeof h), "cudaMalloc(selftest)");
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof h, ctx->stream);
    if (e == cudaSuccess) e = (cudaError_t)bq2_launch_selftest_ratio(mode, num, count, seed, d, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "bq2_selftest_ratio");
    *mismatches = (uint64_t)h;
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

bq2_status bq2_gpu_md2_tangents(bq2_ctx *ctx, const void *blob, size_t blob_bytes, const float *st, float smooth_cosine,
                                int smooth, uint8_t *normals, uint8_t *tangents, uint8_t *binormals)
try {
    if (!ctx) return BQ2_E_INVALID;
    worker_drain(ctx);
    if (!blob || blob_bytes < sizeof(md2_header) || !st || !tangents || !binormals || (!smooth && !normals))
        return fail(ctx, BQ2_E_INVALID, "bq2_gpu_md2_tangents: bad arguments");
    const md2_header *h = (const md2_header *)blob;
    const int V = h->num_xyz, T = h->num_tris, F = h->num_frames, NS = h->num_st;
    {
        bq2_status hs; const char *why = md2_header_problem(h, blob_bytes, &hs);
        if (why) { static thread_local std::string msg; msg = std::string("bq2_gpu_md2_tangents: ") + why; return fail(ctx, hs, msg.c_str()); }
        if (NS <= 0) return fail(ctx, BQ2_E_INVALID, "bq2_gpu_md2_tangents: model has no texture coordinates");
    }
    const unsigned char *base = (const unsigned char *)blob;
    const int16_t *tri = (const int16_t *)(base + h->ofs_tris);
    for (int t = 0; t < T; t++) for (int k = 0; k < 3; k++)
        if (tri[6 * t + k] < 0 || tri[6 * t + k] >= V || tri[6 * t + 3 + k] < 0 || tri[6 * t + 3 + k] >= NS)
            return fail(ctx, BQ2_E_INVALID, "bq2_gpu_md2_tangents: triangle index out of range");
    const int mode = smooth ? 0 : 1;
    if (bq2_tangents_smem_bytes(T, V, NS, mode) > 226u * 1024u) return fail(ctx, BQ2_E_LIMIT, "bq2_gpu_md2_tangents: mesh too large for one CTA");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t rows = smooth ? (size_t)V : (size_t)T, vrow = (size_t)V * 4;
    const size_t o_tri = 0, o_st = (12 * (size_t)T + 15) & ~(size_t)15, o_v = o_st + (((size_t)NS * 8 + 15) & ~(size_t)15),
                 o_out = o_v + ((F * vrow + 15) & ~(size_t)15), total = o_out + 3 * F * rows;
    std::vector<unsigned char> img(o_out, 0);
    memcpy(&img[o_tri], tri, 12 * (size_t)T);
    memcpy(&img[o_st], st, (size_t)NS * 8);
    for (int f = 0; f < F; f++) memcpy(&img[o_v + f * vrow], base + h->ofs_frames + (size_t)f * h->framesize + 40, vrow);
    DevBuf<unsigned char> buf;
    CK(buf.ensure(total + 16), "cudaMalloc(tangents)");
    cudaError_t e = cudaMemcpyAsync(buf.p, img.data(), o_out, cudaMemcpyHostToDevice, ctx->stream);
    const uint16_t *dtri = (const uint16_t *)(buf.p + o_tri);
    uint8_t *dout = buf.p + o_out;
    if (e == cudaSuccess)
        e = (cudaError_t)bq2_launch_tangents(dtri, 6, dtri + 3, 6, (const float *)(buf.p + o_st), NS, buf.p + o_v, vrow, F, T, V, mode,
                                             smooth_cosine, dout, dout + F * rows, dout + 2 * F * rows, rows, ctx->stream);
    if (e == cudaSuccess) { ctx->launches++; e = cudaStreamSynchronize(ctx->stream); }
    if (e == cudaSuccess && !smooth) e = cudaMemcpy(normals, dout, F * rows, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(tangents, dout + F * rows, F * rows, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(binormals, dout + 2 * F * rows, F * rows, cudaMemcpyDeviceToHost);
    buf.release();
    if (e != cudaSuccess) return fail_cuda(ctx, e, "bq2_gpu_md2_tangents");
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

bq2_status bq2_gpu_md3_tangents(bq2_ctx *ctx, void *vertexes, int32_t F, int32_t V, const float *st,
                                const uint16_t *indexes, int32_t T)
try {
    if (!ctx) return BQ2_E_INVALID;
    worker_drain(ctx);
    if (!vertexes || !st || !indexes || F <= 0 || V <= 0 || T <= 0) return fail(ctx, BQ2_E_INVALID, "bq2_gpu_md3_tangents: bad arguments");
    if (V > BQ2_MD3_MAX_VERTS || T > BQ2_MD3_MAX_TRIANGLES) return fail(ctx, BQ2_E_LIMIT, "bq2_gpu_md3_tangents: mesh outside the MD3 limits");
    for (int i = 0; i < 3 * T; i++) if (indexes[i] >= V) return fail(ctx, BQ2_E_INVALID, "bq2_gpu_md3_tangents: index out of range");
    if (bq2_tangents_smem_bytes(T, V, V, 2) > 226u * 1024u) return fail(ctx, BQ2_E_LIMIT, "bq2_gpu_md3_tangents: mesh too large for one CTA");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t o_ix = 0, o_st = (6 * (size_t)T + 15) & ~(size_t)15, o_v = o_st + (((size_t)V * 8 + 15) & ~(size_t)15),
                 vbytes = (size_t)F * V * 16, total = o_v + vbytes;
    DevBuf<unsigned char> buf;
    CK(buf.ensure(total + 16), "cudaMalloc(md3 tangents)");
    cudaError_t e = cudaMemcpyAsync(buf.p + o_ix, indexes, 6 * (size_t)T, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(buf.p + o_st, st, (size_t)V * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(buf.p + o_v, vertexes, vbytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = (cudaError_t)bq2_launch_tangents((const uint16_t *)(buf.p + o_ix), 3, nullptr, 0, (const float *)(buf.p + o_st), V,
                                             buf.p + o_v, (size_t)V * 16, F, T, V, 2, 0.0f, nullptr, nullptr, nullptr, 0, ctx->stream);
    if (e == cudaSuccess) { ctx->launches++; e = cudaMemcpyAsync(vertexes, buf.p + o_v, vbytes, cudaMemcpyDeviceToHost, ctx->stream); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    buf.release();
    if (e != cudaSuccess) return fail_cuda(ctx, e, "bq2_gpu_md3_tangents");
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

bq2_status bq2_upload_brush(bq2_ctx *ctx, int32_t n_polys, const int32_t *numverts, const float *verts,
                            const int32_t *neighbours, const uint8_t *fence, int32_t *brush_id)
try {
    if (!ctx) return BQ2_E_INVALID;
    worker_drain(ctx);
    if (n_polys <= 0 || !numverts || !verts || !neighbours || !brush_id) return fail(ctx, BQ2_E_INVALID, "bq2_upload_brush: bad arguments");
    if (n_polys > BQ2_BRUSH_MAX_POLYS) return fail(ctx, BQ2_E_LIMIT, "bq2_upload_brush: too many polygons");
    std::vector<uint32_t> first((size_t)n_polys + 1);
    uint32_t ns = 0; int worst = 0;
    for (int p = 0; p < n_polys; p++) {
        if (numverts[p] < 3) return fail(ctx, BQ2_E_INVALID, "bq2_upload_brush: polygon with fewer than 3 vertices");
        if (numverts[p] > 255) return fail(ctx, BQ2_E_LIMIT, "bq2_upload_brush: polygon with more than 255 vertices");
        first[p] = ns; ns += (uint32_t)numverts[p]; worst += 6 * numverts[p] + 6 * (numverts[p] - 2);
    }
    first[n_polys] = ns;
    if (ns > BQ2_BRUSH_MAX_SLOTS) return fail(ctx, BQ2_E_LIMIT, "bq2_upload_brush: too many polygon vertices");
    /* per slot: owning polygon | fence << 31, position in the polygon | numverts << 16 */
    std::vector<uint32_t> slot_info(2 * (size_t)ns);
    for (int p = 0; p < n_polys; p++)
        for (uint32_t k = first[p]; k < first[p + 1]; k++) {
            slot_info[2 * (size_t)k] = (uint32_t)p | ((fence && fence[p]) ? 0x80000000u : 0u);
            slot_info[2 * (size_t)k + 1] = (k - first[p]) | ((uint32_t)numverts[p] << 16);
        }
    for (uint32_t k = 0; k < ns; k++) if (neighbours[k] < -1 || neighbours[k] >= n_polys) return fail(ctx, BQ2_E_INVALID, "bq2_upload_brush: neighbour out of range");
    const size_t o_v = 0, o_si = 16 * (size_t)ns, o_nbr = o_si + 8 * (size_t)ns, total = o_nbr + 4 * (size_t)ns;
    std::vector<unsigned char> img(total, 0);
    for (uint32_t k = 0; k < ns; k++) memcpy(&img[o_v + 16 * (size_t)k], verts + 3 * (size_t)k, 12);     /* xyz, w = 0 */
    memcpy(&img[o_si], slot_info.data(), 8 * (size_t)ns);
    memcpy(&img[o_nbr], neighbours, 4 * (size_t)ns);
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    unsigned char *d = nullptr;
    bq2_status st = arena_alloc(ctx, total, &d);
    if (st != BQ2_OK) return st;
    CK(cudaMemcpyAsync(d, img.data(), total, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpy(brush)");
    CK(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize(brush)");
    bq2_brush_dev_h bd{};
    bd.slot_info = d + o_si; bd.nbr = (const int32_t *)(d + o_nbr); bd.verts = d + o_v; bd.n_polys = n_polys; bd.n_slots = (int32_t)ns;
    ctx->brush_dev.push_back(bd);
    bq2_ctx::Brush b; b.n_polys = n_polys; b.n_slots = (int)ns; b.worst_indices = worst;
    ctx->brushes.push_back(b);
    CK(ctx->d_brushes.ensure(ctx->brush_dev.size()), "cudaMalloc(brush table)");
    CK(cudaMemcpy(ctx->d_brushes.p, ctx->brush_dev.data(), ctx->brush_dev.size() * sizeof(bq2_brush_dev_h), cudaMemcpyHostToDevice), "cudaMemcpy(brush table)");
    *brush_id = (int32_t)ctx->brushes.size() - 1;
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

bq2_status bq2_brush_volumes(bq2_ctx *ctx, const bq2_brush_pair *pairs, int32_t n_pairs, const uint8_t *lit, bq2_brush_out *out)
try {
    if (!ctx) return BQ2_E_INVALID;
    worker_drain(ctx);
    if (n_pairs < 0 || (n_pairs && (!pairs || !lit)) || !out) return fail(ctx, BQ2_E_INVALID, "bq2_brush_volumes: bad arguments");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    /* pair records and the returned counts live in page-locked memory of the ctx: the copies are DMA, not staged by the driver */
    const size_t o_lit = (((size_t)n_pairs * sizeof(bq2_brush_pair_dev_h)) + 15) & ~(size_t)15;
    CK(ctx->h_brush.ensure(o_lit + 8 * (size_t)n_pairs + 16), "cudaMallocHost(brush pairs)");
    bq2_brush_pair_dev_h *hp = (bq2_brush_pair_dev_h *)ctx->h_brush.p;
    uint32_t *h_counts = (uint32_t *)(ctx->h_brush.p + o_lit);
    uint32_t lit_bytes = 0, vstride = 4, istride = 4; int max_polys = 1, max_slots = 1;
    for (int k = 0; k < n_pairs; k++) {
        if (pairs[k].brush < 0 || pairs[k].brush >= (int32_t)ctx->brushes.size()) return fail(ctx, BQ2_E_INVALID, "bq2_brush_volumes: unknown brush");
        const bq2_ctx::Brush &b = ctx->brushes[pairs[k].brush];
        hp[k].brush = pairs[k].brush; hp[k].scale = pairs[k].scale; hp[k].lit_off = lit_bytes;
        for (int c = 0; c < 3; c++) hp[k].light[c] = pairs[k].light[c];
        lit_bytes += (uint32_t)b.n_polys;
        vstride = std::max(vstride, (uint32_t)b.n_slots); istride = std::max(istride, (uint32_t)b.worst_indices);
        max_polys = std::max(max_polys, b.n_polys); max_slots = std::max(max_slots, b.n_slots);
    }
    CK(ctx->d_brush_in.ensure(o_lit + lit_bytes + 16), "cudaMalloc(brush pairs)");
    istride = (istride + 1u) & ~1u;                      /* index rows start 8-byte aligned (the kernel stores index pairs) */
    CK(ctx->brush_vcache.ensure(6 * (size_t)n_pairs * vstride + 4), "cudaMalloc(brush vcache)");
    CK(ctx->brush_icache.ensure((size_t)n_pairs * istride + 4), "cudaMalloc(brush icache)");
    CK(ctx->brush_counts.ensure(2 * (size_t)n_pairs + 2), "cudaMalloc(brush counts)");
    ctx->h_brush_counts = h_counts;
    if (n_pairs) {
        /* the lit bytes: straight from the caller's array when that is page-locked memory of this ctx (bq2_host_alloc),
           else through a page-locked staging buffer */
        const unsigned char *lit_src = lit;
        if (!host_pinned(ctx, lit, lit_bytes)) {
            CK(ctx->h_brush_lit.ensure((size_t)lit_bytes + 16), "cudaMallocHost(lit)");
            memcpy(ctx->h_brush_lit.p, lit, lit_bytes);
            lit_src = ctx->h_brush_lit.p;
        }
        CK(cudaMemcpyAsync(ctx->d_brush_in.p, hp, (size_t)n_pairs * sizeof(bq2_brush_pair_dev_h), cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpy(brush pairs)");
        CK(cudaMemcpyAsync(ctx->d_brush_in.p + o_lit, lit_src, lit_bytes, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpy(lit)");
        CK((cudaError_t)bq2_launch_brush(ctx->d_brushes.p, ctx->d_brush_in.p, ctx->d_brush_in.p + o_lit, ctx->brush_vcache.p, ctx->brush_icache.p,
                                         ctx->brush_counts.p, vstride, istride, n_pairs, max_polys, max_slots, ctx->stream), "launch(bq2_brush_kernel)");
        ctx->launches++;
        CK(cudaMemcpyAsync(h_counts, ctx->brush_counts.p, 8 * (size_t)n_pairs, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpy(brush counts)");
    }
    CK(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize(brush)");
    ctx->brush_vstride = vstride; ctx->brush_istride = istride; ctx->brush_pairs = n_pairs;
    out->vcache = ctx->brush_vcache.p; out->icache = ctx->brush_icache.p; out->counts = ctx->h_brush_counts;
    out->vert_stride = vstride; out->index_stride = istride;
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

bq2_status bq2_brush_download(bq2_ctx *ctx, int32_t pair, float *vcache, uint32_t *icache)
try {
    if (!ctx || !vcache || !icache) return BQ2_E_INVALID;
    if (pair < 0 || pair >= ctx->brush_pairs) return fail(ctx, BQ2_E_INVALID, "bq2_brush_download: bad pair");
    const uint32_t vb = ctx->h_brush_counts[2 * (size_t)pair], ib = ctx->h_brush_counts[2 * (size_t)pair + 1];
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    if (vb) CK(cudaMemcpy(vcache, ctx->brush_vcache.p + 6 * (size_t)pair * ctx->brush_vstride, 24 * (size_t)vb, cudaMemcpyDeviceToHost), "cudaMemcpy(vcache)");
    if (ib) CK(cudaMemcpy(icache, ctx->brush_icache.p + (size_t)pair * ctx->brush_istride, 4 * (size_t)ib, cudaMemcpyDeviceToHost), "cudaMemcpy(icache)");
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

void *bq2_host_alloc(bq2_ctx *ctx, size_t bytes)
{
    if (!ctx || !bytes) return nullptr;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { fail_cuda(ctx, e, "cudaHostAlloc"); return nullptr; }
    try { ctx->host_ranges.push_back(bq2_ctx::HostRange{(const unsigned char *)p, bytes}); }
    catch (...) { cudaFreeHost(p); guard_exception(ctx); return nullptr; }
    return p;
}

void bq2_host_free(bq2_ctx *ctx, void *p)
{
    if (!ctx || !p) return;
    worker_drain(ctx);
    for (size_t i = 0; i < ctx->host_ranges.size(); i++)
        if (ctx->host_ranges[i].base == (const unsigned char *)p) {
            for (FrameState &f : ctx->fs) if (f.stream) cudaStreamSynchronize(f.stream);    /* a copy may still read it */
            cudaFreeHost(p);
            ctx->host_ranges.erase(ctx->host_ranges.begin() + (long)i);
            return;
        }
}

static bool host_pinned(const bq2_ctx *ctx, const void *p, size_t bytes)
{
    const unsigned char *q = (const unsigned char *)p;
    for (const bq2_ctx::HostRange &r : ctx->host_ranges)
        if (q >= r.base && q + bytes <= r.base + r.bytes) return true;
    return false;
}

/* ---- the stream operations of a world frame, from what bq2_world_submit left in F.job ---------------------- */
/* the AngleVectors of the frame's rotated entities (their rows hold the angles until now) */
static void finish_bases(FrameState &F)
{
    FrameState::WorldJob &J = F.job;
    if (J.n_rot) { bq2_host_bases_from_angles((float *)(F.hrec.p + J.off_b), J.n_rot); J.n_rot = 0; }
}

static void stage_caller_arrays(FrameState &F)
{
    const FrameState::WorldJob &J = F.job;
    unsigned char *h = F.hrec.p;
    if (J.nL) memcpy(h + J.off_lights, J.lights, (size_t)J.nL * 16);
    if (J.nP) { memcpy(h + J.off_pe, J.pair_ent, (size_t)J.nP * 4); memcpy(h + J.off_pl, J.pair_light, (size_t)J.nP * 4); }
}

static bq2_status world_enqueue(bq2_ctx *ctx, FrameState &F)
{
    const FrameState::WorldJob &J = F.job;
    bq2_world_launch &Wl = F.wlaunch;
    PinBuf &h_records = F.hrec, &h_counts = F.hcnt;
    StageTimer *st = (ctx->trace.gpu && !tl_err_sink) ? &g_stage_timer : nullptr;
    auto enqueue = [&]() -> bq2_status {
        if (st) st->mark(0, F.stream);
        if (J.direct) {
            unsigned char *db = F.d_records.p;
            CK(cudaMemcpyAsync(db + J.off_xf, h_records.p + J.off_xf, J.off_lights - J.off_xf, cudaMemcpyHostToDevice, F.stream), "cudaMemcpy(entity rows)");
            CK(cudaMemcpyAsync(db + J.off_lights, J.lights, (size_t)J.nL * 16, cudaMemcpyHostToDevice, F.stream), "cudaMemcpy(lights)");
            CK(cudaMemcpyAsync(db + J.off_pe, J.pair_ent, (size_t)J.nP * 4, cudaMemcpyHostToDevice, F.stream), "cudaMemcpy(pair_ent)");
            CK(cudaMemcpyAsync(db + J.off_pl, J.pair_light, (size_t)J.nP * 4, cudaMemcpyHostToDevice, F.stream), "cudaMemcpy(pair_light)");
            if (J.basis_bytes) CK(cudaMemcpyAsync(db + J.off_b, h_records.p + J.off_b, J.basis_bytes, cudaMemcpyHostToDevice, F.stream), "cudaMemcpy(bases)");
        } else if (J.copy_end > J.off_xf)
            CK(cudaMemcpyAsync(F.d_records.p + J.off_xf, h_records.p + J.off_xf, J.copy_end - J.off_xf, cudaMemcpyHostToDevice, F.stream), "cudaMemcpy(world records)");
        if (st) st->mark(1, F.stream);
        Wl.n_pairs = (J.want & BQ2_WANT_SHADOW) ? J.nP : 0;         /* the entity records are always built */
        /* totals, pairs per record and chain heads were left cleared by the frame in front (J.clean), else clear them */
        if (!Wl.n_pairs) {  /* no pair thread runs: nothing writes the per-pair counts, clear them too */
            const size_t words = J.clean ? J.tot_word : J.cnt_word + 2 * (size_t)J.nE;
            if (words) CK(cudaMemsetAsync(F.index_count.p, 0, sizeof(uint32_t) * words, F.stream), "cudaMemset(counts, totals, pair lists)");
        } else if (!J.clean)
            CK(cudaMemsetAsync(Wl.totals, 0, sizeof(uint32_t) * (8 + 2 * (size_t)J.nE), F.stream), "cudaMemset(totals, pair lists)");
        if (J.nP || J.nE) {
            CK((cudaError_t)bq2_launch_world_setup(&Wl, F.stream), "launch(bq2_world_link_kernel)");
            COUNT_LAUNCH(ctx, bq2_world_setup_launches(&Wl));
        }
        if (st) st->mark(2, F.stream);
        const int pdl = ctx->pdl_world && !st;
        bq2_status ls = launch_kernels(ctx, F, pdl);
        if (ls != BQ2_OK) return ls;
        if (st) st->mark(3, F.stream);
        /* counts of a small frame are stored to the result buffer by the last kernel itself (5.4 us behind a kernel
           against 7.2 us for a copy node); beyond 64 KB the copy engine is the faster way over PCIe (a kernel's stores to
           host memory run at about a third of its rate) and the last kernel hands over the totals only */
        const uint32_t copy_from = J.tot_word > BQ2_TAIL_STORE_MAX_WORDS ? (uint32_t)J.tot_word : 0u;
        CK((cudaError_t)bq2_launch_frame_tail(F.index_count.p, F.hcnt_dev, F.keep_cnt, copy_from, (uint32_t)J.tot_word + 8u, (uint32_t)J.tot_word,
                                              (uint32_t)J.cnt_word, (uint32_t)(J.cnt_word + (size_t)J.nE), F.stream, pdl),
           "launch(bq2_frame_tail_kernel)");
        if (copy_from)
            CK(cudaMemcpyAsync(h_counts.p, F.index_count.p, sizeof(uint32_t) * J.tot_word, cudaMemcpyDeviceToHost, F.stream), "cudaMemcpy(counts)");
        COUNT_LAUNCH(ctx, 1);
        if (st) { st->mark(4, F.stream); st->finish(); }
        return BQ2_OK;
    };
    const bq2_status es = enqueue_frame(ctx, F, J.use_key ? &J.key : nullptr, enqueue);
    if (es != BQ2_OK) F.tail_clean = false;
    return es;
}

/* ---- submit worker ------------------------------------------------------------------------------------------
 * One thread per ctx, started by the first world frame that can use it.  The caller's thread posts the number of a
 * frame state whose job is complete; the worker stages the caller's page-locked arrays into the state's pinned record
 * buffer and issues the frame's stream operations (a graph launch in the steady state).  bq2_frame_wait -- and every
 * other entry point that touches frame states or CUDA objects of the ctx -- first waits for the worker to have
 * finished what was posted.  The worker spins for a short while between frames (a frame loop posts every few
 * microseconds) and then sleeps on a condition variable. */
/* Best effort, Linux: keep the worker off the physical core the caller's thread is running on (a spinning hyper-thread
   sibling slows the caller's per-entity loop by a third).  caller_cpu = sched_getcpu() of the thread that started it. */
static void worker_avoid_core(int caller_cpu)
{
#if defined(__linux__)
    cpu_set_t allowed, mine;
    if (caller_cpu < 0 || sched_getaffinity(0, sizeof allowed, &allowed) != 0) return;
    char path[96], buf[128];
    snprintf(path, sizeof path, "/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list", caller_cpu);
    mine = allowed;
    CPU_CLR(caller_cpu, &mine);
    FILE *f = fopen(path, "r");
    if (f) {
        if (fgets(buf, sizeof buf, f)) {
            for (char *p = buf; *p;) {                       /* "a-b,c,..." */
                char *end; long lo = strtol(p, &end, 10), hi = lo;
                if (end == p) break;
                if (*end == '-') { hi = strtol(end + 1, &end, 10); }
                for (long c = lo; c <= hi && c < CPU_SETSIZE; c++) CPU_CLR((int)c, &mine);
                p = (*end == ',') ? end + 1 : end;
                if (*end != ',') break;
            }
        }
        fclose(f);
    }
    if (CPU_COUNT(&mine) > 0) sched_setaffinity(0, sizeof mine, &mine);
#else
    (void)caller_cpu;
#endif
}

static void worker_main(bq2_ctx *ctx, int caller_cpu)
{
    worker_avoid_core(caller_cpu);
    cudaSetDevice(ctx->device);
    uint64_t tail = 0;
    for (;;) {
        int spins = 0;
        while (ctx->wq_head.load(std::memory_order_acquire) == tail) {
            if (ctx->wstop.load(std::memory_order_acquire)) return;
            if (++spins < 20000) {
#if defined(__x86_64__) || defined(__i386__)
                __builtin_ia32_pause();
#endif
                continue;
            }
            std::unique_lock<std::mutex> lk(ctx->wmx);
            ctx->wsleeping.store(true, std::memory_order_seq_cst);
            if (ctx->wq_head.load(std::memory_order_seq_cst) == tail && !ctx->wstop.load())
                ctx->wcv.wait_for(lk, std::chrono::milliseconds(50));
            ctx->wsleeping.store(false, std::memory_order_seq_cst);
            spins = 0;
        }
        const int k = ctx->wq[tail % (BQ2_FRAME_STATES + 1)].load(std::memory_order_relaxed);
        FrameState &F = ctx->fs[k];
        tl_err_sink = &F.job.err;
        bq2_status st = BQ2_OK;
        try {
            finish_bases(F);
            if (F.job.stage) stage_caller_arrays(F);
            st = world_enqueue(ctx, F);
        } catch (...) { st = BQ2_E_NOMEM; F.job.err = "submit worker: out of host memory"; }
        tl_err_sink = nullptr;
        F.job.status = st;
        tail++;
        ctx->wq_tail.store(tail, std::memory_order_release);
        F.job_done.store(F.job_posted.load(std::memory_order_relaxed), std::memory_order_release);
    }
}

/* start the worker, or end its sleep, so that the frames that follow find it spinning */
static void worker_wake(bq2_ctx *ctx)
{
    if (!ctx->worker.joinable()) {
        int cpu = -1;
#if defined(__linux__)
        cpu = sched_getcpu();
#endif
        try { ctx->worker = std::thread(worker_main, ctx, cpu); } catch (...) { ctx->use_worker = false; }
        return;
    }
    std::lock_guard<std::mutex> lk(ctx->wmx);
    ctx->wcv.notify_one();
}

static bq2_status worker_post(bq2_ctx *ctx, int k)
{
    FrameState &F = ctx->fs[k];
    F.job.err.clear();
    F.job_posted.store(F.job_posted.load(std::memory_order_relaxed) + 1, std::memory_order_relaxed);
    const uint64_t head = ctx->wq_head.load(std::memory_order_relaxed);
    ctx->wq[head % (BQ2_FRAME_STATES + 1)].store(k, std::memory_order_relaxed);
    ctx->wq_head.store(head + 1, std::memory_order_seq_cst);
    if (ctx->wsleeping.load(std::memory_order_seq_cst)) { std::lock_guard<std::mutex> lk(ctx->wmx); ctx->wcv.notify_one(); }
    return BQ2_OK;
}

/* the worker has issued everything that was posted (entry points that touch frame states call this first) */
static void worker_drain(bq2_ctx *ctx)
{
    if (!ctx || !ctx->worker.joinable()) return;
    const uint64_t head = ctx->wq_head.load(std::memory_order_acquire);
    while (ctx->wq_tail.load(std::memory_order_acquire) != head) {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
}

/* ... for ONE frame state: bq2_frame_wait meets the worker here; an error of the job becomes the caller's */
static bq2_status worker_join_frame(bq2_ctx *ctx, FrameState &F)
{
    if (!ctx->worker.joinable()) return BQ2_OK;
    const uint64_t posted = F.job_posted.load(std::memory_order_relaxed);
    while (F.job_done.load(std::memory_order_acquire) != posted) {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    if (F.job.status != BQ2_OK) { ctx->err = F.job.err; const bq2_status st = F.job.status; F.job.status = BQ2_OK; return st; }
    return BQ2_OK;
}

bq2_status bq2_world_submit(bq2_ctx *ctx, const bq2_world_desc *wd)
try {
    if (!ctx) return BQ2_E_INVALID;
    if (!wd || wd->n_ents < 0 || wd->n_pairs < 0 || wd->n_lights < 0 || (wd->n_ents > 0 && !wd->ents) ||
        (wd->n_pairs > 0 && (!wd->pair_ent || !wd->pair_light || !wd->lights)))
        return fail(ctx, BQ2_E_INVALID, "bq2_world: bad descriptor");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    if (ctx->submit_seq - ctx->wait_seq >= BQ2_MAX_IN_FLIGHT) return fail(ctx, BQ2_E_INVALID, "bq2_world_submit: the maximum number of frames is already in flight (bq2_max_frames_in_flight), call bq2_frame_wait first");
    const int slot_k = pick_state(ctx);
    if (slot_k < 0) return fail(ctx, BQ2_E_INVALID, "bq2_frame_submit: no free frame state");
    FrameState &F = ctx->fs[slot_k];
    PinBuf &h_records = F.hrec, &h_counts = F.hcnt;
    const int nE = wd->n_ents, nP = wd->n_pairs, nL = wd->n_lights;
    const int nModels = (int)ctx->models.size();
    if (wd->want & ~BQ2_WANT_PUBLIC_MASK) return fail(ctx, BQ2_E_INVALID, "bq2_world: unknown bits in want");
    const uint32_t want = wd->want;

    if (ctx->model_rows.size() != ctx->models.size()) {
        ctx->model_rows.resize(ctx->models.size());
        for (size_t i = 0; i < ctx->models.size(); i++) {
            const Model &mo = ctx->models[i];
            const bq2_dev_mesh &me = ctx->meshes[mo.first_mesh];
            ctx->model_rows[i] = bq2_model_row{mo.F, mo.md3 ? 1 : 0, mo.first_mesh, mo.n_mesh == 1 ? 1 : 2,
                                               me.V, me.T, me.Tp, me.flags};
            bq2_model_row_derive(&ctx->model_rows[i]);
            ctx->model_rows[i].pad[0] = (int32_t)mo.rf_flags;
            ctx->model_rows[i].pad[1] = mo.casts[0] ? 0 : 1;     /* a mesh whose skin casts no shadow (M:63811-63816): lerp, no volume */
        }
    }
    const size_t off_xf = (size_t)nE * sizeof(bq2_dev_ent), off_view = off_xf + (size_t)nE * sizeof(bq2_dev_went),
                 off_lights = off_view + 16,
                 off_pe = off_lights + (size_t)nL * 16, off_pl = off_pe + (((size_t)nP * 4 + 15) & ~(size_t)15),
                 off_b = off_pl + (((size_t)nP * 4 + 15) & ~(size_t)15),
                 rec_bytes = off_b + (((size_t)nE * 36 + 15) & ~(size_t)15);
    /* [0, off_xf) = the entity records: device only, written by the record-building kernel; the copy starts at off_xf.
       [off_b, ..) = the AngleVectors bases of the rotated entities, room for all of them, copied as far as they go */
    Trace &tr = ctx->trace; tr.start(); tr.n++;
    bq2_world_stats ws{};
    int er = -3;
    const bool want_tb = (want & (BQ2_WANT_NTB | BQ2_WANT_TSLIGHT)) != 0;
    if (!want_tb || (want & BQ2_WANT_NTB)) {                     /* (TSLIGHT without NTB: the host-built route reports it) */
        CK(h_records.ensure(std::max<size_t>(rec_bytes, 16)), "cudaMallocHost(records)");
        F.ent_slot_first.resize((size_t)nE + 1); F.ent_vert_off.resize((size_t)nE + 1); F.ent_tb_off.assign((size_t)nE + 1, 0u);
        F.slot_record.resize((size_t)nE + 1);
        er = bq2_host_world_entities(wd->ents, nE, ctx->model_rows.data(), nModels, wd->opts, (bq2_dev_went *)(h_records.p + off_xf),
                                     (float *)(h_records.p + off_b), F.slot_record.data(), F.ent_vert_off.data(), &ws);
        if (er == -1) return fail(ctx, BQ2_E_INVALID, "bq2_world: entity refers to an unknown model");
        if (er == -2) return fail(ctx, BQ2_E_INVALID, "bq2_world: frame out of range");
        if (er == -4) return fail(ctx, BQ2_E_LIMIT, "bq2_world: more than 2^32 output rows, split the frame");
        /* tangent-space outputs on this route: every model with per-VERTEX N/T/B rows (MD3, smooth MD2 with tangent /
           binormal bytes), so that an entity's N/T/B rows and a pair's LightP / tsH rows are numbered like its vertex rows */
        if (er == 0 && want_tb && (ws.and_mflags & (BQ2_MESH_SMOOTH | BQ2_MESH_HAS_TB)) != (BQ2_MESH_SMOOTH | BQ2_MESH_HAS_TB)) er = -3;
        if (er == 0 && want_tb) F.ent_tb_off = F.ent_vert_off;
    }
    if (er != 0) {
        /* multi-mesh models, or tangent-space outputs for models with per-triangle rows (non-smooth MD2) or without
           tangent / binormal bytes: host-built records, same slot numbering (filtered pairs
           keep their slot) */
        ctx->tmp_ents.resize((size_t)nE + 1); ctx->tmp_pairs.resize((size_t)nP + 1);
        int32_t kept = 0;
        bq2_status s = build_records_impl(ctx, wd->ents, nE, wd->lights, nL, wd->pair_ent, wd->pair_light, nP, wd->view_world,
                                          wd->opts, ctx->tmp_ents.data(), ctx->tmp_pairs.data(), &kept, true);
        if (s != BQ2_OK) return s;
        bq2_frame_desc fd{};
        fd.ents = ctx->tmp_ents.data(); fd.n_ents = nE; fd.pairs = ctx->tmp_pairs.data(); fd.n_pairs = kept;
        fd.want = want & ~BQ2_WANT_TABLES; fd.index_stride = wd->index_stride;
        return bq2_frame_submit(ctx, &fd);
    }
    tr.lap(2);
    CK(F.d_records.ensure(std::max<size_t>(rec_bytes, 16)), "cudaMalloc(records)");
    for (int e = 0; e <= nE; e++) F.ent_slot_first[e] = e;
    F.k_warp = FrameState::KLaunch(); F.k_team = FrameState::KLaunch();
    F.k_warp.n = ws.n_warp; F.k_warp.max_V = ws.warp_max_V; F.k_warp.max_T = ws.warp_max_T; F.k_warp.any_md2 = ws.warp_md2;
    F.k_team.n = ws.n_team; F.k_team.max_V = ws.team_max_V; F.k_team.max_T = ws.team_max_T; F.k_team.any_md2 = ws.team_md2;
    const uint32_t voff = ws.lerped_verts_padded, maxVpad = ws.max_vpad, maxBitw = ws.max_bitw;
    /* rows are numbered in 32 bits on the device (pair p's rows start at p * stride) */
    if (((uint64_t)nP * maxVpad | (uint64_t)nP * maxBitw) >> 32)
        return fail(ctx, BQ2_E_LIMIT, "bq2_world: more than 2^32 output rows, split the frame");
    const int maxT = ws.max_T;
    const uint64_t total_verts = ws.total_verts;
    {
        float *hv = (float *)(h_records.p + off_view);
        for (int k = 0; k < 3; k++) hv[k] = wd->view_world ? wd->view_world[k] : 0.0f;
        hv[3] = wd->view_world ? 1.0f : 0.0f;
    }
    /* the caller's arrays: staged through the pinned record buffer, or -- when all of them live in memory from
       bq2_host_alloc -- copied to the device straight from where they are (the caller then keeps them unchanged until
       the frame has been waited for) */
    /* (below ~1 MB one staged copy is cheaper than five separate ones: 25.6 vs 29.1 us per frame at C2) */
    const bool direct = nE && nL && nP && !ctx->host_ranges.empty() && !tr.gpu &&
                        (size_t)nL * 16 + (size_t)nP * 8 + (size_t)nE * sizeof(bq2_world_entity) >= ctx->direct_min_bytes &&
                        host_pinned(ctx, wd->ents, (size_t)nE * sizeof(bq2_world_entity)) && host_pinned(ctx, wd->lights, (size_t)nL * 16) &&
                        host_pinned(ctx, wd->pair_ent, (size_t)nP * 4) && host_pinned(ctx, wd->pair_light, (size_t)nP * 4);
    /* staged: by the submit worker when the caller's arrays are page-locked memory of this ctx (they stay valid and
       unchanged until bq2_frame_wait, as for the direct copies), by this thread otherwise */
    const bool caller_arrays_stay = !direct && nL && nP && !ctx->host_ranges.empty() &&
                                    host_pinned(ctx, wd->lights, (size_t)nL * 16) && host_pinned(ctx, wd->pair_ent, (size_t)nP * 4) &&
                                    host_pinned(ctx, wd->pair_light, (size_t)nP * 4);
    F.job.stage = false;
    if (!direct) {
        if (caller_arrays_stay && ctx->use_worker && !tr.gpu && !(want & BQ2_WANT_TABLES)) F.job.stage = true;
        else {
            if (nL) memcpy(h_records.p + off_lights, wd->lights, (size_t)nL * 16);
            if (nP) { memcpy(h_records.p + off_pe, wd->pair_ent, (size_t)nP * 4); memcpy(h_records.p + off_pl, wd->pair_light, (size_t)nP * 4); }
        }
    }
    /* the bases are copied in steps of 64 rows, so that frames whose number of rotated entities differs a little
       still have the same copy sizes (the graph key below) */
    const size_t basis_bytes = std::min((((size_t)ws.n_rotated + 63) & ~(size_t)63) * 36, rec_bytes - off_b);
    const size_t copy_end = basis_bytes ? off_b + basis_bytes : off_b;

    tr.lap(3);
    uint32_t stride = wd->index_stride ? wd->index_stride : 12u * (uint32_t)maxT;
    stride = (stride + 3u) & ~3u;
    const bool tables = (want & BQ2_WANT_TABLES) != 0;

    /* ---- capacity: per-pair rows are bounded by the largest mesh of the frame ------------------------ */
    CK(F.lerped.ensure(3 * (size_t)voff + 4), "cudaMalloc(lerped)");
    if (want & BQ2_WANT_SHADOW) {
        CK(F.extruded.ensure(3 * (size_t)nP * maxVpad + 4), "cudaMalloc(extruded)");
        CK(F.facing_bits.ensure((size_t)nP * maxBitw + 1), "cudaMalloc(facing_bits)");
        CK(F.indices.ensure((size_t)nP * stride + 4), "cudaMalloc(indices)");
    }
    if (ws.team_normal_rows) CK(F.tri_normals.ensure(4 * (size_t)ws.team_normal_rows + 4), "cudaMalloc(triangle normals)");
    if (want & BQ2_WANT_NTB) {
        CK(F.normals.ensure(3 * (size_t)voff + 4), "cudaMalloc(normals)");
        CK(F.tangents.ensure(3 * (size_t)voff + 4), "cudaMalloc(tangents)");
        CK(F.binormals.ensure(3 * (size_t)voff + 4), "cudaMalloc(binormals)");
    }
    if (want & BQ2_WANT_TSLIGHT) {
        CK(F.lightp.ensure(3 * (size_t)nP * maxVpad + 4), "cudaMalloc(lightp)");
        CK(F.tsh.ensure(3 * (size_t)nP * maxVpad + 4), "cudaMalloc(tsh)");
    }
    /* one buffer: index counts | 8-byte aligned, three 64-bit totals | pairs per record | overflow-chain heads | the
       finished frame's pairs per record.  The frame's last kernel hands counts + totals to the host and clears totals
       and pairs per record for the next frame; a frame that does not find them cleared clears totals .. heads itself
       (heads to 0: a chain is only walked for as many nodes as were pushed). */
    const size_t tot_word = ((size_t)nP + 1) & ~(size_t)1, cnt_word = tot_word + 8;
    CK(F.index_count.ensure(cnt_word + 3 * (size_t)nE + 8), "cudaMalloc(index_count)");
    CK(h_counts.ensure(sizeof(uint32_t) * (tot_word + 8)), "cudaMallocHost(counts)");
    if (F.hcnt_dev_of != h_counts.p) {
        void *dp = nullptr;
        CK(cudaHostGetDevicePointer(&dp, h_counts.p, 0), "cudaHostGetDevicePointer(counts)");
        F.hcnt_dev = (uint32_t *)dp; F.hcnt_dev_of = h_counts.p;
    }
    const bool clean = F.tail_clean && F.clean_ptr == F.index_count.p && F.clean_nP == nP && F.clean_nE == nE;
    const size_t w_pairs = ((size_t)nP * sizeof(bq2_dev_pair) + 15) & ~(size_t)15, w_bucket = (size_t)nE * BQ2_BUCKET * 4,
                 w_next = (size_t)nP * 4;
    CK(F.d_world.ensure(w_pairs + w_bucket + w_next + 64), "cudaMalloc(world scratch)");
    if (ctx->meshes_dirty) {
        CK(ctx->d_meshes.ensure(ctx->meshes.size()), "cudaMalloc(mesh table)");
        CK(cudaMemcpyAsync(ctx->d_meshes.p, ctx->meshes.data(), ctx->meshes.size() * sizeof(bq2_dev_mesh),
                           cudaMemcpyHostToDevice, F.stream), "cudaMemcpy(mesh table)");
        CK(cudaStreamSynchronize(F.stream), "sync(mesh table)");
        ctx->meshes_dirty = false;
    }
    if (ctx->dev_models.size() != ctx->models.size()) {          /* device model table: frame headers + first mesh */
        ctx->dev_models.resize(ctx->models.size());
        for (size_t i = 0; i < ctx->models.size(); i++)
            ctx->dev_models[i] = bq2_dev_model{ctx->models[i].d_hdr, ctx->models[i].first_mesh, ctx->models[i].md3 ? 1 : 0};
        CK(ctx->d_models.ensure(ctx->dev_models.size()), "cudaMalloc(model table)");
        CK(cudaMemcpyAsync(ctx->d_models.p, ctx->dev_models.data(), ctx->dev_models.size() * sizeof(bq2_dev_model),
                           cudaMemcpyHostToDevice, F.stream), "cudaMemcpy(model table)");
        CK(cudaStreamSynchronize(F.stream), "sync(model table)");
    }
    if (F.k_team.n && bq2_kernel_smem_bytes(F.k_team.max_V, F.k_team.max_T, F.k_team.any_md2) > 227u * 1024u)
        return fail(ctx, BQ2_E_LIMIT, "bq2_world: mesh too large for one CTA's shared memory");

    /* ---- go ----------------------------------------------------------------------------------------- */
    tr.lap(5);
    unsigned char *dw = F.d_world.p;
    bq2_world_launch &Wl = F.wlaunch;
    /* (reading the pinned host records directly from the kernels instead of copying them was measured: the copy
       stage disappears but the record-building kernels and the frame kernel get 13 + 6 us slower at C2) */
    unsigned char *rb = F.d_records.p;
    Wl.went = (const bq2_dev_went *)(rb + off_xf);
    Wl.bases = (const float *)(rb + off_b);
    Wl.models = ctx->d_models.p; Wl.meshes = ctx->d_meshes.p; Wl.ents_out = (bq2_dev_ent *)rb;
    Wl.lights = (const float *)(rb + off_lights);
    Wl.pair_ent = (const int32_t *)(rb + off_pe); Wl.pair_light = (const int32_t *)(rb + off_pl);
    Wl.n_pairs = nP; Wl.n_ents = nE; Wl.n_lights = nL; Wl.n_recs = nE;
    Wl.dpairs = (bq2_dev_pair *)dw;
    Wl.bucket = (uint32_t *)(dw + w_pairs); Wl.onext = (uint32_t *)(dw + w_pairs + w_bucket);
    Wl.totals = (unsigned long long *)(F.index_count.p + tot_word);
    Wl.bucket_cnt = F.index_count.p + cnt_word; Wl.ohead = Wl.bucket_cnt + nE;
    F.keep_cnt = Wl.ohead + nE;
    Wl.vert_stride = maxVpad; Wl.bits_stride = maxBitw;
    Wl.index_count = F.index_count.p;
    Wl.view = (const float *)(rb + off_view);

    bq2_launch &L = F.launch;
    L.meshes = ctx->d_meshes.p;
    L.ents = (const bq2_dev_ent *)rb;
    L.pairs = Wl.dpairs;
    L.n_ent_slots = nE; L.want = want & ~BQ2_WANT_TABLES; L.index_stride = stride;
    L.lerped = F.lerped.p; L.extruded = F.extruded.p; L.facing_bits = F.facing_bits.p;
    L.indices = F.indices.p; L.index_count = F.index_count.p;
    const bool w_ntb = (want & BQ2_WANT_NTB) != 0, w_ts = (want & BQ2_WANT_TSLIGHT) != 0;
    L.normals = w_ntb ? F.normals.p : nullptr; L.tangents = w_ntb ? F.tangents.p : nullptr; L.binormals = w_ntb ? F.binormals.p : nullptr;
    L.lightp = w_ts ? F.lightp.p : nullptr; L.tsh = w_ts ? F.tsh.p : nullptr;
    L.totals = Wl.totals; L.tri_normals = F.tri_normals.p;
    L.bucket_cnt = Wl.bucket_cnt; L.bucket = Wl.bucket; L.ohead = Wl.ohead; L.onext = Wl.onext; L.rec_base = 0;
    F.n_ent_slots = nE; F.n_pair_slots = nP; F.n_ents = nE; F.n_pairs = nP;
    F.total_verts = total_verts;
    F.world_mode = true; F.world_tables = tables;
    /* per-slot host tables used by bq2_expand_trisverts */
    F.slot_mesh_V.resize((size_t)nP); F.slot_mesh_T.resize((size_t)nP); F.pair_slot_entslot.resize((size_t)nP);

    /* the frame's stream operations: H2D, record building, the fused kernel(s), the kernel that stores counts + totals
       to the page-locked result buffer -- issued by the
       submit worker when the ctx has one and the caller's arrays may be read after this call returns (page-locked
       memory from bq2_host_alloc: the contract already says they stay unchanged until bq2_frame_wait), else here */
    {
        FrameState::WorldJob &J = F.job;
        J.direct = direct; J.lights = wd->lights; J.pair_ent = wd->pair_ent; J.pair_light = wd->pair_light;
        J.off_xf = off_xf; J.off_lights = off_lights; J.off_pe = off_pe; J.off_pl = off_pl; J.off_b = off_b;
        J.basis_bytes = basis_bytes; J.copy_end = copy_end; J.tot_word = tot_word; J.cnt_word = cnt_word;
        J.nE = nE; J.nL = nL; J.nP = nP; J.want = want; J.status = BQ2_OK; J.n_rot = ws.n_rotated;
        J.use_key = !(tr.gpu || tables);
        J.clean = clean;
        F.tail_clean = true; F.clean_ptr = F.index_count.p; F.clean_nP = nP; F.clean_nE = nE;    /* (world_enqueue takes it back if it fails) */
        const int32_t kv[20] = {nE, nP, nL, (int32_t)want, (int32_t)stride, ws.n_warp, ws.n_team, ws.warp_max_V, ws.warp_max_T, ws.warp_md2,
                                ws.team_max_V, ws.team_max_T, ws.team_md2, (int32_t)maxVpad, (int32_t)maxBitw, (int32_t)basis_bytes, clean ? 1 : 0, 0, 0, 0};
        memset(&J.key, 0, sizeof J.key);
        memcpy(J.key.v, kv, sizeof kv);
        const void *kp[24] = {direct ? wd->ents : nullptr, direct ? wd->lights : nullptr, direct ? wd->pair_ent : nullptr,
                              direct ? wd->pair_light : nullptr,
                              F.d_records.p, F.d_world.p, h_records.p, h_counts.p, F.lerped.p, F.extruded.p, F.facing_bits.p,
                              F.indices.p, F.index_count.p, ctx->d_meshes.p, F.tri_normals.p, ctx->d_models.p,
                              L.normals, L.tangents, L.binormals, L.lightp, L.tsh, nullptr, nullptr, nullptr};
        memcpy(J.key.p, kp, sizeof kp);
        /* to the worker unless it is asleep (an engine that submits once per 16 ms frame would pay the wake-up latency on
           every frame: it is woken for the frames that follow only when submits come in quick succession) */
        /* ... and only for a caller that overlaps frames: with nothing in flight there is nothing for the hand-over to
           overlap with, it would only add its latency to the frame's */
        bool to_worker = ctx->use_worker && !tr.gpu && !tables && ctx->submit_seq != ctx->wait_seq;
        if (to_worker) {
            const double now = Trace::now();
            const bool quick = now - ctx->last_submit_us < 300.0;
            ctx->last_submit_us = now;
            if (!ctx->worker.joinable() || ctx->wsleeping.load(std::memory_order_seq_cst)) {
                to_worker = false;
                if (quick) worker_wake(ctx);
            }
        }
        if (to_worker) {
            bq2_status ws_ = worker_post(ctx, slot_k);
            if (ws_ != BQ2_OK) { F.tail_clean = false; return ws_; }      /* nothing was enqueued: assume nothing about the counters */
        } else {
            finish_bases(F);
            if (J.stage) { stage_caller_arrays(F); J.stage = false; }
            bq2_status es = world_enqueue(ctx, F);
            if (es != BQ2_OK) return es;
        }
    }
    if (tables && nP) {
        F.world_pvoff.resize((size_t)nP + 1); F.world_pboff.resize((size_t)nP + 1);
        for (int p = 0; p <= nP; p++) { F.world_pvoff[p] = (uint32_t)p * maxVpad; F.world_pboff[p] = (uint32_t)p * maxBitw; }
        for (int p = 0; p < nP; p++) {                 /* test / shim convenience, not on the fast path */
            int e = wd->pair_ent[p];
            if (e < 0 || e >= nE) { F.slot_mesh_V[p] = 0; F.slot_mesh_T[p] = 0; F.pair_slot_entslot[p] = 0; continue; }
            const bq2_dev_mesh &me = ctx->meshes[ctx->models[wd->ents[e].model].first_mesh];
            F.slot_mesh_V[p] = me.V; F.slot_mesh_T[p] = me.T; F.pair_slot_entslot[p] = e;
        }
    } else if (tables) { F.world_pvoff.assign(1, 0u); F.world_pboff.assign(1, 0u); }
    ctx->submitted = true;
    { FrameState::Meta &m = F.meta; m.world = true; m.tables = tables; m.n_ent_slots = nE; m.n_pair_slots = nP;
      m.total_verts = total_verts; m.stride = stride; }
    ctx->order[ctx->submit_seq % BQ2_FRAME_STATES] = slot_k; ctx->busy[slot_k] = true;
    ctx->submit_seq++; ctx->cur = slot_k;
    tr.lap(6);
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

bq2_status bq2_world_frame(bq2_ctx *ctx, const bq2_world_desc *wd, bq2_frame_out *out)
{
    while (ctx && ctx->submit_seq != ctx->wait_seq) bq2_frame_wait(ctx, nullptr);     /* drain frames in flight */
    bq2_status s = bq2_world_submit(ctx, wd);
    if (s != BQ2_OK) return s;
    return bq2_frame_wait(ctx, out);
}

bq2_status bq2_frame_replay(bq2_ctx *ctx)
try {
    if (!ctx) return BQ2_E_INVALID;
    worker_drain(ctx);
    FrameState &F = ctx->fs[ctx->cur];
    if (!ctx->submitted || !F.n_ent_slots) return fail(ctx, BQ2_E_INVALID, "bq2_frame_replay: nothing submitted");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    F.tail_clean = false;
    if (F.world_mode) {                                          /* the kernels add to the totals: start from zero */
        CK(cudaMemsetAsync(F.wlaunch.totals, 0, 2 * sizeof(unsigned long long), F.stream), "cudaMemset(totals)");
        F.launch.bucket_cnt = F.keep_cnt;                        /* where the frame's last kernel left the pairs per record */
    }
    return launch_kernels(ctx, F, /*overlap=*/1);
} catch (...) { return guard_exception(ctx); }

static bq2_status frame_keep_impl(bq2_ctx *ctx, bq2_resident **out, bool own)
{
    if (!ctx || !out) return BQ2_E_INVALID;
    worker_drain(ctx);
    *out = nullptr;
    FrameState &F = ctx->fs[ctx->cur];
    if (!ctx->submitted || !F.n_ent_slots || ctx->submit_seq != ctx->wait_seq)
        return fail(ctx, BQ2_E_INVALID, "bq2_frame_keep: no completed frame (submit and wait first)");
    if (F.world_mode) return fail(ctx, BQ2_E_INVALID, "bq2_frame_keep: the frame must come from bq2_frame / bq2_frame_submit");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t bytes = (size_t)((const unsigned char *)(F.launch.pairs) - F.d_records.p) + (size_t)F.n_pair_slots * sizeof(bq2_dev_pair);
    bq2_resident *r = new (std::nothrow) bq2_resident();
    if (!r) return fail(ctx, BQ2_E_NOMEM, "bq2_frame_keep: out of host memory");
    cudaError_t e = r->records.ensure(std::max<size_t>(bytes, 16));
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->records.p, F.d_records.p, bytes, cudaMemcpyDeviceToDevice, F.stream);
    r->state = ctx->cur; r->L = F.launch; r->k_warp = F.k_warp; r->k_team = F.k_team;
    r->L.ents = (const bq2_dev_ent *)r->records.p;
    r->L.pairs = (const bq2_dev_pair *)(r->records.p + ((const unsigned char *)F.launch.pairs - F.d_records.p));
    if (own) {
        /* a set of result arrays of the frame's own, as large as the state's, holding the same results */
        r->own = true;
        #define BQ2_OWN(field) do { if (e == cudaSuccess && F.field.p) { e = r->field.ensure(F.field.cap); \
            if (e == cudaSuccess) e = cudaMemcpyAsync(r->field.p, F.field.p, F.field.cap * sizeof(*F.field.p), cudaMemcpyDeviceToDevice, F.stream); \
            r->L.field = r->field.p; r->result_bytes += F.field.cap * sizeof(*F.field.p); } } while (0)
        BQ2_OWN(lerped); BQ2_OWN(extruded); BQ2_OWN(facing_bits); BQ2_OWN(indices); BQ2_OWN(index_count);
        BQ2_OWN(normals); BQ2_OWN(tangents); BQ2_OWN(binormals); BQ2_OWN(lightp); BQ2_OWN(tsh);
        #undef BQ2_OWN
        /* ... and its own scratch rows for the team kernel's triangle normals (not a result: not counted) */
        if (e == cudaSuccess && F.tri_normals.p) { e = r->tri_normals.ensure(F.tri_normals.cap); r->L.tri_normals = r->tri_normals.p; }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(F.stream);
    if (e != cudaSuccess) { r->release(); delete r; return fail_cuda(ctx, e, "bq2_frame_keep"); }
    *out = r;
    return BQ2_OK;
}

bq2_status bq2_frame_keep(bq2_ctx *ctx, bq2_resident **out)
try { return frame_keep_impl(ctx, out, false); } catch (...) { return guard_exception(ctx); }
bq2_status bq2_frame_keep_own(bq2_ctx *ctx, bq2_resident **out)
try { return frame_keep_impl(ctx, out, true); } catch (...) { return guard_exception(ctx); }

static bq2_status resident_launch(bq2_ctx *ctx, const bq2_resident *r, cudaStream_t on = nullptr)
{
    FrameState &F = ctx->fs[r->state];
    cudaStream_t stream = on ? on : F.stream;
    if (!r->own && (r->L.lerped != F.lerped.p || r->L.extruded != F.extruded.p || r->L.indices != F.indices.p ||
        r->L.index_count != F.index_count.p || r->L.facing_bits != F.facing_bits.p))
        return fail(ctx, BQ2_E_INVALID, "bq2_resident_replay: the result arrays of that frame state were re-allocated");
    if ((!r->own && r->L.tri_normals != F.tri_normals.p) || r->L.meshes != ctx->d_meshes.p)
        return fail(ctx, BQ2_E_INVALID, "bq2_resident_replay: the mesh table or the scratch rows of that frame state were re-allocated");
    if (!r->own) F.tail_clean = false;                           /* writes counts into the state's buffer */
    bq2_launch L = r->L;
    L.seq = (uint32_t)ctx->launches;
    if (r->k_warp.n) {
        L.n_ent_slots = r->k_warp.n;
        CK((cudaError_t)bq2_launch_frame_warp(&L, r->k_warp.max_V, r->k_warp.max_T, r->k_warp.any_md2, stream, /*overlap=*/1), "launch(bq2_frame_warp_kernel)");
        ctx->launches++;
    }
    if (r->k_team.n) {
        L.ents = r->L.ents + r->k_warp.n; L.rec_base = (uint32_t)r->k_warp.n; L.n_ent_slots = r->k_team.n;
        CK((cudaError_t)bq2_launch_frame(&L, r->k_team.max_V, r->k_team.max_T, r->k_team.any_md2, stream, /*overlap=*/1), "launch(bq2_frame_kernel)");
        ctx->launches++;
    }
    return BQ2_OK;
}

bq2_status bq2_resident_replay(bq2_ctx *ctx, const bq2_resident *r)
try {
    if (!ctx || !r) return BQ2_E_INVALID;
    worker_drain(ctx);
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    return resident_launch(ctx, r);
} catch (...) { return guard_exception(ctx); }

/* k replays launched from C: ring[(start + i) % n_ring] for i = 0 .. k-1, all on the stream of the ring's frame state.
   ms != NULL: the launches are bracketed by two CUDA events on that stream and the call returns their distance after
   waiting for the second (the device-timed region of bench.py: no Python between the launches). */
bq2_status bq2_resident_run(bq2_ctx *ctx, bq2_resident *const *ring, int32_t n_ring, int32_t start, int32_t k, float *ms)
try {
    if (!ctx || !ring || n_ring <= 0 || k < 0 || start < 0) return BQ2_E_INVALID;
    worker_drain(ctx);
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    for (int i = 0; i < n_ring; i++) {
        if (!ring[i]) return fail(ctx, BQ2_E_INVALID, "bq2_resident_run: NULL frame in the ring");
        if (ring[i]->state != ring[0]->state) return fail(ctx, BQ2_E_INVALID, "bq2_resident_run: the frames of a ring must come from one frame state (one stream)");
    }
    cudaStream_t s = ctx->fs[ring[0]->state].stream;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms) {
        CK(cudaEventCreate(&e0), "cudaEventCreate"); 
        cudaError_t ce = cudaEventCreate(&e1);
        if (ce != cudaSuccess) { cudaEventDestroy(e0); return fail_cuda(ctx, ce, "cudaEventCreate"); }
        cudaEventRecord(e0, s);
    }
    bq2_status st = BQ2_OK;
    for (int i = 0; i < k && st == BQ2_OK; i++) st = resident_launch(ctx, ring[(start + i) % n_ring]);
    if (ms) {
        cudaError_t ce = cudaEventRecord(e1, s);
        if (ce == cudaSuccess) ce = cudaEventSynchronize(e1);
        if (ce == cudaSuccess) ce = cudaEventElapsedTime(ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (st == BQ2_OK && ce != cudaSuccess) return fail_cuda(ctx, ce, "bq2_resident_run(events)");
    }
    return st;
} catch (...) { return guard_exception(ctx); }

/* The same with the k launches dealt round-robin onto n_streams of the ctx's streams (frames in flight on different
   streams, as bq2_world_submit runs them): the frames are independent -- each kept frame owns its result arrays -- so
   nothing orders launch i + 1 behind launch i, and the tail of one overlaps the body of the next.  Needs frames kept
   with bq2_frame_keep_own (result arrays and team-kernel scratch rows of their own).  The events are
   on the first stream; the others start after the first event and are joined before the second. */
bq2_status bq2_resident_run_streams(bq2_ctx *ctx, bq2_resident *const *ring, int32_t n_ring, int32_t start, int32_t k,
                                    int32_t n_streams, float *ms)
try {
    if (!ctx || !ring || n_ring <= 0 || k < 0 || start < 0 || n_streams < 1 || n_streams > BQ2_FRAME_STATES) return BQ2_E_INVALID;
    worker_drain(ctx);
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    for (int i = 0; i < n_ring; i++) {
        if (!ring[i]) return fail(ctx, BQ2_E_INVALID, "bq2_resident_run_streams: NULL frame in the ring");
        if (n_streams > 1 && !ring[i]->own)
            return fail(ctx, BQ2_E_INVALID, "bq2_resident_run_streams: frames on several streams must own their result arrays (bq2_frame_keep_own)");
    }
    cudaStream_t s0 = ctx->fs[ring[0]->state].stream;
    std::vector<cudaStream_t> ss; ss.push_back(s0);
    for (int j = 0; j < BQ2_FRAME_STATES && (int)ss.size() < n_streams; j++) if (ctx->fs[j].stream != s0) ss.push_back(ctx->fs[j].stream);
    std::vector<cudaEvent_t> ev(ss.size() + 1, nullptr);
    for (auto &e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDefault), "cudaEventCreate");
    cudaEvent_t e0 = ev[0], e1 = ev[ss.size()];
    CK(cudaEventRecord(e0, s0), "cudaEventRecord");
    for (size_t j = 1; j < ss.size(); j++) CK(cudaStreamWaitEvent(ss[j], e0, 0), "cudaStreamWaitEvent");
    bq2_status st = BQ2_OK;
    for (int i = 0; i < k && st == BQ2_OK; i++) st = resident_launch(ctx, ring[(start + i) % n_ring], ss[(size_t)i % ss.size()]);
    for (size_t j = 1; j < ss.size(); j++) { cudaEventRecord(ev[j], ss[j]); cudaStreamWaitEvent(s0, ev[j], 0); }
    cudaError_t ce = cudaEventRecord(e1, s0);
    if (ce == cudaSuccess) ce = cudaEventSynchronize(e1);
    float t = 0.0f;
    if (ce == cudaSuccess) ce = cudaEventElapsedTime(&t, e0, e1);
    for (auto &e : ev) cudaEventDestroy(e);
    if (ms) *ms = t;
    if (st == BQ2_OK && ce != cudaSuccess) return fail_cuda(ctx, ce, "bq2_resident_run_streams(events)");
    return st;
} catch (...) { return guard_exception(ctx); }

uint64_t bq2_resident_result_bytes(const bq2_resident *r) { return r ? (uint64_t)r->result_bytes : 0; }

void bq2_resident_free(bq2_ctx *ctx, bq2_resident *r)
{
    if (!r) return;
    worker_drain(ctx);
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->fs[r->state].stream); }
    r->release();
    delete r;
}

/* One 64-bit hash per pair slot over everything the frame produced for it (bq2_hash_kernel); r == NULL: the frame
   bq2_frame_wait returned last.  Both kernel classes, host-built and device-built records. */
bq2_status bq2_hash_results(bq2_ctx *ctx, const bq2_resident *r, uint64_t *host_hashes, int32_t n_pair_slots)
try {
    if (!ctx || !host_hashes || n_pair_slots < 0) return BQ2_E_INVALID;
    worker_drain(ctx);
    if (!r && !ctx->submitted) return fail(ctx, BQ2_E_INVALID, "bq2_hash_results: nothing submitted");
    FrameState &F = ctx->fs[r ? r->state : ctx->last_waited];
    const int32_t nPS = F.n_pair_slots;       /* (a kept frame has the shape of the frame it was kept from) */
    if (!r && n_pair_slots != nPS) return fail(ctx, BQ2_E_INVALID, "bq2_hash_results: n_pair_slots does not match the frame");
    if (!(( r ? r->L.want : F.launch.want) & BQ2_WANT_SHADOW)) return fail(ctx, BQ2_E_INVALID, "bq2_hash_results: the frame has no shadow volumes");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    if (n_pair_slots == 0) return BQ2_OK;
    unsigned long long *d = nullptr;
    CK(cudaMalloc(&d, sizeof(unsigned long long) * (size_t)n_pair_slots), "cudaMalloc(hashes)");
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof(unsigned long long) * (size_t)n_pair_slots, F.stream);
    bq2_launch L = r ? r->L : F.launch;
    if (!r && F.world_mode) L.bucket_cnt = F.keep_cnt;          /* where the frame's last kernel left the pairs per record */
    const FrameState::KLaunch &kw = r ? r->k_warp : F.k_warp, &kt = r ? r->k_team : F.k_team;
    const bq2_dev_ent *ents0 = L.ents;
    if (e == cudaSuccess && kw.n) { L.n_ent_slots = kw.n; L.rec_base = 0; e = (cudaError_t)bq2_launch_hash(&L, d, F.stream); ctx->launches++; }
    if (e == cudaSuccess && kt.n) { L.ents = ents0 + kw.n; L.rec_base = (uint32_t)kw.n; L.n_ent_slots = kt.n; e = (cudaError_t)bq2_launch_hash(&L, d, F.stream); ctx->launches++; }
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_hashes, d, sizeof(unsigned long long) * (size_t)n_pair_slots, cudaMemcpyDeviceToHost, F.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(F.stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "bq2_hash_results");
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

/* cudaProfilerStart / Stop around a measured region (bench.py's DRAM-traffic probe runs under ncu's range replay) */
bq2_status bq2_profiler_range(bq2_ctx *ctx, int on)
{
    if (!ctx) return BQ2_E_INVALID;
    worker_drain(ctx);
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    CK(on ? cudaProfilerStart() : cudaProfilerStop(), "cudaProfilerStart/Stop");
    return BQ2_OK;
}

const uint32_t *bq2_host_index_counts(const bq2_ctx *ctx) { return ctx ? (const uint32_t *)ctx->fs[ctx->last_waited].hcnt.p : nullptr; }

bq2_status bq2_download(bq2_ctx *ctx, bq2_array which, uint64_t first, uint64_t count, void *host)
try {
    if (!ctx || !host) return BQ2_E_INVALID;
    worker_drain(ctx);
    FrameState &F = ctx->fs[ctx->last_waited];                   /* the frame bq2_frame_wait returned last */
    const void *src = nullptr; size_t cap = 0;
    switch (which) {
    case BQ2_A_LERPED:      src = F.lerped.p;      cap = F.lerped.cap; break;
    case BQ2_A_EXTRUDED:    src = F.extruded.p;    cap = F.extruded.cap; break;
    case BQ2_A_FACING_BITS: src = F.facing_bits.p; cap = F.facing_bits.cap; break;
    case BQ2_A_INDICES:     src = F.indices.p;     cap = F.indices.cap; break;
    case BQ2_A_INDEX_COUNT: src = F.index_count.p; cap = F.index_count.cap; break;
    case BQ2_A_NORMALS:     src = F.normals.p;     cap = F.normals.cap; break;
    case BQ2_A_TANGENTS:    src = F.tangents.p;    cap = F.tangents.cap; break;
    case BQ2_A_BINORMALS:   src = F.binormals.p;   cap = F.binormals.cap; break;
    case BQ2_A_LIGHTP:      src = F.lightp.p;      cap = F.lightp.cap; break;
    case BQ2_A_TSH:         src = F.tsh.p;         cap = F.tsh.cap; break;
    default: return fail(ctx, BQ2_E_INVALID, "bq2_download: unknown array");
    }
    if (!src || first + count > cap) return fail(ctx, BQ2_E_INVALID, "bq2_download: range outside the array");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaMemcpyAsync(host, (const uint32_t *)src + first, count * 4, cudaMemcpyDeviceToHost, F.stream), "cudaMemcpy(download)");
    CK(cudaStreamSynchronize(F.stream), "cudaStreamSynchronize(download)");
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

/* The first `words_per_slot` index words of pair slots [first_slot, first_slot + n_slots) of the frame bq2_frame_wait
   returned last, packed into host[n_slots][words_per_slot] by ONE strided copy (a consumer that draws from host memory
   brings a frame's index buffers back with this and bq2_download of the two vertex arrays). */
bq2_status bq2_download_indices(bq2_ctx *ctx, int32_t first_slot, int32_t n_slots, uint32_t words_per_slot, uint32_t *host)
try {
    if (!ctx || !host || first_slot < 0 || n_slots < 0) return BQ2_E_INVALID;
    worker_drain(ctx);
    FrameState &F = ctx->fs[ctx->last_waited];
    const uint32_t stride = F.launch.index_stride;
    if (!ctx->submitted || !F.indices.p || first_slot + n_slots > F.n_pair_slots || words_per_slot > stride)
        return fail(ctx, BQ2_E_INVALID, "bq2_download_indices: range outside the frame");
    if (n_slots == 0 || words_per_slot == 0) return BQ2_OK;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaMemcpy2DAsync(host, (size_t)words_per_slot * 4, F.indices.p + (size_t)first_slot * stride, (size_t)stride * 4,
                         (size_t)words_per_slot * 4, (size_t)n_slots, cudaMemcpyDeviceToHost, F.stream), "cudaMemcpy2D(indices)");
    CK(cudaStreamSynchronize(F.stream), "cudaStreamSynchronize(download)");
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

/* the same from the result arrays a kept frame writes to (its own, or those of its frame state) */
bq2_status bq2_resident_download(bq2_ctx *ctx, const bq2_resident *r, bq2_array which, uint64_t first, uint64_t count, void *host)
try {
    if (!ctx || !r || !host) return BQ2_E_INVALID;
    worker_drain(ctx);
    const FrameState &F = ctx->fs[r->state];
    const void *src = nullptr; size_t cap = 0;
    switch (which) {
    case BQ2_A_LERPED:      src = r->L.lerped;      cap = r->own ? r->lerped.cap : F.lerped.cap; break;
    case BQ2_A_EXTRUDED:    src = r->L.extruded;    cap = r->own ? r->extruded.cap : F.extruded.cap; break;
    case BQ2_A_FACING_BITS: src = r->L.facing_bits; cap = r->own ? r->facing_bits.cap : F.facing_bits.cap; break;
    case BQ2_A_INDICES:     src = r->L.indices;     cap = r->own ? r->indices.cap : F.indices.cap; break;
    case BQ2_A_INDEX_COUNT: src = r->L.index_count; cap = r->own ? r->index_count.cap : F.index_count.cap; break;
    case BQ2_A_NORMALS:     src = r->L.normals;     cap = r->own ? r->normals.cap : F.normals.cap; break;
    case BQ2_A_TANGENTS:    src = r->L.tangents;    cap = r->own ? r->tangents.cap : F.tangents.cap; break;
    case BQ2_A_BINORMALS:   src = r->L.binormals;   cap = r->own ? r->binormals.cap : F.binormals.cap; break;
    case BQ2_A_LIGHTP:      src = r->L.lightp;      cap = r->own ? r->lightp.cap : F.lightp.cap; break;
    case BQ2_A_TSH:         src = r->L.tsh;         cap = r->own ? r->tsh.cap : F.tsh.cap; break;
    default: return fail(ctx, BQ2_E_INVALID, "bq2_resident_download: unknown array");
    }
    if (!src || first + count > cap) return fail(ctx, BQ2_E_INVALID, "bq2_resident_download: range outside the array");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaMemcpyAsync(host, (const uint32_t *)src + first, count * 4, cudaMemcpyDeviceToHost, F.stream), "cudaMemcpy(download)");
    CK(cudaStreamSynchronize(F.stream), "cudaStreamSynchronize(download)");
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

bq2_status bq2_expand_trisverts(bq2_ctx *ctx, int32_t ps, float *tris_verts, int32_t max_verts, int32_t *tris_counter)
try {
    if (!ctx || !tris_verts || !tris_counter) return BQ2_E_INVALID;
    FrameState &F = ctx->fs[ctx->last_waited];
    if (!ctx->submitted || ps < 0 || ps >= F.n_pair_slots) return fail(ctx, BQ2_E_INVALID, "bq2_expand_trisverts: bad pair slot");
    if (F.world_mode && !F.world_tables)
        return fail(ctx, BQ2_E_INVALID, "bq2_expand_trisverts: the frame was submitted without BQ2_WANT_TABLES");
    const int V = F.slot_mesh_V[ps];
    const uint32_t pvoff = F.world_mode ? F.world_pvoff[ps] : F.pair_vert_off[ps];
    const uint32_t stride = F.launch.index_stride;
    uint32_t cnt = 0;
    bq2_status s = bq2_download(ctx, BQ2_A_INDEX_COUNT, (uint64_t)ps, 1, &cnt);
    if (s != BQ2_OK) return s;
    if (cnt > stride || (int32_t)cnt > max_verts) return fail(ctx, BQ2_E_OVERFLOW, "bq2_expand_trisverts: output too small");
    std::vector<float> pool(6 * (size_t)V);
    std::vector<uint32_t> idx(cnt);
    s = bq2_download(ctx, BQ2_A_LERPED, 3ull * F.ent_vert_off[F.pair_slot_entslot[ps]], 3ull * V, pool.data());
    if (s != BQ2_OK) return s;
    s = bq2_download(ctx, BQ2_A_EXTRUDED, 3ull * pvoff, 3ull * V, pool.data() + 3 * (size_t)V);
    if (s != BQ2_OK) return s;
    if (cnt) { s = bq2_download(ctx, BQ2_A_INDICES, (uint64_t)ps * stride, cnt, idx.data()); if (s != BQ2_OK) return s; }
    for (uint32_t k = 0; k < cnt; k++) memcpy(tris_verts + 3 * (size_t)k, &pool[3 * (size_t)idx[k]], 12);
    *tris_counter = (int32_t)cnt;
    return BQ2_OK;
} catch (...) { return guard_exception(ctx); }

} // extern "C"
