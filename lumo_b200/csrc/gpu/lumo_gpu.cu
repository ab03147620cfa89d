// liblumo_gpu.so — the C ABI of include/lumo_gpu.h over the sm_100a kernels of this directory.
// Host side only orchestrates: scene upload, wave allocation, the iteration loop (regen -> trace ->
// shade -> occlude) and the copies in and out.  No computation of the hot path happens on the CPU.
#include "lumo_gpu.h"
#include "bdpt.cuh"
#include <cub/device/device_scan.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace lumo_dev;

static thread_local std::string g_err;
static int32_t fail(int32_t code, const std::string& msg) { g_err = msg; return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(e_ == cudaErrorMemoryAllocation ? LUMO_ERR_OOM : LUMO_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

#define LUMO_ITER_LOG_CAP 16384
#define LUMO_ITER_BATCH 4   /* wave iterations enqueued per host synchronisation */
#define LUMO_TINY_SCENE_PRIMS 256u   /* at or below: closest hits and occlusion replay the reference traversal directly (measured on the 32-triangle Cornell box: 875 vs 714 Mrays/s) */

struct lumo_ctx {
    int device = 0, sm_count = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // wave storage, grown on demand and reused across renders
    void* wave_mem = nullptr; size_t wave_bytes = 0;
    void* bdpt_mem = nullptr; size_t bdpt_bytes = 0; void* bdpt_scan = nullptr; size_t bdpt_scan_bytes = 0;   // BDPT batch storage (vertex buffers), same policy
    unsigned long long* d_cursor = nullptr;            // work cursor of the ray-batch kernels
    uint8_t* blob_cache = nullptr; uint64_t blob_cache_bytes = 0;   // one recycled scene allocation (cudaFree / cudaMalloc stall unpredictably)
    void* film_mem = nullptr; size_t film_bytes = 0;   // device film of lumo_gpu_render (host-buffer entry point), reused across calls
    void* rgb_mem = nullptr; size_t rgb_bytes = 0;     // 8-bit image of lumo_gpu_film_encode*, reused across calls
    void* host_pinned = nullptr;   // IterCounters + RunCounters read-back
    unsigned long long launches = 0;
    Counters* d_visit = nullptr;   // traversal visit counters (CNT passes): [0] closest-hit kernels, [1] occlusion kernels
    int count_visits = 0;
    AhCounters* d_ah = nullptr;    // occlusion-BVH counters (CNT passes and LUMO_OCCLUDE_CHECK)
    ClosestCounters* d_ch = nullptr;   // closest-hit pipeline counters (CNT passes)
    int closest_faithful = 0;      // LUMO_CLOSEST_FAITHFUL=1: Scene::hit replays the reference traversal for every ray (k_wave_trace) instead of closest.cuh
    void* occl_mem = nullptr; size_t occl_bytes = 0;   // queues of the batch occlusion entry point (lumo_gpu_trace_any)
    int occl_faithful = 0;         // LUMO_OCCLUDE_FAITHFUL=1: shadow rays replay the reference's object BVH + kd-trees (k_wave_occlude) instead of the occlusion BVH
    uint64_t shade_stats[4] = {0, 0, 0, 0};   // of the last render: NEE bounces, NEE terms evaluated, shadow rays queued, bounces
    int occl_check = 0;            // LUMO_OCCLUDE_CHECK=1: run both on every shadow ray of a render and count disagreements (counters[9])
    uint32_t* d_iter_log = nullptr; uint32_t iter_log_n = 0;   // (closest-hit rays, shadow rays) per wave iteration of the last render's main pass
    // per-kernel-class device time of the last render (CUDA events on the launching stream)
    cudaEvent_t kev[5 * LUMO_ITER_BATCH] = {};
    double kernel_ms[4] = {0, 0, 0, 0}; unsigned long long kernel_launches[4] = {0, 0, 0, 0};
};
struct lumo_scene {
    lumo_ctx* ctx = nullptr;
    uint8_t* d_blob = nullptr; uint64_t len = 0, cap = 0;
    DevScene S;
    LumoBlobHeader H;
    bool has_textures = false;                       // any LumoTexture record (then materials may refer to textures / bump maps)
    uint32_t kind_mask = 0;   // bit k set: some Standard material of LumoMatKind k exists (which shade kernels to launch)
    bool tiny = false;        // a handful of primitives (Cornell box: 32): the reference traversal itself is a few steps, the BVH pipelines' extra passes cost more than they save
};

extern "C" const char* lumo_gpu_last_error(void) { return g_err.c_str(); }

extern "C" int32_t lumo_gpu_device_count(int32_t* n) {
    if (!n) return fail(LUMO_ERR_INVALID, "device_count: null pointer");
    int c = 0; CU(cudaGetDeviceCount(&c)); *n = c; return LUMO_OK;
}

extern "C" int32_t lumo_gpu_ctx_destroy(lumo_ctx* ctx);
static int32_t ctx_init(lumo_ctx* ctx, int32_t device, int sm_count) {
    ctx->device = device; ctx->sm_count = sm_count;
    CU(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    CU(cudaEventCreate(&ctx->ev0)); CU(cudaEventCreate(&ctx->ev1));
    CU(cudaMallocHost(&ctx->host_pinned, 4096));
    CU(cudaMalloc(&ctx->d_visit, 2 * sizeof(Counters)));
    CU(cudaMemset(ctx->d_visit, 0, 2 * sizeof(Counters)));
    for (auto& e : ctx->kev) CU(cudaEventCreate(&e));
    CU(cudaMalloc(&ctx->d_iter_log, LUMO_ITER_LOG_CAP * 8));
    CU(cudaMalloc(&ctx->d_cursor, 8));
    // the reference-order traversal keeps two explicit stacks per thread (2.3 KB).  The limit is device-wide and shared with the
    // host process (e.g. torch): only ever raised, never lowered.
    { size_t cur = 0; CU(cudaDeviceGetLimit(&cur, cudaLimitStackSize)); if (cur < 4096) CU(cudaDeviceSetLimit(cudaLimitStackSize, 4096)); }
    { const char* e = std::getenv("LUMO_OCCLUDE_FAITHFUL"); if (e) ctx->occl_faithful = std::atoi(e) != 0; }
    { const char* e = std::getenv("LUMO_OCCLUDE_CHECK"); if (e) ctx->occl_check = std::atoi(e) != 0; }
    CU(cudaMalloc(&ctx->d_ah, sizeof(AhCounters)));
    CU(cudaMemset(ctx->d_ah, 0, sizeof(AhCounters)));
    { const char* e = std::getenv("LUMO_CLOSEST_FAITHFUL"); if (e) ctx->closest_faithful = std::atoi(e) != 0; }
    CU(cudaMalloc(&ctx->d_ch, sizeof(ClosestCounters)));
    CU(cudaMemset(ctx->d_ch, 0, sizeof(ClosestCounters)));
    return LUMO_OK;
}
extern "C" int32_t lumo_gpu_ctx_create(int32_t device, lumo_ctx** out) {
    if (!out) return fail(LUMO_ERR_INVALID, "ctx_create: null pointer");
    int c = 0; CU(cudaGetDeviceCount(&c));
    if (device < 0 || device >= c) return fail(LUMO_ERR_INVALID, "ctx_create: no such device");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(LUMO_ERR_UNSUPPORTED, "ctx_create: liblumo_gpu is built for sm_100a only");
    lumo_ctx* ctx = new (std::nothrow) lumo_ctx();
    if (!ctx) return fail(LUMO_ERR_OOM, "ctx_create: out of host memory");
    const int32_t rc = ctx_init(ctx, device, prop.multiProcessorCount);
    if (rc != LUMO_OK) { const std::string msg = g_err; lumo_gpu_ctx_destroy(ctx); g_err = msg; return rc; }   // nothing created so far is leaked
    *out = ctx; return LUMO_OK;
}
extern "C" int32_t lumo_gpu_ctx_destroy(lumo_ctx* ctx) {
    if (!ctx) return LUMO_OK;
    cudaSetDevice(ctx->device);
    if (ctx->wave_mem) cudaFree(ctx->wave_mem);
    if (ctx->bdpt_mem) cudaFree(ctx->bdpt_mem);
    if (ctx->bdpt_scan) cudaFree(ctx->bdpt_scan);
    if (ctx->film_mem) cudaFree(ctx->film_mem);
    if (ctx->rgb_mem) cudaFree(ctx->rgb_mem);
    if (ctx->d_cursor) cudaFree(ctx->d_cursor);
    if (ctx->d_ah) cudaFree(ctx->d_ah);
    if (ctx->d_ch) cudaFree(ctx->d_ch);
    if (ctx->occl_mem) cudaFree(ctx->occl_mem);
    if (ctx->blob_cache) cudaFree(ctx->blob_cache);
    if (ctx->d_visit) cudaFree(ctx->d_visit);
    if (ctx->d_iter_log) cudaFree(ctx->d_iter_log);
    if (ctx->host_pinned) cudaFreeHost(ctx->host_pinned);
    for (auto& e : ctx->kev) if (e) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx; return LUMO_OK;
}
// Runs every later call of this context on the caller's CUDA stream (e.g. the stream the caller's NCCL
// reduce of the film is enqueued on).  NULL restores the context's own stream.
extern "C" int32_t lumo_gpu_ctx_set_stream(lumo_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(LUMO_ERR_INVALID, "set_stream: null ctx");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return LUMO_OK;
}
// Switches the traversal kernels of this context to their visit-counting instantiation (same
// traversal, plus per-thread counters of TLAS nodes / instance transforms / kd nodes / leaf entries /
// triangle tests / sphere tests — the N_* of the byte formula in DESIGN.md).  Not for timed runs.
extern "C" int32_t lumo_gpu_ctx_count_visits(lumo_ctx* ctx, int32_t enable) {
    if (!ctx) return fail(LUMO_ERR_INVALID, "count_visits: null ctx");
    CU(cudaSetDevice(ctx->device));
    ctx->count_visits = enable ? 1 : 0;
    CU(cudaMemset(ctx->d_visit, 0, 2 * sizeof(Counters)));
    CU(cudaMemset(ctx->d_ah, 0, sizeof(AhCounters)));
    CU(cudaMemset(ctx->d_ch, 0, sizeof(ClosestCounters)));
    return LUMO_OK;
}
// Closest hits (Scene::hit) go through the world-space BVH with the reference traversal run on the winning object only, and
// through the full reference traversal wherever that is not provably the same (csrc/gpu/closest.cuh).  mode 0: that (default);
// 1: the reference traversal for every ray.  stats (only while visit counting is on): [0] BVH nodes, [1] leaf primitives,
// [2] triangle tests, [3] sphere tests, [4] rays sent to the reference traversal, [5] rays, [6..13] why they were sent: stack
// overflow, another object at or below the nearest hit, a box above the winner failing, the winner's full hit rejected, its
// any-hit or full distance differing from the nearest hit, and the last three again for the nearest light.
extern "C" int32_t lumo_gpu_ctx_closest_mode(lumo_ctx* ctx, int32_t mode) {
    if (!ctx || mode < 0 || mode > 1) return fail(LUMO_ERR_INVALID, "closest_mode: bad arguments");
    ctx->closest_faithful = mode == 1;
    return LUMO_OK;
}
extern "C" int32_t lumo_gpu_ctx_closest_stats(lumo_ctx* ctx, uint64_t* out6) {   // fourteen values
    if (!ctx || !out6) return fail(LUMO_ERR_INVALID, "closest_stats: null pointer");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ClosestCounters c; CU(cudaMemcpy(&c, ctx->d_ch, sizeof c, cudaMemcpyDeviceToHost));
    out6[0] = c.nodes; out6[1] = c.prims; out6[2] = c.tris; out6[3] = c.spheres; out6[4] = c.fallback; out6[5] = c.rays; for (int k = 0; k < 8; k++) out6[6 + k] = c.why[k];
    return LUMO_OK;
}
// Work of the shading stage in the context's last PathTrace / DirectLight render: [0] bounces that ran next-event estimation,
// [1] NEE terms that reached k_nee_eval (light- and BSDF-sampled), [2] shadow rays queued, [3] bounces shaded.
extern "C" int32_t lumo_gpu_ctx_shade_stats(lumo_ctx* ctx, uint64_t* out4) {
    if (!ctx || !out4) return fail(LUMO_ERR_INVALID, "shade_stats: null pointer");
    for (int k = 0; k < 4; k++) out4[k] = ctx->shade_stats[k];
    return LUMO_OK;
}
// Counters of the occlusion-BVH kernels (occlude.cuh) since the last lumo_gpu_ctx_count_visits call: [0] BVH nodes visited,
// [1] leaf primitives fetched, [2] f64 triangle tests, [3] sphere tests, [4] candidate blockers sent to the confirmation
// pass, [5] confirmed by the reference's per-object traversal, [6] rays sent to the faithful kernel, [8] blockers accepted
// as robust without confirmation — all only while visit counting is on — and [7] rays on which the occlusion BVH and the
// faithful kernel disagreed (LUMO_OCCLUDE_CHECK=1 renders; must stay 0).
extern "C" int32_t lumo_gpu_ctx_occlusion_stats(lumo_ctx* ctx, uint64_t* out8) {   // nine values
    if (!ctx || !out8) return fail(LUMO_ERR_INVALID, "occlusion_stats: null pointer");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    AhCounters c; CU(cudaMemcpy(&c, ctx->d_ah, sizeof c, cudaMemcpyDeviceToHost));
    out8[0] = c.nodes; out8[1] = c.prims; out8[2] = c.tris; out8[3] = c.spheres; out8[4] = c.candidates; out8[5] = c.confirmed; out8[6] = c.fallback; out8[7] = c.mismatches; out8[8] = c.robust;
    return LUMO_OK;
}
// Overrides the LUMO_OCCLUDE_FAITHFUL / LUMO_OCCLUDE_CHECK environment switches of this context: mode 0 = occlusion BVH
// (default), 1 = the reference's traversal for shadow rays too, 2 = both on every shadow ray, disagreements counted.
extern "C" int32_t lumo_gpu_ctx_occlusion_mode(lumo_ctx* ctx, int32_t mode) {
    if (!ctx || mode < 0 || mode > 2) return fail(LUMO_ERR_INVALID, "occlusion_mode: bad arguments");
    ctx->occl_faithful = mode == 1; ctx->occl_check = mode == 2;
    return LUMO_OK;
}
// out12: six counters of the closest-hit kernels, then six of the occlusion kernels
extern "C" int32_t lumo_gpu_ctx_visits(lumo_ctx* ctx, uint64_t* out12) {
    if (!ctx || !out12) return fail(LUMO_ERR_INVALID, "visits: null pointer");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    Counters c[2]; CU(cudaMemcpy(c, ctx->d_visit, sizeof c, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 2; k++) { uint64_t* o = out12 + 6 * k; o[0] = c[k].tlas; o[1] = c[k].inst; o[2] = c[k].kd; o[3] = c[k].leaf; o[4] = c[k].tri; o[5] = c[k].sphere; }
    return LUMO_OK;
}
// Queue sizes of the last render's main pass: out[2i] = rays traced by k_wave_trace in iteration i, out[2i+1] = shadow
// rays traced by k_wave_occlude.  Returns the number of iterations in *n (at most cap are written).
extern "C" int32_t lumo_gpu_ctx_iter_log(lumo_ctx* ctx, uint32_t* out, uint32_t cap, uint32_t* n) {
    if (!ctx || !out || !n) return fail(LUMO_ERR_INVALID, "iter_log: null pointer");
    CU(cudaSetDevice(ctx->device));
    const uint32_t k = ctx->iter_log_n < cap ? ctx->iter_log_n : cap;
    if (k) CU(cudaMemcpy(out, ctx->d_iter_log, (size_t)k * 8, cudaMemcpyDeviceToHost));
    *n = ctx->iter_log_n;
    return LUMO_OK;
}
// Device time (ms) and launch count per kernel class of the last render: regen, trace, shade, occlude.
extern "C" int32_t lumo_gpu_ctx_kernel_times(lumo_ctx* ctx, double* ms4, uint64_t* launches4) {
    if (!ctx || !ms4 || !launches4) return fail(LUMO_ERR_INVALID, "kernel_times: null pointer");
    for (int k = 0; k < 4; k++) { ms4[k] = ctx->kernel_ms[k]; launches4[k] = ctx->kernel_launches[k]; }
    return LUMO_OK;
}

// lumo_math.h as compiled for the device: fn 0 sin, 1 cos, 2 atan2(x, y), 3 acos, 4 atanh, 5 cosh, 6 exp, 7 log, 8 pow(x, y).
// tests/test_lumo_math.py compares the bits with the same header compiled by g++ for the oracle and for the host library.
__global__ void k_math_eval(int fn, const double* x, const double* y, uint64_t n, double* out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double a = x[i], b = y ? y[i] : 0.0;
        double r;
        switch (fn) {
        case 0: r = lm_sin(a); break; case 1: r = lm_cos(a); break; case 2: r = lm_atan2(a, b); break; case 3: r = lm_acos(a); break;
        case 4: r = lm_atanh(a); break; case 5: r = lm_cosh(a); break; case 6: r = lm_exp(a); break; case 7: r = lm_log(a); break;
        default: r = lm_pow(a, b); break;
        }
        out[i] = r;
    }
}
extern "C" int32_t lumo_gpu_math_eval(lumo_ctx* ctx, int32_t fn, const double* x, const double* y, uint64_t n, double* out) {
    if (!ctx || !x || !out) return fail(LUMO_ERR_INVALID, "math_eval: null pointer");
    if (fn < 0 || fn > 8) return fail(LUMO_ERR_INVALID, "math_eval: unknown function");
    if (n == 0) return LUMO_OK;
    CU(cudaSetDevice(ctx->device));
    double *dx = nullptr, *dy = nullptr, *dout = nullptr;
    cudaError_t e = cudaMalloc(&dx, n * 8);
    if (e == cudaSuccess && y) e = cudaMalloc(&dy, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&dout, n * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dx, x, n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && y) e = cudaMemcpyAsync(dy, y, n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) { k_math_eval<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 4096), 256, 0, ctx->stream>>>(fn, dx, dy, n, dout); e = cudaGetLastError(); ctx->launches++; }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (dx) cudaFree(dx); if (dy) cudaFree(dy); if (dout) cudaFree(dout);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? LUMO_ERR_OOM : LUMO_ERR_CUDA, std::string("math_eval: ") + cudaGetErrorString(e));
    return LUMO_OK;
}

// FP64 issue ceiling of this GPU (SURVEY 8d asks for it next to the bandwidth roofline: every kernel of the path computes
// in f64): eight independent DFMA chains per thread, 2048 resident threads per SM.  Returns TFLOP/s (2 flops per DFMA).
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters) {
    double a0 = 1e-9 * threadIdx.x, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double b = 0.99999999, c = 1e-9;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        a0 = __fma_rn(a0, b, c); a1 = __fma_rn(a1, b, c); a2 = __fma_rn(a2, b, c); a3 = __fma_rn(a3, b, c);
        a4 = __fma_rn(a4, b, c); a5 = __fma_rn(a5, b, c); a6 = __fma_rn(a6, b, c); a7 = __fma_rn(a7, b, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == -1.0) out[0] = s;   // never true; keeps the chains alive
}
extern "C" int32_t lumo_gpu_fp64_peak(lumo_ctx* ctx, double* tflops, double* ms_out) {
    if (!ctx || !tflops) return fail(LUMO_ERR_INVALID, "fp64_peak: null pointer");
    CU(cudaSetDevice(ctx->device));
    const int iters = 8192, grid = ctx->sm_count * 8;
    double best = 0.0, best_ms = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
        k_fp64_peak<<<grid, 256, 0, ctx->stream>>>((double*)ctx->d_cursor, iters);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float ms = 0; CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double tf = 2.0 * 8.0 * (double)iters * (double)grid * 256.0 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) { best = tf; best_ms = ms; }
        ctx->launches++;
    }
    *tflops = best; if (ms_out) *ms_out = best_ms;
    return LUMO_OK;
}

// ---- scene -------------------------------------------------------------------------------------------
static bool validate_blob(const uint8_t* b, uint64_t len, LumoBlobHeader& H, std::string& why) {
    if (len < sizeof(LumoBlobHeader)) { why = "blob shorter than its header"; return false; }
    std::memcpy(&H, b, sizeof H);
    if (H.magic != LUMO_BLOB_MAGIC) { why = "bad blob magic"; return false; }
    if (H.version != LUMO_BLOB_VERSION) { why = "unsupported blob version"; return false; }
    if (H.n_sections != LSEC_COUNT || H.total_bytes != len) { why = "blob size / section count mismatch"; return false; }
    static const size_t elem[LSEC_COUNT] = {sizeof(LumoTlasNode), 4, sizeof(LumoObject), sizeof(LumoInstance), sizeof(LumoKdTree), sizeof(LumoKdNode), 4,
                                            sizeof(LumoTriVerts), sizeof(LumoTriShade), 24, 16, sizeof(LumoRect), sizeof(LumoSphere), sizeof(LumoMaterial), 96 * 8, sizeof(LumoLight),
                                            sizeof(LumoTexture), 16, 8, sizeof(LumoAhNode), sizeof(LumoAhPrim), 4, 4};
    for (int s = 0; s < LSEC_COUNT; s++) {
        const LumoSectionRef& r = H.sec[s];
        if (r.offset % 16 || r.offset > len || r.bytes > len - r.offset || r.bytes != r.count * elem[s]) { why = "blob section " + std::to_string(s) + " is malformed"; return false; }
    }
    const LumoSceneParams& P = H.params;
    if (P.n_objects == 0 || P.n_lights == 0) { why = "scene needs at least one object and one light"; return false; }
    if (P.n_lights >= (1u << 24) || P.n_shadow_rays == 0 || P.n_shadow_rays > 64) { why = "more than 2^24 lights, or a shadow-ray count outside [1, 64]"; return false; }
    if (H.sec[LSEC_OBJECTS].count != (uint64_t)P.n_objects + P.n_lights || H.sec[LSEC_LIGHTS].count != P.n_lights) { why = "object / light counts disagree with sections"; return false; }
    if (P.camera.res_x == 0 || P.camera.res_y == 0) { why = "camera resolution is zero"; return false; }
    // index ranges (a malformed blob must not make a kernel read out of bounds)
    const LumoObject* objs = (const LumoObject*)(b + H.sec[LSEC_OBJECTS].offset);
    for (uint64_t i = 0; i < H.sec[LSEC_OBJECTS].count; i++) {
        const LumoObject& o = objs[i];
        const uint64_t lim = (o.kind == LOBJ_KD || o.kind == LOBJ_RECT) ? H.sec[LSEC_KD_TREES].count : (o.kind == LOBJ_SPHERE ? H.sec[LSEC_SPHERES].count : H.sec[LSEC_TRI_VERTS].count);
        if (o.kind > LOBJ_TRI || o.geom >= lim || o.inst >= (int64_t)H.sec[LSEC_INSTANCES].count || o.material < 0 || (uint64_t)o.material >= H.sec[LSEC_MATERIALS].count ||
            (o.kind == LOBJ_RECT && o.rect >= H.sec[LSEC_RECTS].count)) { why = "object record " + std::to_string(i) + " out of range"; return false; }
    }
    const LumoKdTree* kds = (const LumoKdTree*)(b + H.sec[LSEC_KD_TREES].offset);
    for (uint64_t i = P.n_objects; i < H.sec[LSEC_OBJECTS].count; i++)   // lights are intersected with a small kd stack (LUMO_LIGHT_KD_STACK)
        if ((objs[i].kind == LOBJ_KD || objs[i].kind == LOBJ_RECT) && kds[objs[i].geom].n_tris > LUMO_LIGHT_KD_STACK) { why = "light kd-tree too large (lights are Rectangle / Triangle / Sphere objects)"; return false; }
    for (uint64_t i = 0; i < H.sec[LSEC_KD_TREES].count; i++)
        if (kds[i].root >= H.sec[LSEC_KD_NODES].count || (uint64_t)kds[i].tri_base + kds[i].n_tris > H.sec[LSEC_TRI_VERTS].count) { why = "kd tree record out of range"; return false; }
    const LumoKdNode* kn = (const LumoKdNode*)(b + H.sec[LSEC_KD_NODES].offset);
    for (uint64_t i = 0; i < H.sec[LSEC_KD_NODES].count; i++) {
        if (kn[i].b & 0x80000000u) { if ((uint64_t)kn[i].a + (kn[i].b & 0x7FFFFFFFu) > H.sec[LSEC_KD_LEAF].count) { why = "kd leaf range out of bounds"; return false; } }
        else if (kn[i].b > 2 || kn[i].a >= H.sec[LSEC_KD_NODES].count || i + 1 >= H.sec[LSEC_KD_NODES].count) { why = "kd inner node out of bounds"; return false; }
    }
    const LumoTlasNode* tn = (const LumoTlasNode*)(b + H.sec[LSEC_TLAS_NODES].offset);
    if (P.lights_root >= H.sec[LSEC_TLAS_NODES].count) { why = "lights_root out of range"; return false; }
    for (uint64_t i = 0; i < H.sec[LSEC_TLAS_NODES].count; i++)
        if ((uint64_t)tn[i].first + tn[i].count > H.sec[LSEC_TLAS_LEAF].count) { why = "TLAS leaf range out of bounds"; return false; }
    const LumoMaterial* mats = (const LumoMaterial*)(b + H.sec[LSEC_MATERIALS].offset);
    for (uint64_t i = 0; i < H.sec[LSEC_MATERIALS].count; i++)
        if (mats[i].eta_table >= H.sec[LSEC_TABLES].count || mats[i].k_table >= H.sec[LSEC_TABLES].count || mats[i].illum_table >= H.sec[LSEC_TABLES].count) { why = "material table index out of range"; return false; }
    const LumoTexture* tx = (const LumoTexture*)(b + H.sec[LSEC_TEXTURES].offset);
    const uint64_t n_tex = H.sec[LSEC_TEXTURES].count;
    for (uint64_t i = 0; i < n_tex; i++) {
        const LumoTexture& T = tx[i];
        bool ok = T.kind <= LTEX_BUMP;
        if (T.kind == LTEX_CHECKER) ok = T.a < i && T.b < i && tx[T.a].kind != LTEX_BUMP && tx[T.b].kind != LTEX_BUMP;   // children precede: no cycles
        if (T.kind == LTEX_MARBLE) ok = T.data + 1536 <= H.sec[LSEC_TEX_F64].count;
        if (T.kind == LTEX_IMAGE) ok = T.width > 0 && T.height > 0 && T.data + (uint64_t)T.width * T.height <= H.sec[LSEC_TEX_PIXELS].count;
        if (T.kind == LTEX_BUMP) ok = T.width > 0 && T.height > 0 && T.data + 3ull * T.width * T.height <= H.sec[LSEC_TEX_F64].count;
        if (!ok) { why = "texture record " + std::to_string(i) + " out of range"; return false; }
    }
    {   // occlusion BVH: child links, leaf ranges, primitive records; per-object node chains
        const LumoAhNode* an = (const LumoAhNode*)(b + H.sec[LSEC_AH_NODES].offset);
        const LumoAhPrim* ap = (const LumoAhPrim*)(b + H.sec[LSEC_AH_PRIMS].offset);
        const uint64_t n_an = H.sec[LSEC_AH_NODES].count, n_ap = H.sec[LSEC_AH_PRIMS].count, n_obj = H.sec[LSEC_OBJECTS].count;
        if (n_an == 0 || n_ap == 0 || n_ap >= (1ull << 27)) { why = "occlusion BVH is empty or too large"; return false; }
        for (uint64_t i = 0; i < n_an; i++) for (int k = 0; k < 4; k++) {
            const uint32_t c = an[i].child[k];
            if (c == LUMO_NONE) { if (!(an[i].lo_x[k] > an[i].hi_x[k])) { why = "occlusion BVH: empty child with a non-empty box"; return false; } continue; }
            if (c & LUMO_AH_LEAF) { const uint64_t first = c & 0x07FFFFFFu, cnt = ((c >> 27) & 0xFu) + 1u; if (first + cnt > n_ap) { why = "occlusion BVH leaf range out of bounds"; return false; } }
            else if (c >= n_an || c <= i) { why = "occlusion BVH child link out of range"; return false; }   // children follow their parent: no cycles
        }
        for (uint64_t i = 0; i < n_ap; i++) {
            const uint32_t o = ap[i].obj & ~LUMO_AH_INSTANCED;
            if (o >= n_obj || ((ap[i].obj & LUMO_AH_INSTANCED) != 0) != (objs[o].inst >= 0)) { why = "occlusion BVH primitive: bad object"; return false; }
            if (ap[i].tri & LUMO_AH_SPHERE) { if ((ap[i].tri & ~LUMO_AH_SPHERE) >= H.sec[LSEC_SPHERES].count) { why = "occlusion BVH primitive: bad sphere"; return false; } }
            else if (ap[i].tri >= H.sec[LSEC_TRI_VERTS].count) { why = "occlusion BVH primitive: bad triangle"; return false; }
        }
        const uint32_t* po = (const uint32_t*)(b + H.sec[LSEC_OBJ_PATH_OFF].offset); const uint32_t* pn = (const uint32_t*)(b + H.sec[LSEC_OBJ_PATH].offset);
        if (H.sec[LSEC_OBJ_PATH_OFF].count != n_obj + 1) { why = "object path offsets: wrong count"; return false; }
        for (uint64_t i = 0; i < n_obj; i++) if (po[i] >= po[i + 1] || po[i + 1] > H.sec[LSEC_OBJ_PATH].count) { why = "object path offsets out of range (every object sits under at least one BVH node)"; return false; }
        for (uint64_t i = 0; i < H.sec[LSEC_OBJ_PATH].count; i++) if (pn[i] >= H.sec[LSEC_TLAS_NODES].count) { why = "object path node out of range"; return false; }
    }
    for (uint64_t i = 0; i < H.sec[LSEC_MATERIALS].count; i++) {
        const uint32_t col[4] = {mats[i].kd_tex, mats[i].ks_tex, mats[i].tf_tex, mats[i].ke_tex};
        for (uint32_t t : col) if (t != LUMO_NONE && (t >= n_tex || tx[t].kind == LTEX_BUMP)) { why = "material texture index out of range"; return false; }
        if (mats[i].bump_tex != LUMO_NONE && (mats[i].bump_tex >= n_tex || tx[mats[i].bump_tex].kind != LTEX_BUMP)) { why = "material bump map index out of range"; return false; }
    }
    return true;
}

extern "C" int32_t lumo_gpu_scene_upload(lumo_ctx* ctx, const void* blob, uint64_t len, lumo_scene** out) {
    if (!ctx || !blob || !out) return fail(LUMO_ERR_INVALID, "scene_upload: null pointer");
    LumoBlobHeader H; std::string why;
    if (!validate_blob((const uint8_t*)blob, len, H, why)) return fail(LUMO_ERR_INVALID, "scene_upload: " + why);
    CU(cudaSetDevice(ctx->device));
    lumo_scene* sc = new (std::nothrow) lumo_scene();
    if (!sc) return fail(LUMO_ERR_OOM, "scene_upload: out of host memory");
    sc->ctx = ctx; sc->len = len; sc->H = H;
    cudaError_t e = cudaSuccess;
    if (ctx->blob_cache && ctx->blob_cache_bytes >= len) { sc->d_blob = ctx->blob_cache; sc->cap = ctx->blob_cache_bytes; ctx->blob_cache = nullptr; ctx->blob_cache_bytes = 0; }
    else { e = cudaMalloc((void**)&sc->d_blob, len); sc->cap = len; }
    if (e != cudaSuccess) { delete sc; return fail(LUMO_ERR_OOM, std::string("scene_upload: cudaMalloc: ") + cudaGetErrorString(e)); }
    e = cudaMemcpyAsync(sc->d_blob, blob, len, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(sc->d_blob); delete sc; return fail(LUMO_ERR_CUDA, std::string("scene_upload: copy: ") + cudaGetErrorString(e)); }
    DevScene& S = sc->S;
    auto at = [&](int s) { return sc->d_blob + H.sec[s].offset; };
    S.tlas = (const LumoTlasNode*)at(LSEC_TLAS_NODES); S.tlas_leaf = (const uint32_t*)at(LSEC_TLAS_LEAF);
    S.objects = (const LumoObject*)at(LSEC_OBJECTS); S.instances = (const LumoInstance*)at(LSEC_INSTANCES);
    S.kd_trees = (const LumoKdTree*)at(LSEC_KD_TREES); S.kd_nodes = (const LumoKdNode*)at(LSEC_KD_NODES); S.kd_leaf = (const uint32_t*)at(LSEC_KD_LEAF);
    S.tri_verts = (const LumoTriVerts*)at(LSEC_TRI_VERTS); S.tri_shade = (const LumoTriShade*)at(LSEC_TRI_SHADE);
    S.normals = (const double*)at(LSEC_NORMALS); S.uvs = (const double*)at(LSEC_UVS);
    S.rects = (const LumoRect*)at(LSEC_RECTS); S.spheres = (const LumoSphere*)at(LSEC_SPHERES);
    S.materials = (const LumoMaterial*)at(LSEC_MATERIALS); S.tables = (const double*)at(LSEC_TABLES); S.lights = (const LumoLight*)at(LSEC_LIGHTS);
    sc->has_textures = H.sec[LSEC_TEXTURES].count > 0;
    sc->tiny = H.sec[LSEC_AH_PRIMS].count <= LUMO_TINY_SCENE_PRIMS;
    S.textures = (const LumoTexture*)at(LSEC_TEXTURES); S.tex_pixels = (const float*)at(LSEC_TEX_PIXELS); S.tex_f64 = (const double*)at(LSEC_TEX_F64);
    S.ah_nodes = (const LumoAhNode*)at(LSEC_AH_NODES); S.ah_prims = (const LumoAhPrim*)at(LSEC_AH_PRIMS);
    S.obj_path_off = (const uint32_t*)at(LSEC_OBJ_PATH_OFF); S.obj_path = (const uint32_t*)at(LSEC_OBJ_PATH);
    S.P = H.params;
    { const LumoMaterial* mats = (const LumoMaterial*)((const uint8_t*)blob + H.sec[LSEC_MATERIALS].offset);
      for (uint64_t i = 0; i < H.sec[LSEC_MATERIALS].count; i++) if (mats[i].kind < 32) sc->kind_mask |= 1u << mats[i].kind; }
    *out = sc; return LUMO_OK;
}
extern "C" int32_t lumo_gpu_scene_destroy(lumo_scene* sc) {
    if (!sc) return LUMO_OK;
    lumo_ctx* ctx = sc->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (sc->cap > ctx->blob_cache_bytes) { if (ctx->blob_cache) cudaFree(ctx->blob_cache); ctx->blob_cache = sc->d_blob; ctx->blob_cache_bytes = sc->cap; }
    else cudaFree(sc->d_blob);
    delete sc; return LUMO_OK;
}

// ---- ray batches -----------------------------------------------------------------------------------
static int trace_grid(const lumo_ctx* ctx) { return ctx->sm_count * 8; }

struct Carver {   // carves 256-byte aligned arrays out of one allocation
    uint8_t* base; size_t off = 0;
    template <class T> T* take(size_t n) { T* p = base ? (T*)(base + off) : nullptr; off = (off + n * sizeof(T) + 255) & ~(size_t)255; return p; }
};
// The occlusion half of Scene::hit_light through the occlusion BVH (occlude.cuh): BVH walk -> per-object confirmation by the
// reference's own traversal -> faithful kernel for whatever is left.  Q.counters must be zero.
template <class Source, class Sink>
static int32_t launch_occlusion(lumo_scene* sc, const Source& src, const Sink& sink, const OcclQueues& Q, cudaStream_t st) {
    lumo_ctx* ctx = sc->ctx;
    const int g1 = ctx->sm_count * LUMO_BVH_BLOCKS, g2 = ctx->sm_count * LUMO_WAVE_TRACE_BLOCKS;
    if (ctx->count_visits) {
        k_occl_bvh<true><<<g1, 128, 0, st>>>(sc->S, src, sink, Q, ctx->d_ah);
        k_occl_confirm<true><<<g2, 128, 0, st>>>(sc->S, src, sink, Q, ctx->d_visit + 1, ctx->d_ah);
        k_occl_fallback<true><<<g2, 128, 0, st>>>(sc->S, src, sink, Q, ctx->d_visit + 1);
    } else {
        k_occl_bvh<false><<<g1, 128, 0, st>>>(sc->S, src, sink, Q, nullptr);
        k_occl_confirm<false><<<g2, 128, 0, st>>>(sc->S, src, sink, Q, nullptr, nullptr);
        k_occl_fallback<false><<<g2, 128, 0, st>>>(sc->S, src, sink, Q, nullptr);
    }
    ctx->launches += 3;
    CU(cudaGetLastError());
    return LUMO_OK;
}

// Scene::hit through closest.cuh: BVH walk -> finish on the winning object -> reference traversal for the rest.  Q.counters must be zero.
template <class Source, class Sink>
static int32_t launch_closest(lumo_scene* sc, const Source& src, const Sink& sink, const ClosestScratch& Q, cudaStream_t st) {
    lumo_ctx* ctx = sc->ctx;
    const int g1 = ctx->sm_count * LUMO_BVH_BLOCKS, g2 = ctx->sm_count * LUMO_WAVE_TRACE_BLOCKS;
    if (ctx->count_visits) {
        k_closest_bvh<true><<<g1, 128, 0, st>>>(sc->S, src, Q, ctx->d_ch);
        k_closest_finish<true><<<g2, 128, 0, st>>>(sc->S, src, sink, Q, ctx->d_visit, ctx->d_ch);
        k_closest_fallback<true><<<g2, 128, 0, st>>>(sc->S, src, sink, Q, ctx->d_visit, ctx->d_ch);
    } else {
        k_closest_bvh<false><<<g1, 128, 0, st>>>(sc->S, src, Q, nullptr);
        k_closest_finish<false><<<g2, 128, 0, st>>>(sc->S, src, sink, Q, nullptr, nullptr);
        k_closest_fallback<false><<<g2, 128, 0, st>>>(sc->S, src, sink, Q, nullptr, nullptr);
    }
    ctx->launches += 3;
    CU(cudaGetLastError());
    return LUMO_OK;
}

template <int MODE>
static int32_t launch_batch(lumo_scene* sc, const double* o_dev, const double* d_dev, const double* tmax_dev, uint64_t n, unsigned long long* next_dev,
                            uint32_t* obj, uint32_t* tri, double* t, double* bary, uint8_t* occ) {
    lumo_ctx* ctx = sc->ctx;
    if (MODE == 1 && !ctx->occl_faithful) {
        // queues: confirm (index, object) + fallback (index) per ray of a chunk, 8 counters
        const uint64_t CH = 1ull << 26;
        const size_t need = (size_t)std::min<uint64_t>(n, CH) * 12 + 256;
        if (need > ctx->occl_bytes) {
            if (ctx->occl_mem) { cudaFree(ctx->occl_mem); ctx->occl_mem = nullptr; ctx->occl_bytes = 0; }
            CU(cudaMalloc(&ctx->occl_mem, need)); ctx->occl_bytes = need;
        }
        for (uint64_t c0 = 0; c0 < n; c0 += CH) {
            const uint32_t cnt = (uint32_t)std::min<uint64_t>(CH, n - c0);
            OcclQueues Q; Q.counters = (uint32_t*)ctx->occl_mem; Q.confirm_i = Q.counters + 64; Q.confirm_obj = Q.confirm_i + cnt; Q.fallback_i = Q.confirm_obj + cnt;
            CU(cudaMemsetAsync(Q.counters, 0, 32, ctx->stream));
            BatchShadowSource src{o_dev + 3 * c0, d_dev + 3 * c0, tmax_dev + c0, cnt}; BatchShadowSink sink{occ + c0};
            int32_t rc = launch_occlusion(sc, src, sink, Q, ctx->stream); if (rc != LUMO_OK) return rc;
        }
        return LUMO_OK;
    }
    if (MODE == 0 && !ctx->closest_faithful) {
        // per ray of a chunk: t1, tl (8 B each), o1, ol, flags, fallback index (4 B each); 4 counters
        const uint64_t CH = 1ull << 26;
        const size_t need = (size_t)std::min<uint64_t>(n, CH) * 32 + 256;
        if (need > ctx->occl_bytes) {
            if (ctx->occl_mem) { cudaFree(ctx->occl_mem); ctx->occl_mem = nullptr; ctx->occl_bytes = 0; }
            CU(cudaMalloc(&ctx->occl_mem, need)); ctx->occl_bytes = need;
        }
        for (uint64_t c0 = 0; c0 < n; c0 += CH) {
            const uint32_t cnt = (uint32_t)std::min<uint64_t>(CH, n - c0);
            ClosestScratch Q; Q.counters = (uint32_t*)ctx->occl_mem; Q.t1 = (double*)((uint8_t*)ctx->occl_mem + 256); Q.tl = Q.t1 + cnt;
            Q.o1 = (uint32_t*)(Q.tl + cnt); Q.ol = Q.o1 + cnt; Q.flags = Q.ol + cnt; Q.fallback_i = Q.flags + cnt;
            CU(cudaMemsetAsync(Q.counters, 0, 16, ctx->stream));
            BatchRaySource src{o_dev + 3 * c0, d_dev + 3 * c0, tmax_dev ? tmax_dev + c0 : nullptr, cnt};
            BatchHitSink sink{obj + c0, tri + c0, t + c0, bary + 2 * c0};
            int32_t rc = launch_closest(sc, src, sink, Q, ctx->stream); if (rc != LUMO_OK) return rc;
        }
        return LUMO_OK;
    }
    CU(cudaMemsetAsync(next_dev, 0, 8, ctx->stream));
    if (ctx->count_visits) k_trace_batch<MODE, true><<<trace_grid(ctx), 128, 0, ctx->stream>>>(sc->S, o_dev, d_dev, tmax_dev, n, next_dev, obj, tri, t, bary, occ, ctx->d_visit + (MODE == 0 ? 0 : 1));
    else k_trace_batch<MODE, false><<<trace_grid(ctx), 128, 0, ctx->stream>>>(sc->S, o_dev, d_dev, tmax_dev, n, next_dev, obj, tri, t, bary, occ, nullptr);
    ctx->launches++;
    CU(cudaGetLastError());
    return LUMO_OK;
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 16); }
    template <class T> T* as() { return (T*)p; }
};

template <int MODE>
static int32_t trace_host(lumo_scene* sc, const double* o, const double* d, const double* t_max, uint64_t n, uint32_t* obj, uint32_t* tri, double* t, double* bary, uint8_t* occ) {
    if (!sc || !o || !d) return fail(LUMO_ERR_INVALID, "trace: null pointer");
    if (MODE == 0 && (!obj || !tri || !t || !bary)) return fail(LUMO_ERR_INVALID, "trace_closest: null output");
    if (MODE == 1 && (!occ || !t_max)) return fail(LUMO_ERR_INVALID, "trace_any: null t_max / output");
    if (MODE == 2 && !t) return fail(LUMO_ERR_INVALID, "trace_first_found: null output");
    if (n == 0) return LUMO_OK;
    lumo_ctx* ctx = sc->ctx;
    CU(cudaSetDevice(ctx->device));
    DevBuf bo, bd, bt, bobj, btri, btt, bbary, bocc;
    CU(bo.alloc(n * 24)); CU(bd.alloc(n * 24));
    if (t_max) CU(bt.alloc(n * 8));
    if (MODE == 0) { CU(bobj.alloc(n * 4)); CU(btri.alloc(n * 4)); CU(bbary.alloc(n * 16)); }
    if (MODE != 1) CU(btt.alloc(n * 8)); else CU(bocc.alloc(n));
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(bo.p, o, n * 24, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(bd.p, d, n * 24, cudaMemcpyHostToDevice, st));
    if (t_max) CU(cudaMemcpyAsync(bt.p, t_max, n * 8, cudaMemcpyHostToDevice, st));
    int32_t rc = launch_batch<MODE>(sc, bo.as<double>(), bd.as<double>(), t_max ? bt.as<double>() : nullptr, n, ctx->d_cursor,
                                    bobj.as<uint32_t>(), btri.as<uint32_t>(), btt.as<double>(), bbary.as<double>(), bocc.as<uint8_t>());
    if (rc != LUMO_OK) return rc;
    if (MODE == 0) {
        CU(cudaMemcpyAsync(obj, bobj.p, n * 4, cudaMemcpyDeviceToHost, st)); CU(cudaMemcpyAsync(tri, btri.p, n * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(bary, bbary.p, n * 16, cudaMemcpyDeviceToHost, st));
    }
    if (MODE != 1) CU(cudaMemcpyAsync(t, btt.p, n * 8, cudaMemcpyDeviceToHost, st)); else CU(cudaMemcpyAsync(occ, bocc.p, n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return LUMO_OK;
}

extern "C" int32_t lumo_gpu_trace_closest(lumo_scene* sc, const double* o, const double* d, const double* t_max, uint64_t n, uint32_t* obj, uint32_t* tri, double* t, double* bary) {
    return trace_host<0>(sc, o, d, t_max, n, obj, tri, t, bary, nullptr);
}
extern "C" int32_t lumo_gpu_trace_any(lumo_scene* sc, const double* o, const double* d, const double* t_max, uint64_t n, uint8_t* occluded) {
    return trace_host<1>(sc, o, d, t_max, n, nullptr, nullptr, nullptr, nullptr, occluded);
}
extern "C" int32_t lumo_gpu_trace_first_found(lumo_scene* sc, const double* o, const double* d, uint64_t n, double* t) {
    return trace_host<2>(sc, o, d, nullptr, n, nullptr, nullptr, t, nullptr, nullptr);
}
extern "C" int32_t lumo_gpu_trace_closest_dev(lumo_scene* sc, const double* o_dev, const double* d_dev, uint64_t n, uint32_t* obj_dev, uint32_t* tri_dev, double* t_dev,
                                              double* bary_dev, float* kernel_ms) {
    if (!sc || !o_dev || !d_dev || !obj_dev || !tri_dev || !t_dev || !bary_dev) return fail(LUMO_ERR_INVALID, "trace_closest_dev: null pointer");
    lumo_ctx* ctx = sc->ctx;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    int32_t rc = launch_batch<0>(sc, o_dev, d_dev, nullptr, n, ctx->d_cursor, obj_dev, tri_dev, t_dev, bary_dev, nullptr);
    if (rc != LUMO_OK) return rc;
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (kernel_ms) CU(cudaEventElapsedTime(kernel_ms, ctx->ev0, ctx->ev1));
    return LUMO_OK;
}

// ---- render ----------------------------------------------------------------------------------------
static void carve_wave(Wave& W, Carver& c, uint32_t N, uint32_t shadow_cap, uint32_t n_tiles, size_t film_px) {
    W.n_slots = N; W.shadow_cap = shadow_cap;
    for (int b = 0; b < 2; b++) {
        W.ox[b] = c.take<double>(N); W.oy[b] = c.take<double>(N); W.oz[b] = c.take<double>(N); W.dx[b] = c.take<double>(N); W.dy[b] = c.take<double>(N); W.dz[b] = c.take<double>(N);
        W.gathered[b] = c.take<double>(4 * (size_t)N); W.draws[b] = c.take<uint32_t>(N);
    }
    for (int k = 0; k < LUMO_N_CLASSES; k++) W.cls[k] = c.take<uint32_t>(N);
    W.ht = c.take<double>(N); W.hb0 = c.take<double>(N); W.hb1 = c.take<double>(N); W.hb2 = c.take<double>(N); W.hobj = c.take<uint32_t>(N); W.htri = c.take<uint32_t>(N);
    W.radiance = c.take<double>(4 * (size_t)N); W.lam = c.take<double>(4 * (size_t)N);
    W.rx = c.take<double>(N); W.ry = c.take<double>(N);
    W.pixel = c.take<uint32_t>(N); W.sample = c.take<uint32_t>(N); W.depth = c.take<uint32_t>(N); W.flags = c.take<uint32_t>(N); W.witem = c.take<uint32_t>(N);
    W.active = c.take<uint32_t>(N);
    for (int b = 0; b < 2; b++) W.done[b] = c.take<uint32_t>(N);
    const size_t C = shadow_cap;
    W.sox = c.take<double>(C); W.soy = c.take<double>(C); W.soz = c.take<double>(C); W.sdx = c.take<double>(C); W.sdy = c.take<double>(C); W.sdz = c.take<double>(C);
    W.stmax = c.take<double>(C); W.sc = c.take<double>(4 * C); W.sslot = c.take<uint32_t>(C);
    W.oq_i = c.take<uint32_t>(C); W.oq_obj = c.take<uint32_t>(C); W.oq_fb = c.take<uint32_t>(C); W.occ_record = c.take<uint8_t>(C);
    W.ch_t1 = c.take<double>(N); W.ch_tl = c.take<double>(N); W.ch_o1 = c.take<uint32_t>(N); W.ch_ol = c.take<uint32_t>(N); W.ch_flags = c.take<uint32_t>(N); W.ch_fb = c.take<uint32_t>(N);
    W.nee_ctx = c.take<double>((size_t)LUMO_NEE_CTX_DOUBLES * N); W.nee_meta = c.take<uint32_t>(N); W.neeq = c.take<uint32_t>(N);
    { NeeTermQueue& T = W.tq; T.ox = c.take<double>(C); T.oy = c.take<double>(C); T.oz = c.take<double>(C); T.dx = c.take<double>(C); T.dy = c.take<double>(C); T.dz = c.take<double>(C);
      T.wx = c.take<double>(C); T.wy = c.take<double>(C); T.wz = c.take<double>(C); T.tmax = c.take<double>(C); T.p_lig = c.take<double>(C); T.pdf_light = c.take<double>(C);
      T.le = c.take<double>(4 * C); T.slot = c.take<uint32_t>(C); }
    { NeeSurvivors& V = W.surv; const size_t H = C + 1; V.slot = c.take<uint32_t>(H); V.li = c.take<uint32_t>(H); V.wx = c.take<double>(H); V.wy = c.take<double>(H); V.wz = c.take<double>(H); }
    W.it = c.take<IterCounters>(1); W.run = c.take<RunCounters>(1); W.qc = c.take<QueueCounters>(1);
    W.tile_delta = c.take<double>(n_tiles); W.tile_delta_next = c.take<double>(n_tiles);
    W.pilot_lum = c.take<double>((size_t)n_tiles * LUMO_PILOT_N); W.pilot_cost = c.take<uint32_t>((size_t)n_tiles * LUMO_PILOT_N);
    (void)film_px;
}

struct HostCounters { QueueCounters qc; RunCounters run; };

// Runs waves until the work counter is exhausted and no path is alive.
template <int K>
static void launch_shade_kind(lumo_scene* sc, const Wave& W, const WaveParams& P, int grid, int nee_grid, cudaStream_t st, unsigned long long& launches) {
    if (!(sc->kind_mask & (1u << K))) return;
    // no LumoTexture record in the scene: the texture-free instantiations (shade.cuh LUMO_K_SOLID)
    if (sc->has_textures) {
        k_scatter<K><<<grid, 128, 0, st>>>(sc->S, W, P);
#if LUMO_NEE_A_SPLIT
        k_nee_a1<<<nee_grid, 128, 0, st>>>(sc->S, W, P, (uint32_t)K); launches++;
        k_nee_b<K><<<nee_grid, 128, 0, st>>>(sc->S, W, P);
        k_nee_a<true><<<nee_grid, 128, 0, st>>>(sc->S, W, P, (uint32_t)K);
#else
        k_nee_a<true><<<nee_grid, 128, 0, st>>>(sc->S, W, P, (uint32_t)K);
        k_nee_b<K><<<nee_grid, 128, 0, st>>>(sc->S, W, P);
#endif
        k_nee_eval<K><<<nee_grid, 128, 0, st>>>(sc->S, W, P);
    } else {
        k_scatter<K | LUMO_K_SOLID><<<grid, 128, 0, st>>>(sc->S, W, P);
#if LUMO_NEE_A_SPLIT
        k_nee_a1<<<nee_grid, 128, 0, st>>>(sc->S, W, P, (uint32_t)K); launches++;
        k_nee_b<K | LUMO_K_SOLID><<<nee_grid, 128, 0, st>>>(sc->S, W, P);
        k_nee_a<false><<<nee_grid, 128, 0, st>>>(sc->S, W, P, (uint32_t)K);
#else
        k_nee_a<false><<<nee_grid, 128, 0, st>>>(sc->S, W, P, (uint32_t)K);
        k_nee_b<K | LUMO_K_SOLID><<<nee_grid, 128, 0, st>>>(sc->S, W, P);
#endif
        k_nee_eval<K | LUMO_K_SOLID><<<nee_grid, 128, 0, st>>>(sc->S, W, P);
    }
    k_terms_reset<<<1, 1, 0, st>>>(W.it);
    launches += 5;
}

// Runs waves until the work counter is exhausted and no path is alive.
static int32_t run_wave(lumo_scene* sc, const Wave& W, WaveParams P, uint64_t& iterations) {
    lumo_ctx* ctx = sc->ctx;
    cudaStream_t st = ctx->stream;
    HostCounters* hc = (HostCounters*)ctx->host_pinned;
    const int tgrid = ctx->sm_count * LUMO_WAVE_TRACE_BLOCKS;   // persistent warps: one grid of resident CTAs
    const int rgrid = (int)std::min<uint64_t>((W.n_slots + 255) / 256, (uint64_t)ctx->sm_count * 16);
    const int sgrid = ctx->sm_count * 16;
    const int ngrid = (int)std::min<uint64_t>(((uint64_t)W.n_slots * sc->S.P.n_shadow_rays + 127) / 128, (uint64_t)ctx->sm_count * 32);
    CU(cudaMemsetAsync(W.flags, 0, (size_t)W.n_slots * 4, st));
    CU(cudaMemsetAsync(&W.run->next_work, 0, 8, st));
    // every slot starts out free: the first k_retire finds all of them in done[1] (no PF_DONE flag -> nothing to film)
    const QueueCounters qc0 = {{0u, 0u}, {0u, W.n_slots}};
    CU(cudaMemcpyAsync(W.qc, &qc0, sizeof qc0, cudaMemcpyHostToDevice, st));
    k_iota<<<rgrid, 256, 0, st>>>(W.done[1], W.n_slots);
    ctx->launches++;
    P.cur = 0;
    // Kernels read their queue sizes from device memory, so several iterations are enqueued back to back
    // and the host looks at the live-path count only once per batch (an iteration over empty queues costs
    // a few microseconds).
    const int BATCH = LUMO_ITER_BATCH;
    uint32_t main_iter = 0;
    for (;;) {
        for (int b = 0; b < BATCH; b++) {
            cudaEvent_t* ev = ctx->kev + 5 * b;
            CU(cudaMemsetAsync(W.it, 0, sizeof(IterCounters), st));
            CU(cudaEventRecord(ev[0], st));
            k_retire<<<rgrid, 256, 0, st>>>(sc->S, W, P);
            k_compact<<<rgrid, 256, 0, st>>>(W);
            CU(cudaEventRecord(ev[1], st));
            if (ctx->closest_faithful || sc->tiny) {
                if (ctx->count_visits) k_wave_trace<true><<<tgrid, 128, 0, st>>>(sc->S, W, P.cur, ctx->d_visit);
                else k_wave_trace<false><<<tgrid, 128, 0, st>>>(sc->S, W, P.cur, nullptr);
            } else {
                WaveRaySource src{W, P.cur}; WaveHitSink sink{W};
                ClosestScratch Q; Q.t1 = W.ch_t1; Q.tl = W.ch_tl; Q.o1 = W.ch_o1; Q.ol = W.ch_ol; Q.flags = W.ch_flags; Q.fallback_i = W.ch_fb; Q.counters = W.it->closest;
                int32_t rc = launch_closest(sc, src, sink, Q, st); if (rc != LUMO_OK) return rc;
                k_wave_classify<<<rgrid, 256, 0, st>>>(sc->S, W);
            }
            CU(cudaEventRecord(ev[2], st));
            k_terminal<<<sgrid, 128, 0, st>>>(sc->S, W, P);
            ctx->launches += 3;
            launch_shade_kind<LMAT_LAMBERTIAN>(sc, W, P, sgrid, ngrid, st, ctx->launches);
            launch_shade_kind<LMAT_MFDIFFUSE>(sc, W, P, sgrid, ngrid, st, ctx->launches);
            launch_shade_kind<LMAT_MFCONDUCTOR>(sc, W, P, sgrid, ngrid, st, ctx->launches);
            launch_shade_kind<LMAT_MFDIELECTRIC>(sc, W, P, sgrid, ngrid, st, ctx->launches);
            CU(cudaEventRecord(ev[3], st));
            if (ctx->occl_faithful || (sc->tiny && !ctx->occl_check)) {
                if (ctx->count_visits) k_wave_occlude<true><<<tgrid, 128, 0, st>>>(sc->S, W, ctx->d_visit + 1, nullptr, nullptr);
                else k_wave_occlude<false><<<tgrid, 128, 0, st>>>(sc->S, W, nullptr, nullptr, nullptr);
                ctx->launches++;
            } else {
                WaveShadowSource src{W}; WaveShadowSink sink{W, W.occ_record};
                OcclQueues Q; Q.confirm_i = W.oq_i; Q.confirm_obj = W.oq_obj; Q.fallback_i = W.oq_fb; Q.counters = W.it->occl;
                int32_t rc = launch_occlusion(sc, src, sink, Q, st); if (rc != LUMO_OK) return rc;
                if (ctx->occl_check) k_wave_occlude<false><<<tgrid, 128, 0, st>>>(sc->S, W, nullptr, W.occ_record, ctx->d_ah);
                else k_shadow_apply<<<sgrid, 256, 0, st>>>(W, W.occ_record);
                ctx->launches++;
            }
            CU(cudaEventRecord(ev[4], st));
            k_queue_reset<<<1, 1, 0, st>>>(W.qc, P.cur, W.it, P.mode == WM_MAIN ? ctx->d_iter_log : nullptr, main_iter, LUMO_ITER_LOG_CAP);
            if (P.mode == WM_MAIN) main_iter++;
            ctx->launches += 2; iterations++;
            P.cur ^= 1u;
        }
        CU(cudaMemcpyAsync(&hc->qc, W.qc, sizeof(QueueCounters), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&hc->run, W.run, sizeof(RunCounters), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        if (P.mode == WM_MAIN) for (int b = 0; b < BATCH; b++) for (int k = 0; k < 4; k++) {
            float ms = 0; CU(cudaEventElapsedTime(&ms, ctx->kev[5 * b + k], ctx->kev[5 * b + k + 1])); ctx->kernel_ms[k] += ms; ctx->kernel_launches[k]++;
        }
        if (hc->qc.n_active[P.cur] == 0 && hc->qc.n_done[P.cur ^ 1u] == 0) break;   // no survivors and nothing left to retire
    }
    if (P.mode == WM_MAIN) ctx->iter_log_n = main_iter < LUMO_ITER_LOG_CAP ? main_iter : LUMO_ITER_LOG_CAP;
    return LUMO_OK;
}

// BDPT: batches of camera samples through walk -> scan -> connect -> finish (bdpt.cuh)
// Samples per BDPT batch.  Every bounce of the walk wavefront costs at least one traversal's latency (~0.3 ms once a few
// thousand subpaths are left), and long specular chains give 70+ bounces: large batches pay that tail once.  A sample
// owns 2 x LUMO_BDPT_MAXV vertices in place (27 KB; longer subpaths take blocks from an overflow pool, up to the reference's
// 1024), so 2^20 samples are 28 GB of the 180 GB — sized down if memory is short.
#define LUMO_BDPT_BATCH_MAX (1u << 20)
#define LUMO_BW_TAIL 32768u          /* live subpaths at or below which the walk wavefront hands over to k_bw_tail (walks, caustics 1 spp: 0: 133.9 ms, 2048: 130.4, 8192: 127.1, 32768: 123.2, 131072: 130.1) */
#define LUMO_BDPT_QUEUE (1u << 22)   /* terms per chunk of the visibility-ray queue (100 B each) */
struct BdptStorage { BdptBatch B; void* scan_tmp = nullptr; size_t scan_bytes = 0; };
static void bdpt_carve(BdptBatch& B, Carver& c, uint32_t cap) {
    B.cap = cap;
    B.lp = c.take<Vtx>((size_t)cap * LUMO_BDPT_MAXV); B.cp = c.take<Vtx>((size_t)cap * LUMO_BDPT_MAXV);
    // overflow blocks for subpaths longer than LUMO_BDPT_MAXV vertices (specular chains): one block per 16 samples is ~60x what the
    // caustics scene uses; a subpath is cut (and counted) only when the pool is exhausted
    B.pool_blocks = std::max(1024u, cap / 16u);
    B.pool = c.take<Vtx>((size_t)B.pool_blocks * LUMO_BDPT_MAXV); B.pool_next = c.take<uint32_t>(4);
    B.ovf_l = c.take<uint32_t>((size_t)cap * LUMO_BDPT_OVF_BLOCKS); B.ovf_c = c.take<uint32_t>((size_t)cap * LUMO_BDPT_OVF_BLOCKS);
    B.ns = c.take<int>(cap); B.nt = c.take<int>(cap);
    B.lam = c.take<double>(4 * (size_t)cap); B.rx = c.take<double>(cap); B.ry = c.take<double>(cap); B.radiance = c.take<double>(4 * (size_t)cap);
    B.pixel = c.take<uint32_t>(cap); B.sample = c.take<uint32_t>(cap); B.draws = c.take<uint32_t>(cap); B.witem = c.take<uint32_t>(cap); B.valid = c.take<uint32_t>(cap);
    for (int k = 0; k < 3; k++) { B.n_terms[k] = c.take<unsigned long long>(cap + 1); B.term_off[k] = c.take<unsigned long long>(cap + 1); }
    B.w_ray = c.take<double>(6 * (size_t)cap); B.w_cam = c.take<double>(6 * (size_t)cap); B.w_gathered = c.take<double>(4 * (size_t)cap);
    B.w_pdf_fwd = c.take<double>(cap); B.w_delta = c.take<double>(cap);
    B.w_phase = c.take<uint32_t>(cap); B.w_n = c.take<uint32_t>(cap); B.w_depth = c.take<uint32_t>(cap);
    B.w_ht = c.take<double>(cap); B.w_hb0 = c.take<double>(cap); B.w_hb1 = c.take<double>(cap); B.w_hb2 = c.take<double>(cap);
    B.w_hobj = c.take<uint32_t>(cap); B.w_htri = c.take<uint32_t>(cap); B.w_have = c.take<uint32_t>(cap);
    B.act[0] = c.take<uint32_t>(cap); B.act[1] = c.take<uint32_t>(cap); B.n_act = c.take<uint32_t>(4);
    const uint32_t qcap = LUMO_BDPT_QUEUE; B.qcap = qcap;
    B.q_term = c.take<unsigned long long>(qcap); B.q_ray = c.take<double>(6 * (size_t)qcap);
    B.q_ht = c.take<double>(qcap); B.q_hb0 = c.take<double>(qcap); B.q_hb1 = c.take<double>(qcap); B.q_hb2 = c.take<double>(qcap);
    B.q_hobj = c.take<uint32_t>(qcap); B.q_htri = c.take<uint32_t>(qcap); B.q_have = c.take<uint32_t>(qcap); B.q_n = c.take<uint32_t>(4);
}
// batch storage for up to `want` samples, kept by the context and grown on demand
static int32_t bdpt_alloc(lumo_ctx* ctx, BdptStorage& st, unsigned long long want) {
    unsigned long long cap_max = LUMO_BDPT_BATCH_MAX;
    if (const char* e = std::getenv("LUMO_BDPT_BATCH")) cap_max = std::max<unsigned long long>(1024ull, (unsigned long long)std::atoll(e));
    uint32_t cap = (uint32_t)std::min<unsigned long long>(std::max<unsigned long long>((want + 1023ull) & ~1023ull, 1024ull), cap_max);
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    for (;;) {
        Carver dry{nullptr}; BdptBatch tmp; bdpt_carve(tmp, dry, cap);
        if (dry.off <= ctx->bdpt_bytes || dry.off <= (free_b + ctx->bdpt_bytes) / 2 || cap <= 16384u) {
            if (dry.off > ctx->bdpt_bytes) {
                if (ctx->bdpt_mem) { cudaFree(ctx->bdpt_mem); ctx->bdpt_mem = nullptr; ctx->bdpt_bytes = 0; }
                CU(cudaMalloc(&ctx->bdpt_mem, dry.off)); ctx->bdpt_bytes = dry.off;
            }
            break;
        }
        cap /= 2;
    }
    Carver c{(uint8_t*)ctx->bdpt_mem}; bdpt_carve(st.B, c, cap);
    size_t need = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, need, st.B.n_terms[0], st.B.term_off[0], (int)(cap + 1)));
    if (need > ctx->bdpt_scan_bytes) {
        if (ctx->bdpt_scan) { cudaFree(ctx->bdpt_scan); ctx->bdpt_scan = nullptr; ctx->bdpt_scan_bytes = 0; }
        CU(cudaMalloc(&ctx->bdpt_scan, need ? need : 16)); ctx->bdpt_scan_bytes = need;
    }
    st.scan_tmp = ctx->bdpt_scan; st.scan_bytes = ctx->bdpt_scan_bytes;
    return LUMO_OK;
}
static int32_t run_bdpt(lumo_scene* sc, const Wave& W, const WaveParams& P, BdptStorage& bs, uint64_t& iterations) {
    lumo_ctx* ctx = sc->ctx;
    cudaStream_t st = ctx->stream;
    const BdptBatch& B = bs.B;
    cudaEvent_t* ev = ctx->kev;
    static const uint32_t bw_tail = std::getenv("LUMO_BW_TAIL") ? (uint32_t)std::atoll(std::getenv("LUMO_BW_TAIL")) : LUMO_BW_TAIL;
    static const bool bdpt_log = std::getenv("LUMO_BDPT_LOG") != nullptr;   // diagnostics: live subpaths and elapsed time every 8 bounces
    for (unsigned long long w0 = 0; w0 < P.total_work; w0 += B.cap) {
        const uint32_t n = (uint32_t)std::min<unsigned long long>(B.cap, P.total_work - w0);
        CU(cudaEventRecord(ev[0], st));
        CU(cudaMemsetAsync(B.n_act, 0, 16, st));
        CU(cudaMemsetAsync(B.pool_next, 0, 16, st));
        k_bw_setup<<<ctx->sm_count * 8, 128, 0, st>>>(sc->S, W, P, B, w0, n);
        ctx->launches += 1;
        for (uint32_t it = 0;;) {                                        // one bounce of every live subpath per iteration; 8 iterations per host look
            for (int k = 0; k < 8; k++, it++) {
                k_bw_trace<<<ctx->sm_count * 8, 128, 0, st>>>(sc->S, W, B, it & 1u);
                k_bw_step<<<ctx->sm_count * 8, 128, 0, st>>>(sc->S, W, P, B, it & 1u);
            }
            ctx->launches += 16;
            uint32_t live = 0;
            CU(cudaMemcpyAsync(&live, B.n_act + (it & 1u), 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            if (bdpt_log) { CU(cudaEventRecord(ev[4], st)); CU(cudaEventSynchronize(ev[4])); float ms = 0; CU(cudaEventElapsedTime(&ms, ev[0], ev[4]));
                            std::fprintf(stderr, "[bdpt] batch %llu bounce %u live %u t %.3f ms\n", w0 / B.cap, it, live, ms); }
            if (live == 0) break;
            if (live <= bw_tail) {                                  // the stragglers run their walks to the end in one launch
                k_bw_tail<<<(live + 63u) / 64u, 64, 0, st>>>(sc->S, W, P, B, it & 1u);
                ctx->launches += 1;
                break;
            }
        }
        for (int k = 0; k < 3; k++) {
            CU(cudaMemsetAsync(B.n_terms[k] + n, 0, 8, st));
            CU(cub::DeviceScan::ExclusiveSum(bs.scan_tmp, bs.scan_bytes, B.n_terms[k], B.term_off[k], (int)(n + 1), st));
        }
        CU(cudaEventRecord(ev[1], st));
        k_bdpt_connect<BC_EMISSION><<<ctx->sm_count * 16, 128, 0, st>>>(sc->S, W, P, B, n);
        k_bdpt_connect<BC_NEE><<<ctx->sm_count * 16, 128, 0, st>>>(sc->S, W, P, B, n);
        unsigned long long totals[3] = {0, 0, 0};                        // terms per class of this batch
        for (int k = 0; k < 3; k++) CU(cudaMemcpyAsync(&totals[k], B.term_off[k] + n, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        auto queued = [&](auto cls_tag) {                               // light tracing / connections: prepare -> trace -> finish per chunk of terms
            constexpr int CLS = decltype(cls_tag)::value;
            for (unsigned long long t0 = 0; t0 < totals[CLS]; t0 += B.qcap) {
                const uint32_t count = (uint32_t)std::min<unsigned long long>(B.qcap, totals[CLS] - t0);
                cudaMemsetAsync(B.q_n, 0, 16, st);
                k_bq_prepare<CLS><<<ctx->sm_count * 16, 128, 0, st>>>(sc->S, W, P, B, n, t0, count);
                k_bq_trace<CLS><<<ctx->sm_count * 8, 128, 0, st>>>(sc->S, B);
                k_bq_finish<CLS><<<ctx->sm_count * 16, 128, 0, st>>>(sc->S, W, P, B);
                ctx->launches += 3;
            }
        };
        queued(std::integral_constant<int, BC_LIGHT_TRACE>());
        queued(std::integral_constant<int, BC_CONNECT>());
        CU(cudaEventRecord(ev[2], st));
        k_bdpt_finish<<<ctx->sm_count * 4, 256, 0, st>>>(sc->S, W, P, B, n);
        CU(cudaEventRecord(ev[3], st));
        ctx->launches += 4; iterations++;
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        if (P.mode == WM_MAIN) {   // kernel classes for BDPT: [1] walks (incl. their traversal), [2] connections, [0] finish / film
            float a = 0, b = 0, c = 0;
            CU(cudaEventElapsedTime(&a, ev[0], ev[1])); CU(cudaEventElapsedTime(&b, ev[1], ev[2])); CU(cudaEventElapsedTime(&c, ev[2], ev[3]));
            ctx->kernel_ms[1] += a; ctx->kernel_ms[2] += b; ctx->kernel_ms[0] += c;
            ctx->kernel_launches[0]++; ctx->kernel_launches[1]++; ctx->kernel_launches[2]++;
        }
    }
    return LUMO_OK;
}

static int32_t render_impl(lumo_scene* sc, const lumo_render_params* rp, double* pixels_dev, double* splats_dev, uint64_t* counters, double* tile_deltas_host, double* device_ms) {
    lumo_ctx* ctx = sc->ctx;
    const LumoSceneParams& SP = sc->S.P;
    if (rp->integrator < 0 || rp->integrator > 2) return fail(LUMO_ERR_INVALID, "render: unknown integrator");
    if (rp->sampler < 0 || rp->sampler > 3) return fail(LUMO_ERR_INVALID, "render: unknown sampler");
    if (rp->sampler == LUMO_SAMPLER_SOBOL && rp->total_spp > 1023u) return fail(LUMO_ERR_INVALID, "render: the Sobol sampler has 1023 points (samplers/sobol_seq.rs:3)");
    if (rp->tone_map < 0 || rp->tone_map > 2) return fail(LUMO_ERR_INVALID, "render: unknown tone map");
    if (rp->spp_end < rp->spp_begin || rp->spp_end > rp->total_spp || rp->total_spp == 0) return fail(LUMO_ERR_INVALID, "render: bad sample range");
    if (!(rp->rr_delta >= 0.0)) return fail(LUMO_ERR_INVALID, "render: rr_delta must be >= 0");
    const uint32_t Wd = SP.camera.res_x, Hd = SP.camera.res_y;
    const uint32_t tiles_x = (Wd + 15) / 16, tiles_y = (Hd + 15) / 16, n_tiles = tiles_x * tiles_y;
    const size_t film_px = (size_t)Wd * Hd;
    const uint32_t spp = rp->spp_end - rp->spp_begin;
    const unsigned long long main_work = (unsigned long long)n_tiles * 256ull * spp;
    uint32_t N = rp->wave_paths ? rp->wave_paths : (1u << 21);   // measured on B200: 2^21 slots amortise the per-iteration launches and the tail; 2^22 gains nothing more
    N = (uint32_t)std::min<unsigned long long>(N, std::max<unsigned long long>(main_work, (unsigned long long)n_tiles * LUMO_PILOT_N));
    N = std::max(N, 1024u);
    const uint32_t per_path = 2u * SP.n_shadow_rays;
    const uint32_t shadow_cap = (uint32_t)std::min<unsigned long long>((unsigned long long)N * per_path, 0xFFFFFF00ull);
    Wave W; std::memset(&W, 0, sizeof W);
    { Carver dry{nullptr}; carve_wave(W, dry, N, shadow_cap, n_tiles, film_px);
      if (dry.off > ctx->wave_bytes) {
          if (ctx->wave_mem) { cudaFree(ctx->wave_mem); ctx->wave_mem = nullptr; ctx->wave_bytes = 0; }
          CU(cudaMalloc(&ctx->wave_mem, dry.off)); ctx->wave_bytes = dry.off;
      } }
    Carver c{(uint8_t*)ctx->wave_mem}; carve_wave(W, c, N, shadow_cap, n_tiles, film_px);
    W.pixels = pixels_dev; W.splats = splats_dev;
    cudaStream_t st = ctx->stream;
    const unsigned long long launches0 = ctx->launches;
    for (int k = 0; k < 4; k++) { ctx->kernel_ms[k] = 0; ctx->kernel_launches[k] = 0; }
    CU(cudaEventRecord(ctx->ev0, st));
    CU(cudaMemsetAsync(pixels_dev, 0, film_px * 32, st));
    CU(cudaMemsetAsync(splats_dev, 0, film_px * 24, st));
    CU(cudaMemsetAsync(W.run, 0, sizeof(RunCounters), st));
    if (ctx->occl_check) CU(cudaMemsetAsync(&ctx->d_ah->mismatches, 0, 8, st));   // per render
    k_fill<<<64, 256, 0, st>>>(W.tile_delta, n_tiles, rp->rr_delta > 0.0 ? rp->rr_delta : 1e-5);
    ctx->launches++;
    WaveParams P; std::memset(&P, 0, sizeof P);
    P.seed = rp->seed; P.integrator = (uint32_t)rp->integrator; P.sampler = (uint32_t)rp->sampler; P.tone_map = (uint32_t)rp->tone_map; P.tone_map_arg = rp->tone_map_arg;
    P.spp_begin = rp->spp_begin; P.spp_count = spp; P.total_spp = rp->total_spp; P.tiles_x = tiles_x; P.tiles_y = tiles_y;
    { const char* e = std::getenv("LUMO_DEBUG_PIXEL"); P.debug_pixel = e ? (uint32_t)std::atoll(e) : LUMO_NONE; }
    P.check = ctx->occl_check ? 1u : 0u;
    uint64_t iterations = 0;
    const bool bdpt = rp->integrator == LUMO_BD_PATH_TRACE;
    BdptStorage bs;
    if (bdpt) { int32_t rc = bdpt_alloc(ctx, bs, std::max<unsigned long long>(main_work, (unsigned long long)n_tiles * LUMO_PILOT_N)); if (rc != LUMO_OK) return rc; }
    if (rp->rr_delta <= 0.0 && rp->integrator != LUMO_DIRECT_LIGHT) {
        // Per-tile Russian-roulette threshold (the role of task.rs:42-53): two pilot rounds, the second
        // using the first round's estimate.  Pilot paths never touch the film or the reported counters.
        for (uint32_t round = 0; round < 2; round++) {
            P.mode = WM_PILOT; P.pilot_round = round; P.total_work = (unsigned long long)n_tiles * LUMO_PILOT_N;
            int32_t rc = bdpt ? run_bdpt(sc, W, P, bs, iterations) : run_wave(sc, W, P, iterations); if (rc != LUMO_OK) return rc;
            k_pilot_reduce<<<(n_tiles + 127) / 128, 128, 0, st>>>(W, n_tiles, W.tile_delta_next);
            ctx->launches++;
            CU(cudaMemcpyAsync(W.tile_delta, W.tile_delta_next, (size_t)n_tiles * 8, cudaMemcpyDeviceToDevice, st));
        }
        CU(cudaStreamSynchronize(st));
        CU(cudaMemsetAsync(W.run, 0, sizeof(RunCounters), st));
    }
    P.mode = WM_MAIN; P.total_work = main_work;
    if (spp > 0) { int32_t rc = bdpt ? run_bdpt(sc, W, P, bs, iterations) : run_wave(sc, W, P, iterations); if (rc != LUMO_OK) return rc; }
    CU(cudaEventRecord(ctx->ev1, st));
    HostCounters* hc = (HostCounters*)ctx->host_pinned;
    CU(cudaMemcpyAsync(&hc->run, W.run, sizeof(RunCounters), cudaMemcpyDeviceToHost, st));
    if (tile_deltas_host) CU(cudaMemcpyAsync(tile_deltas_host, W.tile_delta, (size_t)n_tiles * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (hc->run.shadow_dropped && !bdpt) return fail(LUMO_ERR_CUDA, "render: shadow queue overflow (internal sizing error)");
    float ms = 0; CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (device_ms) *device_ms = ms;
    if (counters) {
        counters[0] = hc->run.camera_paths; counters[1] = hc->run.closest; counters[2] = hc->run.occlusion; counters[3] = hc->run.cost;
        counters[4] = ctx->launches - launches0; counters[5] = hc->run.max_depth; counters[6] = iterations; counters[7] = hc->run.nonfinite;
        counters[8] = bdpt ? hc->run.shadow_dropped : 0ull;   // BDPT: subpaths cut for lack of vertex storage
        ctx->shade_stats[0] = hc->run.nee_bounces; ctx->shade_stats[1] = hc->run.nee_terms; ctx->shade_stats[2] = hc->run.occlusion; ctx->shade_stats[3] = hc->run.closest;
        counters[9] = 0;
        if (ctx->occl_check) { AhCounters ac; CU(cudaMemcpy(&ac, ctx->d_ah, sizeof ac, cudaMemcpyDeviceToHost)); counters[9] = ac.mismatches + hc->run.prereject_bad; }
    }
    return LUMO_OK;
}

extern "C" int32_t lumo_gpu_render(lumo_scene* sc, const lumo_render_params* rp, lumo_film_accum* out) {
    if (!sc || !rp || !out || !out->pixels || !out->splats) return fail(LUMO_ERR_INVALID, "render: null pointer");
    lumo_ctx* ctx = sc->ctx;
    CU(cudaSetDevice(ctx->device));
    const size_t film_px = (size_t)sc->S.P.camera.res_x * sc->S.P.camera.res_y;
    // no allocation on the steady-state path: cudaMalloc / cudaFree were measured to stall a call by up to 0.3 s now and then
    if (film_px * 56 > ctx->film_bytes) {
        if (ctx->film_mem) { cudaFree(ctx->film_mem); ctx->film_mem = nullptr; ctx->film_bytes = 0; }
        CU(cudaMalloc(&ctx->film_mem, film_px * 56)); ctx->film_bytes = film_px * 56;
    }
    double* px = (double*)ctx->film_mem; double* sp = px + film_px * 4;
    int32_t rc = render_impl(sc, rp, px, sp, out->counters, out->tile_deltas, &out->device_ms);
    if (rc != LUMO_OK) return rc;
    CU(cudaMemcpyAsync(out->pixels, px, film_px * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out->splats, sp, film_px * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LUMO_OK;
}
extern "C" int32_t lumo_gpu_render_dev(lumo_scene* sc, const lumo_render_params* rp, double* pixels_dev, double* splats_dev, uint64_t* counters10, double* device_ms) {
    if (!sc || !rp || !pixels_dev || !splats_dev) return fail(LUMO_ERR_INVALID, "render_dev: null pointer");
    CU(cudaSetDevice(sc->ctx->device));
    return render_impl(sc, rp, pixels_dev, splats_dev, counters10, nullptr, device_ms);
}

// ---------------------------------------------------------------------------------------------------
// Film finalisation on the device: Film::rgb_image (src/tracer/film.rs:173-193) = Pixel::value
// (film.rs:82-90) + splat_scale * splat / filter.integral(), then ColorSpace::encode ->
// TransferFunction::apply (color/space.rs:8-36, 135-141) with Rust's saturating `as u8`.
// Pure streaming, 59 B per pixel (56 read as 128- and 64-bit loads, 3 written), HBM bound once the film is larger
// than a launch's fixed cost.  The host then reads 3 B/pixel instead of 56 B/pixel.
// ---------------------------------------------------------------------------------------------------
// TransferFunction::apply in f64, as the reference computes it.
__device__ __noinline__ uint32_t trc_apply(double c, int transfer) {
    double ec;
    if (transfer == 1) {   // rec. 2020
        const double beta = 0.018053968510807, alpha = 1.0 + 5.5 * beta;
        ec = c <= beta ? 4.5 * c : alpha * lm_pow(c, 0.45) - (alpha - 1.0);
    } else {
        ec = c <= 0.0031308 ? 12.92 * c : 1.055 * lm_pow(c, 1.0 / 2.4) - 0.055;
    }
    const double v = ec * 255.0;
    return !(v > 0.0) ? 0u : (v >= 255.0 ? 255u : (uint32_t)v);   // NaN and negatives -> 0, like `as u8`
}
// The f64 divisions and pow of a pixel only decide three 8-bit codes, and done for every pixel they make the kernel
// FP64-pipe bound (ncu: pipe 59 % busy at 27 % of HBM bandwidth, profiles/r1_ncu_film_encode.txt).  So a pixel is first
// evaluated in f32.  For non-negative terms and a weight in [1e-20, 1e20] the f32 linear value is within 3e-7
// (relative) of the f64 one, and 255*ec computed from it is within
//   tier 1: 3.4e-4 with __powf (ex2.approx(y * lg2.approx(x)): 1e-6 relative for x in [0.003, 1.1], y < 1),
//   tier 2: 1.9e-4 with powf (4 ulp)
// of the f64 result (both curves are continuous at their knee to 1e-7, so the side of the knee does not matter).
// A code is taken from tier 1 when the value is at least 1.5e-3 away from the neighbouring codes, else from tier 2 at
// 5e-4; what is left (0.1 % of the values, NaN, negative terms, weights or splat factors
// outside the f32-safe range) takes the reference's f64 path.  tests/test_film_encode.py compares the result with the
// f64-only kernel (LUMO_FILM_F64=1) byte for byte and walks every code boundary.
template <bool ACCURATE>
__device__ __forceinline__ bool trc_estimate(float c, int transfer, uint32_t& code) {
    float vf;
    if (transfer == 1) {
        const float beta = 0.018053968510807f, alpha = 1.0f + 5.5f * beta;
        vf = 255.0f * (c <= beta ? 4.5f * c : alpha * (ACCURATE ? powf(c, 0.45f) : __powf(c, 0.45f)) - (alpha - 1.0f));
    } else {
        vf = 255.0f * (c <= 0.0031308f ? 12.92f * c : 1.055f * (ACCURATE ? powf(c, 1.0f / 2.4f) : __powf(c, 1.0f / 2.4f)) - 0.055f);
    }
    if (vf >= 255.5f) { code = 255u; return true; }
    const float margin = ACCURATE ? 5e-4f : 1.5e-3f;
    const float fl = floorf(vf), fr = vf - fl;
    if (vf >= 1.0f && fr > margin && fr < 1.0f - margin) { code = (uint32_t)fl; return true; }
    if (vf >= 0.0f && vf < 1.0f - margin) { code = 0u; return true; }   // dark pixels: the linear segment, absolute error below 1e-6
    return false;   // also NaN
}
__device__ __forceinline__ void film_pixel_rgb8(const double* __restrict__ px, const double* __restrict__ sp, size_t i, double splat_scale,
                                                double filter_integral, int transfer, float splat_f, uint32_t out[3]) {
    const double2 a = __ldg((const double2*)(px + 4 * i)), b = __ldg((const double2*)(px + 4 * i) + 1);
    const double c[3] = {a.x, a.y, b.x}, w = b.y;
    const double s[3] = {__ldg(sp + 3 * i), __ldg(sp + 3 * i + 1), __ldg(sp + 3 * i + 2)};
    const float wf = (float)w, rw = __frcp_rn(wf);
    const bool w_ok = splat_f >= 0.0f && wf >= 1e-20f && wf <= 1e20f;     // splat_f < 0: the host found the splat factor unsafe for f32
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        const float cf = (float)c[k], sf = (float)s[k];
        if (w_ok && cf >= 0.0f && sf >= 0.0f) {
            const float lin = cf * rw + splat_f * sf;
            if (trc_estimate<false>(lin, transfer, out[k])) continue;
            if (trc_estimate<true>(lin, transfer, out[k])) continue;
        }
        out[k] = trc_apply(c[k] / w + splat_scale * s[k] / filter_integral, transfer);
    }
}
// One pixel per thread; a warp's 96 bytes are staged in shared memory and leave as 24 aligned 32-bit words, so the only
// synchronisation is within the warp (a block-wide barrier made seven warps wait for the one with an f64 pixel).
// Warps stride over tiles of 32 pixels.
__global__ void __launch_bounds__(256, 6) k_film_encode(const double* __restrict__ px, const double* __restrict__ sp, size_t n_pixels, double splat_scale,
                                                        double filter_integral, int transfer, float splat_f, uint8_t* __restrict__ rgb) {
    __shared__ __align__(16) uint8_t stage_all[8][96];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint8_t* stage = stage_all[warp];
    const size_t tiles = (n_pixels + 31) / 32, stride = (size_t)gridDim.x * 8;
    for (size_t tile = (size_t)blockIdx.x * 8 + warp; tile < tiles; tile += stride) {
        const size_t base = tile * 32, i = base + lane;
        if (i < n_pixels) {
            uint32_t v[3]; film_pixel_rgb8(px, sp, i, splat_scale, filter_integral, transfer, splat_f, v);
            stage[3 * lane] = (uint8_t)v[0]; stage[3 * lane + 1] = (uint8_t)v[1]; stage[3 * lane + 2] = (uint8_t)v[2];
        }
        __syncwarp();
        const size_t n_here = n_pixels - base < 32 ? n_pixels - base : 32, bytes = 3 * n_here;   // 3 * base = 96 * tile: word aligned
        if (lane < 24) {
            const size_t b = 4 * (size_t)lane;
            if (b + 4 <= bytes) ((uint32_t*)(rgb + 3 * base))[lane] = ((const uint32_t*)stage)[lane];
            else for (size_t q = b; q < bytes; q++) rgb[3 * base + q] = stage[q];
        }
        __syncwarp();
    }
}
static int32_t film_encode_impl(lumo_ctx* ctx, const double* px_dev, const double* sp_dev, uint64_t n_pixels, double splat_scale, double filter_integral,
                                int32_t transfer, uint8_t* rgb8_host, float* kernel_ms) {
    if (transfer != 0 && transfer != 1) return fail(LUMO_ERR_INVALID, "film_encode: transfer must be 0 (sRGB curve) or 1 (rec. 2020 curve)");
    if (n_pixels == 0) return LUMO_OK;
    const size_t out_bytes = (size_t)n_pixels * 3, out_cap = (out_bytes + 15) & ~(size_t)15;
    if (out_cap > ctx->rgb_bytes) {
        if (ctx->rgb_mem) { cudaFree(ctx->rgb_mem); ctx->rgb_mem = nullptr; ctx->rgb_bytes = 0; }
        CU(cudaMalloc(&ctx->rgb_mem, out_cap)); ctx->rgb_bytes = out_cap;
    }
    const unsigned grid = (unsigned)std::min<size_t>((n_pixels + 255) / 256, (size_t)ctx->sm_count * 24);   // 6 resident CTAs per SM, 4 rounds
    cudaStream_t st = ctx->stream;
    CU(cudaEventRecord(ctx->ev0, st));
    // the f32 pre-pass needs a splat factor that f32 represents well; otherwise (splat_f = -1) every pixel takes the f64 path
    const bool fast_ok = splat_scale >= 0.0 && splat_scale <= 1e10 && filter_integral >= 1e-10 && filter_integral <= 1e10 && !std::getenv("LUMO_FILM_F64");
    const float splat_f = fast_ok ? (float)(splat_scale / filter_integral) : -1.0f;
    k_film_encode<<<grid, 256, 0, st>>>(px_dev, sp_dev, (size_t)n_pixels, splat_scale, filter_integral, transfer, splat_f, (uint8_t*)ctx->rgb_mem);
    CU(cudaGetLastError());
    ctx->launches++;
    CU(cudaEventRecord(ctx->ev1, st));
    CU(cudaMemcpyAsync(rgb8_host, ctx->rgb_mem, out_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (kernel_ms) CU(cudaEventElapsedTime(kernel_ms, ctx->ev0, ctx->ev1));
    return LUMO_OK;
}
extern "C" int32_t lumo_gpu_film_encode_dev(lumo_ctx* ctx, const double* pixels_dev, const double* splats_dev, uint64_t n_pixels, double splat_scale,
                                            double filter_integral, int32_t transfer, uint8_t* rgb8, float* kernel_ms) {
    if (!ctx || !pixels_dev || !splats_dev || !rgb8) return fail(LUMO_ERR_INVALID, "film_encode_dev: null pointer");
    CU(cudaSetDevice(ctx->device));
    return film_encode_impl(ctx, pixels_dev, splats_dev, n_pixels, splat_scale, filter_integral, transfer, rgb8, kernel_ms);
}
extern "C" int32_t lumo_gpu_film_encode(lumo_ctx* ctx, const double* pixels, const double* splats, uint64_t n_pixels, double splat_scale,
                                        double filter_integral, int32_t transfer, uint8_t* rgb8) {
    if (!ctx || !pixels || !splats || !rgb8) return fail(LUMO_ERR_INVALID, "film_encode: null pointer");
    CU(cudaSetDevice(ctx->device));
    if (n_pixels * 56 > ctx->film_bytes) {
        if (ctx->film_mem) { cudaFree(ctx->film_mem); ctx->film_mem = nullptr; ctx->film_bytes = 0; }
        CU(cudaMalloc(&ctx->film_mem, n_pixels * 56)); ctx->film_bytes = n_pixels * 56;
    }
    double* px = (double*)ctx->film_mem; double* sp = px + n_pixels * 4;
    CU(cudaMemcpyAsync(px, pixels, n_pixels * 32, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(sp, splats, n_pixels * 24, cudaMemcpyHostToDevice, ctx->stream));
    return film_encode_impl(ctx, px, sp, n_pixels, splat_scale, filter_integral, transfer, rgb8, nullptr);
}

// ---------------------------------------------------------------------------------------------------
// lumo_gpu_render_multi: one host process, n GPUs.  The scene is replicated (one lumo_scene per
// context), sample indices [spp_begin, spp_end) are cut into n contiguous ranges (sizes differ by at most one, like
// lumo_b200/distributed.py: sample_range), one host thread per GPU runs the wave pipeline into that GPU's film, and the
// films are summed on scenes[0]'s GPU by ONE kernel that loads the peers' accumulators straight over NVLink (peer
// memory; no staging copies, no NCCL dependency in this library) in the fixed order 0..n-1.  GPUs without peer access
// to scenes[0]'s are staged through cudaMemcpyPeerAsync first.
// ---------------------------------------------------------------------------------------------------
#define LUMO_MAX_MULTI 16
struct FilmPeers { const double* p[LUMO_MAX_MULTI]; };
__global__ void __launch_bounds__(256) k_film_reduce(FilmPeers src, int n, double* dst, size_t n_doubles) {
    const size_t n2 = n_doubles / 2;   // 7 doubles per pixel: W*H*7 may be odd -> 128-bit body + scalar tail
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        double2 acc = ((const double2*)src.p[0])[i];
        for (int g = 1; g < n; g++) { const double2 v = ((const double2*)src.p[g])[i]; acc.x += v.x; acc.y += v.y; }
        ((double2*)dst)[i] = acc;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (n_doubles & 1)) {
        double acc = src.p[0][n_doubles - 1];
        for (int g = 1; g < n; g++) acc += src.p[g][n_doubles - 1];
        dst[n_doubles - 1] = acc;
    }
}
#include <thread>
// The sample range of GPU g of n: contiguous, disjoint, covering [begin, end), sizes differing by at most one
// (the same rule as lumo_b200/distributed.py: sample_range).  Pure host arithmetic; exported so that it can be tested without a GPU.
extern "C" int32_t lumo_gpu_sample_range(int32_t g, int32_t n, uint32_t begin, uint32_t end, uint32_t* g_begin, uint32_t* g_end) {
    if (!g_begin || !g_end) return fail(LUMO_ERR_INVALID, "sample_range: null pointer");
    if (n < 1 || g < 0 || g >= n || end < begin) return fail(LUMO_ERR_INVALID, "sample_range: bad arguments");
    const uint32_t total = end - begin, base = total / (uint32_t)n, extra = total % (uint32_t)n;
    *g_begin = begin + (uint32_t)g * base + std::min<uint32_t>((uint32_t)g, extra);
    *g_end = *g_begin + base + ((uint32_t)g < extra ? 1u : 0u);
    return LUMO_OK;
}
static int32_t render_multi_impl(lumo_scene** scenes, int32_t n, const lumo_render_params* rp, lumo_film_accum* out, std::vector<void*>& staged);
extern "C" int32_t lumo_gpu_render_multi(lumo_scene** scenes, int32_t n, const lumo_render_params* rp, lumo_film_accum* out) {
    std::vector<void*> staged;          // peer films copied to GPU 0 where peer access is unavailable: freed on every exit
    int32_t rc;
    try { rc = render_multi_impl(scenes, n, rp, out, staged); }
    catch (const std::bad_alloc&) { rc = fail(LUMO_ERR_OOM, "render_multi: out of host memory"); }
    catch (const std::exception& e) { rc = fail(LUMO_ERR_CUDA, std::string("render_multi: ") + e.what()); }
    catch (...) { rc = fail(LUMO_ERR_CUDA, "render_multi: unknown exception"); }
    if (!staged.empty()) { const std::string msg = g_err; if (scenes && scenes[0]) cudaSetDevice(scenes[0]->ctx->device); for (void* t : staged) cudaFree(t); g_err = msg; }
    return rc;
}
static int32_t render_multi_impl(lumo_scene** scenes, int32_t n, const lumo_render_params* rp, lumo_film_accum* out, std::vector<void*>& staged) {
    if (!scenes || !rp || !out || !out->pixels || !out->splats) return fail(LUMO_ERR_INVALID, "render_multi: null pointer");
    if (n < 1 || n > LUMO_MAX_MULTI) return fail(LUMO_ERR_INVALID, "render_multi: n must be in [1, 16]");
    for (int g = 0; g < n; g++) {
        if (!scenes[g]) return fail(LUMO_ERR_INVALID, "render_multi: null scene");
        for (int h = 0; h < g; h++) if (scenes[h]->ctx == scenes[g]->ctx) return fail(LUMO_ERR_INVALID, "render_multi: every scene needs its own context");
        if (scenes[g]->S.P.camera.res_x != scenes[0]->S.P.camera.res_x || scenes[g]->S.P.camera.res_y != scenes[0]->S.P.camera.res_y)
            return fail(LUMO_ERR_INVALID, "render_multi: the scenes differ in film resolution");
    }
    if (rp->spp_end < rp->spp_begin) return fail(LUMO_ERR_INVALID, "render: bad sample range");
    const size_t film_px = (size_t)scenes[0]->S.P.camera.res_x * scenes[0]->S.P.camera.res_y, film_doubles = film_px * 7;
    struct Part { int32_t rc = LUMO_OK; std::string err; uint64_t counters[10] = {}; double ms = 0.0; };
    std::vector<Part> parts((size_t)n);
    std::vector<std::thread> workers;
    workers.reserve((size_t)n);
    struct Joiner { std::vector<std::thread>& w; ~Joiner() { for (auto& t : w) if (t.joinable()) t.join(); } } joiner{workers};   // also when a later thread cannot be started
    for (int g = 0; g < n; g++) {
        workers.emplace_back([&, g]() {
            Part& pt = parts[(size_t)g];
            lumo_scene* sc = scenes[g]; lumo_ctx* ctx = sc->ctx;
            pt.rc = [&]() -> int32_t {
                CU(cudaSetDevice(ctx->device));
                if (film_doubles * 8 > ctx->film_bytes) {
                    if (ctx->film_mem) { cudaFree(ctx->film_mem); ctx->film_mem = nullptr; ctx->film_bytes = 0; }
                    CU(cudaMalloc(&ctx->film_mem, film_doubles * 8)); ctx->film_bytes = film_doubles * 8;
                }
                lumo_render_params p = *rp;
                { const int32_t rc = lumo_gpu_sample_range(g, n, rp->spp_begin, rp->spp_end, &p.spp_begin, &p.spp_end); if (rc != LUMO_OK) return rc; }
                double* px = (double*)ctx->film_mem;
                return render_impl(sc, &p, px, px + film_px * 4, pt.counters, g == 0 ? out->tile_deltas : nullptr, &pt.ms);
            }();
            if (pt.rc != LUMO_OK) pt.err = g_err;   // g_err is per thread: carry the message to the caller's thread
        });
    }
    for (auto& w : workers) w.join();
    for (int g = 0; g < n; g++) if (parts[(size_t)g].rc != LUMO_OK) return fail(parts[(size_t)g].rc, "render_multi, gpu " + std::to_string(g) + ": " + parts[(size_t)g].err);
    // ---- the one exchange of the path: sum of the film accumulators on scenes[0]'s GPU -------------------
    lumo_ctx* c0 = scenes[0]->ctx;
    CU(cudaSetDevice(c0->device));
    cudaStream_t st = c0->stream;
    FilmPeers peers{};
    peers.p[0] = (const double*)c0->film_mem;
    for (int g = 1; g < n; g++) {
        lumo_ctx* cg = scenes[g]->ctx;
        bool direct = cg->device == c0->device;
        if (!direct) {
            int can = 0; CU(cudaDeviceCanAccessPeer(&can, c0->device, cg->device));
            if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(cg->device, 0); if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); direct = true; } else cudaGetLastError(); }
        }
        if (direct) peers.p[g] = (const double*)cg->film_mem;
        else {
            void* tmp = nullptr; CU(cudaMalloc(&tmp, film_doubles * 8)); staged.push_back(tmp);
            CU(cudaMemcpyPeerAsync(tmp, c0->device, cg->film_mem, cg->device, film_doubles * 8, st));
            peers.p[g] = (const double*)tmp;
        }
    }
    if (n > 1) {
        const unsigned grid = (unsigned)std::min<size_t>((film_doubles / 2 + 255) / 256, (size_t)c0->sm_count * 16);
        CU(cudaEventRecord(c0->ev0, st));
        k_film_reduce<<<std::max(grid, 1u), 256, 0, st>>>(peers, n, (double*)c0->film_mem, film_doubles);
        CU(cudaGetLastError());
        CU(cudaEventRecord(c0->ev1, st));
        c0->launches++;
    }
    const double* px = (const double*)c0->film_mem;
    CU(cudaMemcpyAsync(out->pixels, px, film_px * 32, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out->splats, px + film_px * 4, film_px * 24, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    float reduce_ms = 0.f; if (n > 1) CU(cudaEventElapsedTime(&reduce_ms, c0->ev0, c0->ev1));
    // counters: sums, except [5] deepest path (max) and [6] wave iterations (max); device_ms: slowest GPU + the reduce
    double ms = 0.0;
    for (int k = 0; k < 10; k++) out->counters[k] = 0;
    for (int g = 0; g < n; g++) {
        const Part& pt = parts[(size_t)g];
        for (int k = 0; k < 10; k++) out->counters[k] = (k == 5 || k == 6) ? std::max(out->counters[k], pt.counters[k]) : out->counters[k] + pt.counters[k];
        ms = std::max(ms, pt.ms);
    }
    if (n > 1) out->counters[4] += 1;   // k_film_reduce
    out->device_ms = ms + reduce_ms;
    return LUMO_OK;
}
