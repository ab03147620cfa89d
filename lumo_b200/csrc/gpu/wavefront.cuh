// Wavefront pipeline: the GPU replacement of ThreadPool + RenderTaskExecutor::exec
// (src/pool.rs:10-55, src/renderer/task.rs:25-81) and of the per-sample integrator loops
// (src/tracer/integrator/path_trace.rs:5-82, direct_light.rs:5-73, integrator.rs:74-184).
//
// A wave is N path slots whose state lives in HBM as structure-of-arrays.  One iteration is
//   regen  : retire finished paths into the film (tone map, XYZ->RGB, filter footprint clamped to
//            the 16x16 tile, f64 atomics), refill free slots from the global work counter, and
//            stream-compact the live slots into the active queue;
//   trace  : Scene::hit for every queued ray (persistent warps pulling 32-ray batches);
//   shade  : one bounce of the integrator per live path — Hit reconstruction, BSDF sampling, NEE
//            (light sampling + MIS; occlusion rays are appended to the shadow queue with their
//            already-weighted contribution), throughput update, Russian roulette;
//   occlude: Scene::hit_light's occlusion test for every shadow-queue entry; unoccluded
//            contributions are added to the owning path's radiance.
// Random numbers are Philox streams keyed by (seed, pixel, global sample index) and consumed in the
// reference's draw order, so a path is independent of wave size, scheduling and GPU
// count.
#pragma once
#include "shade.cuh"
#include "occlude.cuh"
#include "closest.cuh"
#include <cooperative_groups.h>
#include <cstdio>

namespace lumo_dev {
namespace cg = cooperative_groups;

enum { PF_ALIVE = 1u, PF_LAST_SPECULAR = 2u, PF_DONE = 4u, PF_NEE = 8u };
enum { WM_MAIN = 0, WM_PILOT = 1 };
#define LUMO_RR_DEPTH 5u           /* path_trace.rs:3 */
#define LUMO_DL_MAX_RECURSION 50u  /* direct_light.rs:6 */
#define LUMO_PILOT_N 64u
#define LUMO_N_CLASSES 5           /* shade queues: 0 = terminal (miss / light / blank), 1..4 = LumoMatKind of a Standard material */

struct IterCounters {   // zeroed before every iteration
    uint32_t n_shadow, trace_next, occl_next, n_active;
    uint32_t n_class[LUMO_N_CLASSES], pad[3];
    uint32_t occl[8];   // OcclQueues::counters of the occlusion-BVH kernels (occlude.cuh)
    uint32_t n_terms, n_nee, n_surv, pad2;   // NEE term queue; bounces of the current material family that run NEE (Wave::neeq); light-sampled terms that passed the sign tests (Wave::surv)
    uint32_t closest[4];         // ClosestScratch::counters of the closest-hit pipeline (closest.cuh)
};
// Counters that live across iterations (double-buffered by the parity of the iteration): done[p] = slots whose
// path ended in an iteration of parity p, retired into the film (and refilled) at the start of the next one;
// n_active[p] = how many paths survive into an iteration of parity p (termination test on the host).  The
// active queue itself is rebuilt in slot order every iteration (k_compact) so that the structure-of-arrays
// path state is read with coalesced accesses.
struct QueueCounters { uint32_t n_active[2], n_done[2]; };
struct RunCounters {    // zeroed once per render
    unsigned long long next_work, camera_paths, closest, occlusion, cost, shadow_queued, shadow_dropped, nonfinite;
    unsigned long long nee_bounces, nee_terms;   // bounces that ran NEE; terms that reached k_nee_eval (bench.py: designed queue traffic of the shading stage)
    uint32_t max_depth, prereject_bad;   // prereject_bad: WaveParams::check only — BSDF-sampled NEE terms whose light test passed although the light's bounding sphere said it could not
};

// Path state, structure of arrays.  Ray, throughput and RNG position are double-buffered: the scatter
// kernel writes the NEXT bounce's values into buffer cur^1 while the NEE kernel of the same iteration
// still reads this bounce's values from buffer cur.
struct NeeTermQueue;
struct Wave;
struct NeeSurvivors { uint32_t *slot, *li; double *wx, *wy, *wz; };    // k_nee_a1, k_nee_b -> k_nee_a; capacity shadow_cap = 2 * n_slots * n_shadow_rays; li: bit 31 term B, bit 30 check-mode flag
struct NeeTermQueue {   // NEE stage 1 -> stage 2 (see k_nee_a): structure of arrays, capacity = shadow_cap; reused by one material family after the other
    double *ox, *oy, *oz, *dx, *dy, *dz, *wx, *wy, *wz, *tmax, *p_lig, *pdf_light, *le;   // le[k * cap + i]
    uint32_t* slot;     // bit 31: the BSDF-sampled term (B)
};
struct Wave {
    uint32_t n_slots, shadow_cap;
    double *ox[2], *oy[2], *oz[2], *dx[2], *dy[2], *dz[2];
    double* gathered[2];                 // [k * n_slots + slot]
    uint32_t* draws[2];
    double *ht, *hb0, *hb1, *hb2; uint32_t *hobj, *htri;
    double *radiance, *lam;              // [k * n_slots + slot]
    double *rx, *ry;
    uint32_t *pixel, *sample, *depth, *flags, *witem;
    uint32_t* active; uint32_t* done[2];
    uint32_t* cls[LUMO_N_CLASSES];       // per-class shade queues (slot indices)
    // shadow queue (SoA)
    double *sox, *soy, *soz, *sdx, *sdy, *sdz, *stmax, *sc; uint32_t* sslot;
    double *ch_t1, *ch_tl; uint32_t *ch_o1, *ch_ol, *ch_flags, *ch_fb;   // closest-hit pipeline: per-ray scratch of k_closest_bvh, fallback queue (capacity n_slots)
    double* nee_ctx; uint32_t* nee_meta; NeeTermQueue tq;   // NEE: per-slot shading context [k * n_slots + slot] (17 doubles), term queue
    uint32_t* neeq;                                          // slots of the current material family whose bounce runs NEE (dense; written by k_scatter)
    NeeSurvivors surv;                                       // k_nee_a1 -> k_nee_a
    uint32_t *oq_i, *oq_obj, *oq_fb; uint8_t* occ_record;   // occlusion-BVH pipeline: confirm queue, fallback queue; verdicts (LUMO_OCCLUDE_CHECK only)
    IterCounters* it; RunCounters* run; QueueCounters* qc;
    // film + RR thresholds
    double *pixels, *splats, *tile_delta, *tile_delta_next;
    double* pilot_lum; uint32_t* pilot_cost;
};

struct WaveParams {
    unsigned long long seed, total_work;
    uint32_t integrator, sampler, tone_map, mode;
    double tone_map_arg;
    uint32_t spp_begin, spp_count, total_spp, pilot_round;
    uint32_t tiles_x, tiles_y, debug_pixel, cur;   // debug_pixel: LUMO_DEBUG_PIXEL env (printf trace of one pixel's paths), LUMO_NONE = off
    uint32_t check, pad[3];                        // check: the context's cross-check mode (lumo_gpu_ctx_occlusion_mode 2) — shortcuts are verified and disagreements counted
};                                                  // cur: which half of the double-buffered state holds this iteration's rays

__device__ __forceinline__ uint32_t agg_inc(uint32_t* ctr) {   // warp-aggregated atomicAdd(ctr, 1)
    cg::coalesced_group g = cg::coalesced_threads();
    uint32_t base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(ctr, g.size());
    return g.shfl(base, 0) + g.thread_rank();
}
__device__ __forceinline__ unsigned long long agg_inc64(unsigned long long* ctr) {
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(ctr, (unsigned long long)g.size());
    return g.shfl(base, 0) + g.thread_rank();
}
__device__ __forceinline__ uint32_t tile_of(const WaveParams& P, const DevScene& S, uint32_t pixel) {
    const uint32_t W = S.P.camera.res_x;
    return ((pixel / W) / 16u) * P.tiles_x + (pixel % W) / 16u;
}
__device__ __forceinline__ C4 load_c4(const double* a, uint32_t n, uint32_t slot) { C4 c; for (int k = 0; k < 4; k++) c.s[k] = a[(size_t)k * n + slot]; return c; }
__device__ __forceinline__ void store_c4(double* a, uint32_t n, uint32_t slot, const C4& c) { for (int k = 0; k < 4; k++) a[(size_t)k * n + slot] = c.s[k]; }

// Intra-pixel offset of sample s (samplers.rs:54-192).  Uniform and Jittered are the reference's
// formulas; MultiJittered keeps the reference's two-level stratification (samplers.rs:178-189) but
// the two Fisher-Yates tables become a keyed hash permutation, so that any sample index can be
// produced independently.
__device__ __forceinline__ void raster_jitter(const WaveParams& P, uint32_t pixel, uint32_t s, Rng& rng, double& jx, double& jy) {
    if (P.sampler == 0) { jx = rng_float(rng); jy = rng_float(rng); return; }
    if (P.sampler == 3) {
        // SobolSampler (samplers.rs:193-247, samplers/sobol_seq.rs): point s + 1 of two Gray-code Sobol dimensions of degree 10
        // (direction numbers m << (64 - i - 1)), XOR-scrambled with the pixel's sampler seed.  The reference steps
        // prev ^= V[ctz(n)]; unrolled that is the XOR of V[i] over the set bits of gray(n) = n ^ (n >> 1), so any sample
        // index is produced independently.  The seed is the pixel's Philox key where the reference draws it from the tile stream.
        const unsigned long long m1[10] = {1, 1, 7, 15, 5, 19, 69, 51, 121, 695}, m2[10] = {1, 1, 7, 7, 7, 53, 57, 229, 473, 533};
        Rng keyr = rng_make(P.seed, pixel, 0xFFFFFFFFu, 1u, 0u);
        const unsigned long long k = rng_u64(keyr);
        const unsigned long long n = (unsigned long long)s + 1ull, g = n ^ (n >> 1);
        unsigned long long a = 0ull, b = 0ull;
#pragma unroll
        for (int i = 0; i < 10; i++) if ((g >> i) & 1ull) { a ^= m1[i] << (64 - i - 1); b ^= m2[i] << (64 - i - 1); }
        jx = __ull2double_rn(a ^ k) * 5.421010862427522170037e-20; jy = __ull2double_rn(b ^ k) * 5.421010862427522170037e-20;
        return;
    }
    const unsigned long long total = P.total_spp;
    const unsigned long long dim = sat_u64(ceil(sqrt((double)total)));
    const double s0x = 1.0 / (double)dim, s0y = (double)dim / (double)total;
    const unsigned long long x0 = s % dim, y0 = s / dim;
    const double o0x = s0x * (double)x0, o0y = s0y * (double)y0;
    if (P.sampler == 1) { const double a = rng_float(rng), b = rng_float(rng); jx = s0x * a + o0x; jy = s0y * b + o0y; return; }
    Rng keyr = rng_make(P.seed, pixel, 0xFFFFFFFFu, 1u, 0u);
    const unsigned long long k = rng_u64(keyr);
    const uint32_t kx = (uint32_t)k, ky = (uint32_t)(k >> 32);
    const double s1x = s0x / (double)dim, s1y = s0y / (double)dim;
    const uint32_t x1 = cmj_permute((uint32_t)y0, (uint32_t)dim, kx), y1 = cmj_permute((uint32_t)x0, (uint32_t)dim, ky);
    const double o1x = s1x * (double)x1, o1y = s1y * (double)y1;
    const double a = rng_float(rng), b = rng_float(rng);
    jx = o0x + o1x + s1x * a; jy = o0y + o1y + s1y * b;
}

// ---- retire: film + refill, densely over the slots whose path ended in the previous iteration ----------
__global__ void k_iota(uint32_t* p, uint32_t n) { for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = i; }
// end of an iteration: recycle the queue counters and log the iteration's queue sizes (rays traced / shadow rays)
__global__ void k_queue_reset(QueueCounters* qc, uint32_t cur, const IterCounters* it, uint32_t* log, uint32_t iter, uint32_t log_cap) {
    qc->n_active[cur] = 0u; qc->n_done[cur ^ 1u] = 0u;
    if (log && iter < log_cap) { log[2u * iter] = it->n_active; log[2u * iter + 1u] = it->n_shadow; }
}

__global__ void __launch_bounds__(256) k_retire(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P) {
    const uint32_t N = W.n_slots, c = P.cur;
    const uint32_t n = W.qc->n_done[c ^ 1u];
    unsigned long long cost_sum = 0, paths = 0; uint32_t depth_max = 0;     // run counters: per thread here, one atomic per warp at the end
    for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < n; qi += gridDim.x * blockDim.x) {
        const uint32_t slot = W.done[c ^ 1u][qi];
        uint32_t f = W.flags[slot];
        if (f & PF_DONE) {
            // RenderTaskExecutor::exec tail (task.rs:64-76)
            Lam lam; for (int k = 0; k < 4; k++) lam.l[k] = W.lam[(size_t)k * N + slot];
            const C4 rad = load_c4(W.radiance, N, slot);
            const uint32_t depth = W.depth[slot];
            const uint32_t cost = P.integrator == 1 ? depth + 1u : depth;
            if (P.mode == WM_PILOT) {
                const uint32_t w = W.witem[slot];
                W.pilot_lum[w] = luminance(S, rad, lam); W.pilot_cost[w] = cost;
            } else {
                bool finite = true;
                for (int k = 0; k < 4; k++) finite = finite && isfinite(rad.s[k]);
                if (!finite) atomicAdd(&W.run->nonfinite, 1ull);
                film_add_sample(S, W.pixels, W.splats, tone_map(S, (int)P.tone_map, P.tone_map_arg, rad, lam), lam, W.rx[slot], W.ry[slot], false);
                cost_sum += cost; paths += 1ull; depth_max = max(depth_max, depth);
            }
        }
        f = 0;
        const uint32_t Wd = S.P.camera.res_x, Hd = S.P.camera.res_y;
        while (W.run->next_work < P.total_work) {       // cheap pre-check; the atomic decides.  Work items outside the image are skipped
            const unsigned long long w = agg_inc64(&W.run->next_work);
            if (w >= P.total_work) break;
            uint32_t px, py, sample; bool ok = true;
            if (P.mode == WM_PILOT) {                       // two pilot rounds of 64 paths per tile estimate the RR threshold
                const uint32_t tile = (uint32_t)(w / LUMO_PILOT_N), k = (uint32_t)(w % LUMO_PILOT_N);
                const uint32_t x0 = (tile % P.tiles_x) * 16u, y0 = (tile / P.tiles_x) * 16u;
                px = min(x0 + 2u * (k % 8u), Wd - 1u); py = min(y0 + 2u * (k / 8u), Hd - 1u);
                sample = 0xFFFFFF00u + P.pilot_round;
            } else {                                          // sample-major, tile by tile, 8x4 pixel blocks per warp
                const unsigned long long per_s = (unsigned long long)P.tiles_x * P.tiles_y * 256ull;
                const uint32_t si = (uint32_t)(w / per_s); const unsigned long long r = w % per_s;
                const uint32_t tile = (uint32_t)(r / 256ull), q = (uint32_t)(r % 256ull);
                const uint32_t blk = q / 32u, in = q % 32u;   // 8 blocks of 8x4 in a 16x16 tile
                px = (tile % P.tiles_x) * 16u + (blk % 2u) * 8u + (in % 8u);
                py = (tile / P.tiles_x) * 16u + (blk / 2u) * 4u + (in / 8u);
                sample = P.spp_begin + si;
                ok = px < Wd && py < Hd;
            }
            if (!ok) continue;
            const uint32_t pixel = px + py * Wd;
            Rng rng = rng_make(P.seed, pixel, sample, 0u, 0u);
            double jx, jy;
            if (P.mode == WM_PILOT) { jx = rng_float(rng); jy = rng_float(rng); } else raster_jitter(P, pixel, sample, rng, jx, jy);
            const double rx = (double)px + jx, ry = (double)py + jy;
            const double l0 = rng_float(rng), l1 = rng_float(rng);                 // integrator.rs:56-57
            const Ray r = camera_generate_ray(S.P.camera, rx, ry, l0, l1);
            const Lam lam = lam_sample(rng_float(rng));
            W.ox[c][slot] = r.o.x; W.oy[c][slot] = r.o.y; W.oz[c][slot] = r.o.z; W.dx[c][slot] = r.d.x; W.dy[c][slot] = r.d.y; W.dz[c][slot] = r.d.z;
            for (int k = 0; k < 4; k++) { W.lam[(size_t)k * N + slot] = lam.l[k]; W.gathered[c][(size_t)k * N + slot] = 1.0; W.radiance[(size_t)k * N + slot] = 0.0; }
            W.rx[slot] = rx; W.ry[slot] = ry;
            W.pixel[slot] = pixel; W.sample[slot] = sample; W.depth[slot] = 0u; W.draws[c][slot] = rng.draws; W.witem[slot] = (uint32_t)w;
            f = PF_ALIVE | PF_LAST_SPECULAR;
            break;
        }
        W.flags[slot] = f;
    }
    __syncwarp();
    for (int o = 16; o > 0; o >>= 1) {
        cost_sum += __shfl_down_sync(0xFFFFFFFFu, cost_sum, o); paths += __shfl_down_sync(0xFFFFFFFFu, paths, o);
        depth_max = max(depth_max, __shfl_down_sync(0xFFFFFFFFu, depth_max, o));
    }
    if ((threadIdx.x & 31u) == 0u && paths) { atomicAdd(&W.run->cost, cost_sum); atomicAdd(&W.run->camera_paths, paths); atomicMax(&W.run->max_depth, depth_max); }
}
// stream compaction of the live slots, in slot order within a warp
__global__ void __launch_bounds__(256) k_compact(const __grid_constant__ Wave W) {
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < W.n_slots; slot += gridDim.x * blockDim.x)
        if (W.flags[slot] & PF_ALIVE) W.active[agg_inc(&W.it->n_active)] = slot;
}

// ---- trace: Scene::hit over the active queue, then binning by what the shading stage has to do --------
template <bool CNT>
__global__ void __launch_bounds__(128, LUMO_WAVE_TRACE_BLOCKS) k_wave_trace(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, uint32_t cur, Counters* gc) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = W.it->n_active;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&W.it->trace_next, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + lane;
        uint32_t slot = 0, klass = LUMO_N_CLASSES;
        if (i < n) {
            slot = W.active[i];
            Ray r; r.o = d3(W.ox[cur][slot], W.oy[cur][slot], W.oz[cur][slot]); r.d = d3(W.dx[cur][slot], W.dy[cur][slot], W.dz[cur][slot]);
            HitRec h;
            klass = 0;
            if (scene_hit<CNT, LUMO_WAVE_KD_ROUND>(S, r, LUMO_INF, h, &cnt)) {
                W.ht[slot] = h.t; W.hb0[slot] = h.bary.x; W.hb1[slot] = h.bary.y; W.hb2[slot] = h.bary.z; W.hobj[slot] = h.obj; W.htri[slot] = h.tri;
                const uint32_t kind = S.materials[S.objects[h.obj].material].kind;
                if (kind >= LMAT_LAMBERTIAN && kind <= LMAT_MFDIELECTRIC) klass = kind;
            } else W.hobj[slot] = LUMO_NONE;
        }
        // stream compaction into the per-class shade queues (one atomic per class per warp)
#pragma unroll
        for (uint32_t c = 0; c < LUMO_N_CLASSES; c++) {
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, klass == c);
            if (m) {
                uint32_t b = 0;
                if (lane == (uint32_t)(__ffs(m) - 1)) b = atomicAdd(&W.it->n_class[c], (uint32_t)__popc(m));
                b = __shfl_sync(0xFFFFFFFFu, b, __ffs(m) - 1);
                if (klass == c) W.cls[c][b + __popc(m & ((1u << lane) - 1u))] = slot;
            }
        }
    }
    if (lane == 0 && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&W.run->closest, (unsigned long long)n);
    if (CNT) { atomicAdd(&gc->tlas, cnt.tlas); atomicAdd(&gc->inst, cnt.inst); atomicAdd(&gc->kd, cnt.kd); atomicAdd(&gc->leaf, cnt.leaf); atomicAdd(&gc->tri, cnt.tri); atomicAdd(&gc->sphere, cnt.sphere); }
}

// ---- the same through the closest-hit pipeline of closest.cuh: ray source, hit sink, class binning --------------------------------
struct WaveRaySource {
    Wave W; uint32_t cur;
    __device__ __forceinline__ uint32_t n() const { return W.it->n_active; }
    __device__ __forceinline__ void load(uint32_t i, Ray& r, double& t_max) const {
        const uint32_t slot = W.active[i];
        r.o = d3(W.ox[cur][slot], W.oy[cur][slot], W.oz[cur][slot]); r.d = d3(W.dx[cur][slot], W.dy[cur][slot], W.dz[cur][slot]);
        t_max = LUMO_INF;
    }
};
struct WaveHitSink {
    Wave W;
    __device__ __forceinline__ void store(uint32_t i, bool have, const HitRec& h) const {
        const uint32_t slot = W.active[i];
        if (have) { W.ht[slot] = h.t; W.hb0[slot] = h.bary.x; W.hb1[slot] = h.bary.y; W.hb2[slot] = h.bary.z; W.hobj[slot] = h.obj; W.htri[slot] = h.tri; }
        else W.hobj[slot] = LUMO_NONE;
    }
};
// stream compaction of the traced slots into the per-class shade queues, in active-queue order (what k_wave_trace does at its end)
__global__ void __launch_bounds__(256) k_wave_classify(const __grid_constant__ DevScene S, const __grid_constant__ Wave W) {
    const uint32_t n = W.it->n_active, lane = threadIdx.x & 31u;
    const uint32_t n_pad = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
        uint32_t slot = 0, klass = LUMO_N_CLASSES;
        if (i < n) {
            slot = W.active[i]; klass = 0;
            const uint32_t obj = W.hobj[slot];
            if (obj != LUMO_NONE) { const uint32_t kind = S.materials[S.objects[obj].material].kind; if (kind >= LMAT_LAMBERTIAN && kind <= LMAT_MFDIELECTRIC) klass = kind; }
        }
#pragma unroll
        for (uint32_t c = 0; c < LUMO_N_CLASSES; c++) {
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, klass == c);
            if (m) {
                uint32_t b = 0;
                if (lane == (uint32_t)(__ffs(m) - 1)) b = atomicAdd(&W.it->n_class[c], (uint32_t)__popc(m));
                b = __shfl_sync(0xFFFFFFFFu, b, __ffs(m) - 1);
                if (klass == c) W.cls[c][b + __popc(m & ((1u << lane) - 1u))] = slot;
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&W.run->closest, (unsigned long long)n);
}
// caller-supplied ray batches (lumo_gpu_trace_closest)
struct BatchRaySource {
    const double *o, *d, *t_max; uint32_t count;
    __device__ __forceinline__ uint32_t n() const { return count; }
    __device__ __forceinline__ void load(uint32_t i, Ray& r, double& tm) const {
        r.o = d3(o[3 * (size_t)i], o[3 * (size_t)i + 1], o[3 * (size_t)i + 2]); r.d = d3(d[3 * (size_t)i], d[3 * (size_t)i + 1], d[3 * (size_t)i + 2]); tm = t_max ? t_max[i] : LUMO_INF;
    }
};
struct BatchHitSink {
    uint32_t *obj, *tri; double *t, *bary;
    __device__ __forceinline__ void store(uint32_t i, bool have, const HitRec& h) const {
        if (have) { obj[i] = h.obj; tri[i] = h.tri; t[i] = h.t; bary[2 * (size_t)i] = h.bary.x; bary[2 * (size_t)i + 1] = h.bary.y; }
        else { obj[i] = LUMO_NONE; tri[i] = LUMO_NONE; t[i] = LUMO_INF; bary[2 * (size_t)i] = 0.0; bary[2 * (size_t)i + 1] = 0.0; }
    }
};

// ---- occlude: the occlusion half of Scene::hit_light over the shadow queue ------------------------
template <bool CNT>
__global__ void __launch_bounds__(128, LUMO_WAVE_TRACE_BLOCKS) k_wave_occlude(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, Counters* gc,
                                                                              const uint8_t* expect, AhCounters* ah) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = min(W.it->n_shadow, W.shadow_cap);
    const uint32_t N = W.n_slots, C = W.shadow_cap;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    unsigned long long mismatches = 0;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&W.it->occl_next, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + lane;
        if (i < n) {
            Ray r; r.o = d3(W.sox[i], W.soy[i], W.soz[i]); r.d = d3(W.sdx[i], W.sdy[i], W.sdz[i]);
            const bool occluded = scene_occluded<CNT, LUMO_WAVE_KD_ROUND>(S, r, W.stmax[i], &cnt);
            if (expect && (expect[i] != 0) != occluded) mismatches++;       // LUMO_OCCLUDE_CHECK: the occlusion BVH's verdict for the same ray
            if (!occluded) {
                const uint32_t slot = W.sslot[i];
                for (int k = 0; k < 4; k++) { const double v = W.sc[(size_t)k * C + i]; if (v != 0.0) atomicAdd(&W.radiance[(size_t)k * N + slot], v); }
            }
        }
    }
    if (lane == 0 && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&W.run->occlusion, (unsigned long long)n);
    if (CNT) { atomicAdd(&gc->tlas, cnt.tlas); atomicAdd(&gc->inst, cnt.inst); atomicAdd(&gc->kd, cnt.kd); atomicAdd(&gc->leaf, cnt.leaf); atomicAdd(&gc->tri, cnt.tri); atomicAdd(&gc->sphere, cnt.sphere); }
    if (expect && mismatches) atomicAdd(&ah->mismatches, mismatches);
}
// the shadow queue as a ray source / verdict sink of the occlusion-BVH kernels (occlude.cuh)
struct WaveShadowSource {
    Wave W;
    __device__ __forceinline__ uint32_t n() const { return min(W.it->n_shadow, W.shadow_cap); }
    __device__ __forceinline__ void load(uint32_t i, Ray& r, double& t_max) const {
        r.o = d3(W.sox[i], W.soy[i], W.soz[i]); r.d = d3(W.sdx[i], W.sdy[i], W.sdz[i]); t_max = W.stmax[i];
    }
};
// Verdicts are only noted here; the contributions of the unoccluded rays are added by the kernel that follows — k_shadow_apply, or
// the faithful k_wave_occlude in check mode.  (Adding them from inside the walk made every lane of a warp wait for one lane's
// cold loads of the colour and its four atomics: 12 % of k_occl_bvh's stall samples.)
struct WaveShadowSink {
    Wave W; uint8_t* record;
    __device__ __forceinline__ void verdict(uint32_t i, bool occluded) const { record[i] = occluded ? 1 : 0; }
    __device__ __forceinline__ void done(uint32_t) const {}
};
__global__ void __launch_bounds__(256) k_shadow_apply(const __grid_constant__ Wave W, const uint8_t* __restrict__ record) {
    const uint32_t n = min(W.it->n_shadow, W.shadow_cap), N = W.n_slots, C = W.shadow_cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (record[i]) continue;
        const uint32_t slot = W.sslot[i];
        for (int k = 0; k < 4; k++) { const double v = W.sc[(size_t)k * C + i]; if (v != 0.0) atomicAdd(&W.radiance[(size_t)k * N + slot], v); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&W.run->occlusion, (unsigned long long)n);
}
// caller-supplied ray batches (lumo_gpu_trace_any)
struct BatchShadowSource {
    const double *o, *d, *t_max; uint32_t count;
    __device__ __forceinline__ uint32_t n() const { return count; }
    __device__ __forceinline__ void load(uint32_t i, Ray& r, double& tm) const {
        r.o = d3(o[3 * (size_t)i], o[3 * (size_t)i + 1], o[3 * (size_t)i + 2]); r.d = d3(d[3 * (size_t)i], d[3 * (size_t)i + 1], d[3 * (size_t)i + 2]); tm = t_max[i];
    }
};
struct BatchShadowSink {
    uint8_t* occ;
    __device__ __forceinline__ void verdict(uint32_t i, bool occluded) const { occ[i] = occluded ? 1 : 0; }
    __device__ __forceinline__ void done(uint32_t) const {}
};

// ---- shade ------------------------------------------------------------------------------------------
// The shading stage of one iteration is three kinds of kernels over the per-class queues:
//   k_terminal   class 0: the ray missed, or hit a Light / Blank material (bsdf_sample -> None):
//                emission if the previous bounce was specular (path_trace.rs:24-29), path ends;
//   k_scatter<K> class K: the BSDF sample of this bounce, throughput update, Russian roulette, next ray
//                (path_trace.rs:23,42-78; direct_light.rs:23-68) — decides whether NEE runs;
//   k_nee_a / k_nee_b<K> / k_nee_eval<K>   next-event estimation, integrator.rs:74-184 (see below); the occlusion half of
//                hit_light goes to the shadow queue.
// One material family per kernel keeps warps on one code path and the instruction footprint small.
__device__ __forceinline__ void push_shadow(const Wave& W, uint32_t slot, const Ray& r, double t_max, const C4& c) {
    const uint32_t i = agg_inc(&W.it->n_shadow);
    if (i >= W.shadow_cap) { atomicAdd(&W.run->shadow_dropped, 1ull); return; }
    W.sox[i] = r.o.x; W.soy[i] = r.o.y; W.soz[i] = r.o.z; W.sdx[i] = r.d.x; W.sdy[i] = r.d.y; W.sdz[i] = r.d.z;
    W.stmax[i] = t_max; W.sslot[i] = slot;
    for (int k = 0; k < 4; k++) W.sc[(size_t)k * W.shadow_cap + i] = c.s[k];
}
__device__ __forceinline__ void load_path(const Wave& W, uint32_t cur, uint32_t slot, Ray& ro, HitRec& rec) {
    ro.o = d3(W.ox[cur][slot], W.oy[cur][slot], W.oz[cur][slot]); ro.d = d3(W.dx[cur][slot], W.dy[cur][slot], W.dz[cur][slot]);
    rec.t = W.ht[slot]; rec.bary = d3(W.hb0[slot], W.hb1[slot], W.hb2[slot]); rec.obj = W.hobj[slot]; rec.tri = W.htri[slot];
}
// dispersion (bxdf/microfacet.rs:282-286): a dielectric with a non-constant eta keeps only the hero wavelength
template <int K> __device__ __forceinline__ void maybe_terminate(const Mat& m, Lam& l) {
    if ((K & 7) == LMAT_MFDIELECTRIC && !(m.flags & LMF_ETA_CONST)) { l.l[1] = 0.0; l.l[2] = 0.0; l.l[3] = 0.0; }
}

__global__ void __launch_bounds__(128) k_terminal(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P) {
    const uint32_t N = W.n_slots, n = W.it->n_class[0];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = W.cls[0][i];
        const uint32_t f = W.flags[slot];
        if (W.hobj[slot] != LUMO_NONE && (P.integrator == 1 || (f & PF_LAST_SPECULAR))) {
            Ray ro; HitRec rec; load_path(W, P.cur, slot, ro, rec);
            const DevHit ho = reconstruct_hit(S, ro, rec);
            const Mat& m = S.materials[ho.material];
            Lam lam; for (int k = 0; k < 4; k++) lam.l[k] = W.lam[(size_t)k * N + slot];
            const C4 e = load_c4(W.gathered[P.cur], N, slot) * mat_emit(S, m, lam, ho);
            if (!is_black(e)) { C4 rad = load_c4(W.radiance, N, slot); rad = rad + e; store_c4(W.radiance, N, slot, rad); }
        }
        W.flags[slot] = PF_DONE;
        W.done[P.cur][agg_inc(&W.qc->n_done[P.cur])] = slot;
    }
}

// resident CTAs per SM the shade kernels are compiled for (register budget = 65536 / (128 * blocks))
#ifndef LUMO_SCATTER_BLOCKS
#define LUMO_SCATTER_BLOCKS 4
#endif
#ifndef LUMO_NEE_BLOCKS
#define LUMO_NEE_BLOCKS 4
#endif
__device__ __forceinline__ void nee_ctx_store(const Wave& W, uint32_t slot, const DevHit& h, D3 nb);
template <int K>
__global__ void __launch_bounds__(128, LUMO_SCATTER_BLOCKS) k_scatter(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P) {
    const uint32_t N = W.n_slots, n = W.it->n_class[K & 7], cur = P.cur, nxt = P.cur ^ 1u;
    const uint32_t n_pad = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
        __syncwarp();                                 // keep the warp's lanes on the same item index (see k_nee)
        if (i >= n) continue;
        const uint32_t slot = W.cls[K & 7][i];
        Ray ro; HitRec rec; load_path(W, cur, slot, ro, rec);
        const DevHit ho = reconstruct_hit(S, ro, rec);
        const Mat& m = S.materials[ho.material];
        const Onb uvw = shading_onb<K>(S, m, ho);
        const uint32_t pixel = W.pixel[slot];
        Rng rng = rng_make(P.seed, pixel, W.sample[slot], 0u, W.draws[cur][slot]);
        Lam lam; for (int k = 0; k < 4; k++) lam.l[k] = W.lam[(size_t)k * N + slot];
        C4 gathered = load_c4(W.gathered[cur], N, slot);
        uint32_t depth = W.depth[slot];
        const D3 wo = -ro.d;
        const bool dbg = pixel == P.debug_pixel && P.mode == WM_MAIN;
        if (dbg) printf("[gpu] depth=%u obj=%u tri=%u t=%.17g g0=%.17g rad0=%.17g\n", depth, rec.obj, rec.tri, rec.t, gathered.s[0], W.radiance[slot]);
        const double ru = rng_float(rng), r0 = rng_float(rng), r1 = rng_float(rng);
        D3 wi;
        uint32_t f = 0;
        bool done = false, nee = false;
        if (!bsdf_sample<K>(S, m, uvw, wo, ho, lam, ru, r0, r1, wi)) done = true;   // a Standard material emits nothing (material.rs:220-231)
        else if (P.integrator == 1) {
            if (!mat_is_specular(m)) { nee = true; done = true; }
            else if (depth >= LUMO_DL_MAX_RECURSION) done = true;
        } else nee = !mat_is_delta(S, m, lam);
        if (nee) { rng.draws += 6u * S.P.n_shadow_rays; nee_ctx_store(W, slot, ho, uvw.w); }   // the draws the NEE stages consume (integrator.rs:96-118); their view of this hit
        if (!done) {
            const Ray ri = hit_generate_ray(ho, wi);
            wi = ri.d;
            const double p_scatter = bsdf_pdf<K>(S, m, uvw, wo, wi, ho, lam, false);
            if (p_scatter <= 0.0) done = true;
            else {
                const C4 bsdf = bsdf_f<K>(S, m, uvw, wo, wi, lam, 0, ho);
                gathered = gathered * (bsdf * shading_cosine(m, wi, ho.ns) / p_scatter);
                if (P.integrator == 0 && depth >= LUMO_RR_DEPTH) {
                    const double delta = W.tile_delta[tile_of(P, S, pixel)];
                    const double lum = luminance(S, gathered, lam);
                    const double rr = fmin(lum / delta, 1.0);
                    rng = rng_make(P.seed, pixel, W.sample[slot], 0u, rng.draws);
                    if (rng_float(rng) > rr) done = true;
                    else gathered = gathered / rr;
                }
                if (!done) {
                    f = PF_ALIVE | (mat_is_specular(m) ? PF_LAST_SPECULAR : 0u);
                    W.ox[nxt][slot] = ri.o.x; W.oy[nxt][slot] = ri.o.y; W.oz[nxt][slot] = ri.o.z; W.dx[nxt][slot] = ri.d.x; W.dy[nxt][slot] = ri.d.y; W.dz[nxt][slot] = ri.d.z;
                    store_c4(W.gathered[nxt], N, slot, gathered);
                    W.depth[slot] = depth + 1u;
                    W.draws[nxt][slot] = rng.draws;
                }
            }
        }
        if ((K & 7) == LMAT_MFDIELECTRIC) for (int k = 1; k < 4; k++) W.lam[(size_t)k * N + slot] = lam.l[k];
        W.flags[slot] = (done ? PF_DONE : f) | (nee ? PF_NEE : 0u);
        if (nee) W.neeq[agg_inc(&W.it->n_nee)] = slot;
        if (done) W.done[cur][agg_inc(&W.qc->n_done[cur])] = slot;
        else agg_inc(&W.qc->n_active[nxt]);
    }
}

// ---- next-event estimation (integrator.rs:74-184) in three dense stages ---------------------------------------------------
// A bounce that runs NEE evaluates n_shadow_rays shadow samples (1 for up to 3 lights, log2 #lights beyond: 12 for a street
// with 4097 lights), each with a light-sampled term (A) and a BSDF-sampled term (B).  As one kernel this was the largest
// and the worst-running piece of the pipeline: 16 k SASS instructions (37 % of the stall samples waiting for instruction
// fetch), 128 registers, and terms that end early (A below the horizon, B missing the light) thinning the warps.  Now:
//   k_scatter<K>   (already there) also leaves the bounce's shading context — reconstructed Hit, shading normals — in HBM
//   k_nee_a        one thread per (path, sample): light pick, point on the light, the two sign tests, the light's own
//                  intersection test, its pdf and emission                  -> term queue      (no material code at all)
//   k_nee_b<K>     one thread per (path, sample): BSDF sample of material family K, the chosen light's intersection
//                  test, its pdf and emission                               -> term queue      (only BxDF::sample of K)
//   k_nee_eval<K>  one thread per surviving term: BSDF pdf and value, MIS weight, contribution -> shadow queue
// Every function is called with the arguments the single kernel passed, so the arithmetic — and the film — is unchanged.
#define LUMO_NEE_CTX_DOUBLES 17   /* p, fp_error, ng, ns, shading normal after the normal map, u, v */
__device__ __forceinline__ void nee_ctx_store(const Wave& W, uint32_t slot, const DevHit& h, D3 nb) {
    const size_t N = W.n_slots; double* c = W.nee_ctx + slot;
    c[0] = h.p.x; c[N] = h.p.y; c[2 * N] = h.p.z; c[3 * N] = h.fp_error.x; c[4 * N] = h.fp_error.y; c[5 * N] = h.fp_error.z;
    c[6 * N] = h.ng.x; c[7 * N] = h.ng.y; c[8 * N] = h.ng.z; c[9 * N] = h.ns.x; c[10 * N] = h.ns.y; c[11 * N] = h.ns.z;
    c[12 * N] = nb.x; c[13 * N] = nb.y; c[14 * N] = nb.z; c[15 * N] = h.u; c[16 * N] = h.v;
    W.nee_meta[slot] = (uint32_t)h.material | (h.backface ? 0x80000000u : 0u);
}
__device__ __forceinline__ void nee_ctx_load(const Wave& W, uint32_t slot, DevHit& h, D3& nb) {
    const size_t N = W.n_slots; const double* c = W.nee_ctx + slot;
    h.t = 0.0;
    h.p = d3(c[0], c[N], c[2 * N]); h.fp_error = d3(c[3 * N], c[4 * N], c[5 * N]);
    h.ng = d3(c[6 * N], c[7 * N], c[8 * N]); h.ns = d3(c[9 * N], c[10 * N], c[11 * N]);
    nb = d3(c[12 * N], c[13 * N], c[14 * N]); h.u = c[15 * N]; h.v = c[16 * N];
    const uint32_t m = W.nee_meta[slot];
    h.material = (int)(m & 0x7FFFFFFFu); h.backface = (m & 0x80000000u) != 0u;
}
__device__ __forceinline__ void push_term(const Wave& W, uint32_t slot, bool term_b, const Ray& r, D3 wi, double t_max, double p_lig, double pdf_light, const C4& le) {
    const uint32_t i = agg_inc(&W.it->n_terms);
    if (i >= W.shadow_cap) { atomicAdd(&W.run->shadow_dropped, 1ull); return; }
    const NeeTermQueue& T = W.tq;
    T.ox[i] = r.o.x; T.oy[i] = r.o.y; T.oz[i] = r.o.z; T.dx[i] = r.d.x; T.dy[i] = r.d.y; T.dz[i] = r.d.z; T.wx[i] = wi.x; T.wy[i] = wi.y; T.wz[i] = wi.z;
    T.tmax[i] = t_max; T.p_lig[i] = p_lig; T.pdf_light[i] = pdf_light; T.slot[i] = slot | (term_b ? 0x80000000u : 0u);
    for (int k = 0; k < 4; k++) T.le[(size_t)k * W.shadow_cap + i] = le.s[k];
}
// (path, shadow sample) of a thread: 32 consecutive entries of the family's NEE queue share a warp and a sample index
__device__ __forceinline__ bool nee_item(const Wave& W, unsigned long long it, uint32_t ns, uint32_t nq, uint32_t& slot, uint32_t& i) {
    const unsigned long long grp = it / (32ull * ns);
    i = (uint32_t)((it / 32ull) % ns);
    const uint32_t qi = (uint32_t)(grp * 32ull + (it % 32ull));
    if (qi >= nq) return false;
    slot = W.neeq[qi];
    return true;
}
// (same-box A/B, bistro 4 spp, shade class ms: 4 CTAs per SM 245.6, 5: 250.6, 6: 265.9)
#ifndef LUMO_NEE_A_BLOCKS
#define LUMO_NEE_A_BLOCKS 4
#endif
// the light-sampled term, up to the point where the material comes in (integrator.rs:96-110).
// (Measured and dropped: collecting the survivors of the two sign tests per warp in shared memory and running the second half
// on 32 of them at a time.  Lanes per instruction went from 12 to 19 and the shade class from 209.0 to 217.6 ms on bistro
// 4 spp — the kernel waits on instruction fetch (39 % of its stall samples) and f64 latency, not on issue slots.)
#ifndef LUMO_NEE_B_PREREJECT
#define LUMO_NEE_B_PREREJECT 1
#endif
// LUMO_NEE_A_SPLIT: the light side of both terms is one kernel of its own.  k_nee_a1 picks the light, samples the point and runs the
// two sign tests of the light-sampled term; k_nee_b<K> samples the BSDF and tests the line against the light's bounding sphere; the
// survivors of both (half of the A items, a tenth of the B items) go through a queue in HBM to k_nee_a, which runs the light's
// intersection test, pdf and emission on dense warps.  Each kernel carries a fraction of the code.
// (A further split of k_nee_a — the intersection test | the light's pdf and emission, through the term queue — was neutral to
// slightly worse: shade 175.9 -> 178.7 ms on bistro 4 spp.)
#ifndef LUMO_NEE_A_SPLIT
#define LUMO_NEE_A_SPLIT 1
#endif
#if LUMO_NEE_A_SPLIT
__global__ void __launch_bounds__(128, LUMO_NEE_A_BLOCKS) k_nee_a1(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P, uint32_t klass) {
    const uint32_t N = W.n_slots, nq = W.it->n_nee, cur = P.cur, ns = S.P.n_shadow_rays;
    const unsigned long long padded = (unsigned long long)((nq + 31u) / 32u) * 32ull * ns;
    for (unsigned long long it = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; it < padded; it += (unsigned long long)gridDim.x * blockDim.x) {
        __syncwarp();
        uint32_t slot, i;
        if (!nee_item(W, it, ns, nq, slot, i)) continue;
        const double* c = W.nee_ctx + slot;
        const D3 xo = d3(c[0], c[(size_t)N], c[2 * (size_t)N]);
        // draws of this bounce: 3 for the scatter sample, then 6 per shadow sample: light pick, light point (2), BSDF sample (3)
        Rng rng = rng_make(P.seed, W.pixel[slot], W.sample[slot], 0u, W.draws[cur][slot] + 3u + 6u * i);
        const uint32_t li = sample_light(S, rng_float(rng));
        const LumoObject lo = S.objects[S.P.n_objects + li];
        const double r0 = rng_float(rng), r1 = rng_float(rng);
        const D3 wi = light_sample_towards(S, lo, xo, r0, r1);
        // mis_sample returns black when either pdf is zero (integrator.rs:150-152) whatever the light test says.  For the
        // reflection-only BxDFs the pdf starts with two sign tests (bsdf.rs:88-90, bxdf.rs:136-139, scatter.rs:14-17):
        // a light-sampled direction below the surface stops here, before any light or microfacet arithmetic.
        if (klass != LMAT_MFDIELECTRIC) {
            const D3 ng = d3(c[6 * (size_t)N], c[7 * (size_t)N], c[8 * (size_t)N]), nb = d3(c[12 * (size_t)N], c[13 * (size_t)N], c[14 * (size_t)N]);
            const D3 wo = d3(-W.dx[cur][slot], -W.dy[cur][slot], -W.dz[cur][slot]);
            if (!is_reflection(wo, wi, ng)) continue;
            const Onb uvw = onb_new(nb);
            if (!same_hemisphere(to_local(uvw, wo), to_local(uvw, wi))) continue;
        }
        const uint32_t j = agg_inc(&W.it->n_surv);
        if (j >= W.shadow_cap) { atomicAdd(&W.run->shadow_dropped, 1ull); continue; }     // cannot happen unless shadow_cap was clamped (render fails loudly)
        W.surv.slot[j] = slot; W.surv.li[j] = li; W.surv.wx[j] = wi.x; W.surv.wy[j] = wi.y; W.surv.wz[j] = wi.z;
    }
}
template <bool TEX>
__global__ void __launch_bounds__(128, LUMO_NEE_A_BLOCKS) k_nee_a(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P, uint32_t klass) {
    const uint32_t N = W.n_slots, n = min(W.it->n_surv, W.shadow_cap);
    const uint32_t n_pad = (n + 31u) & ~31u;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_pad; j += gridDim.x * blockDim.x) {
        __syncwarp();
        if (j >= n) continue;
        const uint32_t slot = W.surv.slot[j], lf = W.surv.li[j], li = lf & 0x00FFFFFFu, lobj = S.P.n_objects + li;   // n_lights < 2^24 (validate_blob)
        const bool term_b = (lf & 0x80000000u) != 0u, far_off = (lf & 0x40000000u) != 0u;
        const D3 wi = d3(W.surv.wx[j], W.surv.wy[j], W.surv.wz[j]);
        DevHit ho; D3 nb; nee_ctx_load(W, slot, ho, nb);
        const Ray ri = hit_generate_ray(ho, wi);
        DevHit hi;
        if (!light_hit<TEX>(S, lobj, ri, hi)) continue;
        if (far_off) atomicAdd(&W.run->prereject_bad, 1u);                            // check mode only: must never happen
        const double p_lig = light_sample_towards_pdf(S, S.objects[lobj], ri, hi.p, hi.ng);
        Lam lam; for (int k = 0; k < 4; k++) lam.l[k] = W.lam[(size_t)k * N + slot];   // already terminated by k_scatter if dispersive
        const C4 le = mat_emit<TEX ? -1 : LUMO_K_SOLID>(S, S.materials[hi.material], lam, hi);
        push_term(W, slot, term_b, ri, wi, hi.t - LUMO_EPS, p_lig, S.lights[li].pdf, le);
    }
}
#else
template <bool TEX>
__global__ void __launch_bounds__(128, LUMO_NEE_A_BLOCKS) k_nee_a(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P, uint32_t klass) {
    const uint32_t N = W.n_slots, nq = W.it->n_nee, cur = P.cur, ns = S.P.n_shadow_rays;
    const unsigned long long padded = (unsigned long long)((nq + 31u) / 32u) * 32ull * ns;
    for (unsigned long long it = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; it < padded; it += (unsigned long long)gridDim.x * blockDim.x) {
        __syncwarp();                                 // a lane whose item ended early waits here instead of running ahead into its next one
        uint32_t slot, i;
        if (!nee_item(W, it, ns, nq, slot, i)) continue;
        DevHit ho; D3 nb; nee_ctx_load(W, slot, ho, nb);
        // draws of this bounce: 3 for the scatter sample, then 6 per shadow sample: light pick, light point (2), BSDF sample (3)
        Rng rng = rng_make(P.seed, W.pixel[slot], W.sample[slot], 0u, W.draws[cur][slot] + 3u + 6u * i);
        const uint32_t li = sample_light(S, rng_float(rng));
        const uint32_t lobj = S.P.n_objects + li;
        const LumoObject lo = S.objects[lobj];
        const double r0 = rng_float(rng), r1 = rng_float(rng);
        const D3 wi = light_sample_towards(S, lo, ho.p, r0, r1);
        // mis_sample returns black when either pdf is zero (integrator.rs:150-152) whatever the light test says.  For the
        // reflection-only BxDFs the pdf starts with two sign tests (bsdf.rs:88-90, bxdf.rs:136-139, scatter.rs:14-17):
        // a light-sampled direction below the surface stops here, before any light or microfacet arithmetic.
        if (klass != LMAT_MFDIELECTRIC) {
            const D3 wo = d3(-W.dx[cur][slot], -W.dy[cur][slot], -W.dz[cur][slot]);
            if (!is_reflection(wo, wi, ho.ng)) continue;
            const Onb uvw = onb_new(nb);
            if (!same_hemisphere(to_local(uvw, wo), to_local(uvw, wi))) continue;
        }
        const Ray ri = hit_generate_ray(ho, wi);
        DevHit hi;
        if (!light_hit<TEX>(S, lobj, ri, hi)) continue;
        const double p_lig = light_sample_towards_pdf(S, lo, ri, hi.p, hi.ng);
        Lam lam; for (int k = 0; k < 4; k++) lam.l[k] = W.lam[(size_t)k * N + slot];   // already terminated by k_scatter if dispersive
        const C4 le = mat_emit<TEX ? -1 : LUMO_K_SOLID>(S, S.materials[hi.material], lam, hi);
        push_term(W, slot, false, ri, wi, hi.t - LUMO_EPS, p_lig, S.lights[li].pdf, le);
    }
}
#endif
// the BSDF-sampled term up to the same point (integrator.rs:112-134)
template <int K>
__global__ void __launch_bounds__(128, LUMO_NEE_A_BLOCKS) k_nee_b(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P) {
    const uint32_t N = W.n_slots, nq = W.it->n_nee, cur = P.cur, ns = S.P.n_shadow_rays;
    const unsigned long long padded = (unsigned long long)((nq + 31u) / 32u) * 32ull * ns;
    for (unsigned long long it = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; it < padded; it += (unsigned long long)gridDim.x * blockDim.x) {
        __syncwarp();
        uint32_t slot, i;
        if (!nee_item(W, it, ns, nq, slot, i)) continue;
        DevHit ho; D3 nb; nee_ctx_load(W, slot, ho, nb);
        const Mat& m = S.materials[ho.material];
        const uint32_t pixel = W.pixel[slot], sample = W.sample[slot], d0 = W.draws[cur][slot] + 3u + 6u * i;
        Rng rng = rng_make(P.seed, pixel, sample, 0u, d0);
        const uint32_t li = sample_light(S, rng_float(rng));
#if !LUMO_NEE_A_SPLIT
        const uint32_t lobj = S.P.n_objects + li;
#endif
        rng = rng_make(P.seed, pixel, sample, 0u, d0 + 3u);
        const double ru = rng_float(rng), r0 = rng_float(rng), r1 = rng_float(rng);
        Lam lam; for (int k = 0; k < 4; k++) lam.l[k] = W.lam[(size_t)k * N + slot];
        const Onb uvw = onb_new(nb);
        const D3 wo = d3(-W.dx[cur][slot], -W.dy[cur][slot], -W.dz[cur][slot]);
        D3 wi;
        Lam l2 = lam;
        if (!bsdf_sample<K>(S, m, uvw, wo, ho, l2, ru, r0, r1, wi)) continue;
        const LumoLight L = S.lights[li];
#if LUMO_NEE_B_PREREJECT
        // The BSDF-sampled direction nearly never points at the ONE light this shadow sample picked (1 of 4097 on the street):
        // a line that stays outside the light's padded bounding sphere cannot pass its intersection test (scene_blob.h LumoLight).
        // Tested on the surface point itself, before the spawn offset and the normalisation of hit_generate_ray: the offset is
        // at most |fp_error|_1 plus an ulp (hit.rs:85-111), added to the radius; the 1e-12 |v|^2 covers the direction's rounding.
        bool far_off;
        {
            const D3 v = d3(L.bound_c[0], L.bound_c[1], L.bound_c[2]) - ho.p;
            const double pad = 2.0 * (fabs(ho.fp_error.x) + fabs(ho.fp_error.y) + fabs(ho.fp_error.z)) + 1e-14 * (fabs(ho.p.x) + fabs(ho.p.y) + fabs(ho.p.z));
            const double R = L.bound_r + pad;
            const double vv = dot(v, v), dd = dot(wi, wi), b = dot(v, wi);
            far_off = vv * dd - b * b > (R * R + 1e-12 * vv) * dd;
        }
        if (far_off && !P.check) continue;
#else
        const bool far_off = false;
#endif
#if LUMO_NEE_A_SPLIT
        // the light's intersection test, pdf and emission are k_nee_a's (the second stage of both terms): bit 31 marks the BSDF-sampled term
        { const uint32_t j = agg_inc(&W.it->n_surv);
          if (j >= W.shadow_cap) { atomicAdd(&W.run->shadow_dropped, 1ull); continue; }
          W.surv.slot[j] = slot; W.surv.li[j] = li | 0x80000000u | (far_off ? 0x40000000u : 0u); W.surv.wx[j] = wi.x; W.surv.wy[j] = wi.y; W.surv.wz[j] = wi.z; }
#else
        const Ray ri = hit_generate_ray(ho, wi);
        DevHit hi;
        if (!light_hit<LUMO_TEX(K)>(S, lobj, ri, hi)) continue;
        if (far_off) atomicAdd(&W.run->prereject_bad, 1u);                            // check mode only: must never happen
        const double p_lig = light_sample_towards_pdf(S, S.objects[lobj], ri, hi.p, hi.ng);
        const C4 le = mat_emit<K>(S, S.materials[hi.material], lam, hi);
        push_term(W, slot, true, ri, wi, hi.t - LUMO_EPS, p_lig, L.pdf, le);
#endif
    }
}
// BSDF pdf and value, MIS weight (integrator.rs:139-184), contribution -> shadow queue.  Dense over the term queue.
template <int K>
__global__ void __launch_bounds__(128, LUMO_NEE_BLOCKS) k_nee_eval(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P) {
    const uint32_t N = W.n_slots, cur = P.cur, C = W.shadow_cap;
    const uint32_t n = min(W.it->n_terms, C);
    const uint32_t n_pad = (n + 31u) & ~31u;
    const double ns = (double)S.P.n_shadow_rays;
    if (blockIdx.x == 0 && threadIdx.x == 0 && P.mode == WM_MAIN) { atomicAdd(&W.run->nee_terms, (unsigned long long)n); atomicAdd(&W.run->nee_bounces, (unsigned long long)W.it->n_nee); }
    const NeeTermQueue& T = W.tq;
    for (uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x; ti < n_pad; ti += gridDim.x * blockDim.x) {
        __syncwarp();
        if (ti >= n) continue;
        const uint32_t sl = T.slot[ti], slot = sl & 0x7FFFFFFFu;
        const bool light_sampled = (sl & 0x80000000u) == 0u;
        const double p_lig = T.p_lig[ti];
        DevHit ho; D3 nb; nee_ctx_load(W, slot, ho, nb);
        const Mat& m = S.materials[ho.material];
        const Onb uvw = onb_new(nb);
        const D3 wo = d3(-W.dx[cur][slot], -W.dy[cur][slot], -W.dz[cur][slot]);
        const D3 wi = d3(T.wx[ti], T.wy[ti], T.wz[ti]);
        Lam lam; for (int k = 0; k < 4; k++) lam.l[k] = W.lam[(size_t)k * N + slot];
        const double p_sct = bsdf_pdf<K>(S, m, uvw, wo, wi, ho, lam, false);
        if (p_lig == 0.0 || p_sct == 0.0) continue;                                   // integrator.rs:150-152
        const C4 bsdf = bsdf_f<K>(S, m, uvw, wo, wi, lam, 0, ho);
        const double denom = p_lig * p_lig + p_sct * p_sct;
        const double weight = light_sampled ? (p_lig * p_lig) / denom : (p_sct * p_sct) / denom;
        const double p_denom = light_sampled ? p_lig : p_sct;
        C4 le; for (int k = 0; k < 4; k++) le.s[k] = T.le[(size_t)k * C + ti];
        const C4 c = bsdf * c4(1.0) * le * shading_cosine(m, wi, ho.ns) * weight / p_denom;
        if (W.pixel[slot] == P.debug_pixel && P.mode == WM_MAIN) printf("  [gpu] %c vis t=%.17g p_lig=%.17g p_sct=%.17g c0=%.17g\n", light_sampled ? 'A' : 'B', T.tmax[ti] + LUMO_EPS, p_lig, p_sct, c.s[0]);
        if (is_black(c)) continue;
        Ray ri; ri.o = d3(T.ox[ti], T.oy[ti], T.oz[ti]); ri.d = d3(T.dx[ti], T.dy[ti], T.dz[ti]);
        push_shadow(W, slot, ri, T.tmax[ti], load_c4(W.gathered[cur], N, slot) * (c / T.pdf_light[ti]) / ns);
    }
}
__global__ void k_terms_reset(IterCounters* it) { it->n_terms = 0u; it->n_nee = 0u; it->n_surv = 0u; }   // after each material family

// RR threshold of a tile from its 64 pilot paths, summed in index order (task.rs:42-53 applied to
// the pilot set): var = sum f^2 - (sum f)^2 / n; delta = var <= 0 ? 1e-5 : sqrt(var / sum cost)
__global__ void k_pilot_reduce(const Wave W, uint32_t n_tiles, double* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    double f = 0.0, f2 = 0.0; unsigned long long cost = 0;
    for (uint32_t k = 0; k < LUMO_PILOT_N; k++) { const double l = W.pilot_lum[t * LUMO_PILOT_N + k]; f += l; f2 += l * l; cost += W.pilot_cost[t * LUMO_PILOT_N + k]; }
    const double var = f2 - f * f / (double)LUMO_PILOT_N;
    out[t] = var <= 0.0 ? 1e-5 : sqrt(var / (double)cost);
}
__global__ void k_fill(double* p, size_t n, double v) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v; }

// ---- batch entry points of the C ABI (parity kernels) ----------------------------------------------
// mode 0: Scene::hit; 1: hit_light occlusion; 2: Scene::hit_t
template <int MODE, bool CNT>
__global__ void __launch_bounds__(128) k_trace_batch(const __grid_constant__ DevScene S, const double* __restrict__ o, const double* __restrict__ d,
                                                     const double* __restrict__ t_max, unsigned long long n, unsigned long long* next,
                                                     uint32_t* obj, uint32_t* tri, double* t, double* bary, uint8_t* occ, Counters* gc) {
    const uint32_t lane = threadIdx.x & 31u;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(next, 32ull);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const unsigned long long i = base + lane;
        if (i < n) {
            Ray r; r.o = d3(o[3 * i], o[3 * i + 1], o[3 * i + 2]); r.d = d3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
            if (MODE == 0) {
                HitRec h;
                if (scene_hit<CNT>(S, r, t_max ? t_max[i] : LUMO_INF, h, &cnt)) { obj[i] = h.obj; tri[i] = h.tri; t[i] = h.t; bary[2 * i] = h.bary.x; bary[2 * i + 1] = h.bary.y; }
                else { obj[i] = LUMO_NONE; tri[i] = LUMO_NONE; t[i] = LUMO_INF; bary[2 * i] = 0.0; bary[2 * i + 1] = 0.0; }
            } else if (MODE == 1) occ[i] = scene_occluded<CNT>(S, r, t_max[i], &cnt) ? 1 : 0;
            else t[i] = scene_hit_t<CNT>(S, r, &cnt);
        }
    }
    if (CNT) { atomicAdd(&gc->tlas, cnt.tlas); atomicAdd(&gc->inst, cnt.inst); atomicAdd(&gc->kd, cnt.kd); atomicAdd(&gc->leaf, cnt.leaf); atomicAdd(&gc->tri, cnt.tri); atomicAdd(&gc->sphere, cnt.sphere); }
}

}  // namespace lumo_dev
