/* lumo_gpu.h — C ABI of liblumo_gpu.so, the B200 (sm_100a) implementation of lumo's rendering
 * hot path.  This is the boundary a `lumo-gpu-sys` Rust crate binds (INTEGRATION.md shows the
 * extern "C" block and the Renderer::render wiring).
 *
 * What it replaces in the reference (ekarpp/lumo v0.6.1, paths relative to the crate root):
 *   - the generic executor seam `trait Executor<T,R> { fn exec(&mut self, task: T) -> R }`
 *     (src/pool.rs:6-8) as instantiated by `RenderTaskExecutor` (src/renderer/task.rs:24-81) and
 *     driven by `ThreadPool` from `Renderer::render` (src/renderer.rs:159-244);
 *   - underneath it `Integrator::integrate` (src/tracer/integrator.rs:45-69), `Scene::hit /
 *     hit_t / hit_light` (src/tracer/scene.rs:119-189), `BVH::_hit` (src/tracer/object/bvh.rs:
 *     315-362), `KdTree::_hit` (src/tracer/object/kdtree.rs:101-169) and `Triangle::_hit`
 *     (src/tracer/object/triangle.rs:63-187).
 *
 * Conventions: every function returns 0 on success and a negative lumo_status on failure; the
 * message is available from lumo_gpu_last_error() (thread-local).  Nothing unwinds or aborts.
 * All pointers are HOST pointers unless a name ends in `_dev`; the library copies in and out.
 * A context is not re-entrant; different contexts may be used from different threads.
 */
#ifndef LUMO_GPU_H
#define LUMO_GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lumo_ctx lumo_ctx;     /* one per GPU: stream, scratch queues */
typedef struct lumo_scene lumo_scene; /* an uploaded scene blob (csrc/common/scene_blob.h) */

enum lumo_status {
    LUMO_OK = 0,
    LUMO_ERR_INVALID = -1,   /* bad argument / malformed blob */
    LUMO_ERR_CUDA = -2,      /* CUDA runtime error, no device */
    LUMO_ERR_OOM = -3,
    LUMO_ERR_UNSUPPORTED = -4
};

/* Integrator::{PathTrace, DirectLight, BDPathTrace}  (src/tracer/integrator.rs:17-27) */
enum lumo_integrator { LUMO_PATH_TRACE = 0, LUMO_DIRECT_LIGHT = 1, LUMO_BD_PATH_TRACE = 2 };
/* SamplerType (src/samplers.rs:8-21); Sobol has 1023 points (samplers/sobol_seq.rs:3): total_spp <= 1023 */
enum lumo_sampler { LUMO_SAMPLER_UNIFORM = 0, LUMO_SAMPLER_JITTERED = 1, LUMO_SAMPLER_MULTI_JITTERED = 2, LUMO_SAMPLER_SOBOL = 3 };
/* ToneMap (src/tone_mapping.rs:13-20) */
enum lumo_tone_map { LUMO_TONE_NONE = 0, LUMO_TONE_CLAMP = 1, LUMO_TONE_REINHARD = 2 };

/* What Renderer::{samples,integrator,seed,sampler,tone_map} configure (src/renderer.rs:66-99).
 * [spp_begin, spp_end) of total_spp lets several GPUs (or calls) share one render: the RNG is
 * keyed by (pixel, global sample index), so the union of disjoint ranges equals one full call. */
typedef struct lumo_render_params {
    int32_t integrator;      /* lumo_integrator */
    int32_t sampler;         /* lumo_sampler */
    int32_t tone_map;        /* lumo_tone_map */
    int32_t flags;           /* reserved, 0 */
    double tone_map_arg;     /* Clamp(max) */
    double rr_delta;         /* > 0: fixed Russian-roulette threshold; 0: per-tile pilot estimate */
    uint64_t seed;           /* Renderer::seed */
    uint32_t spp_begin, spp_end, total_spp;
    uint32_t wave_paths;     /* paths in flight per wave; 0 = default */
} lumo_render_params;

/* Film accumulators (src/tracer/film.rs:52-76,92-100): the caller allocates pixels[W*H*4] =
 * (sum r*w, sum g*w, sum b*w, sum w) and splats[W*H*3]; the library OVERWRITES them.  The host
 * then finishes exactly like Film::rgb_image (film.rs:173-193). */
typedef struct lumo_film_accum {
    double* pixels;
    double* splats;
    uint64_t counters[10];   /* [0] camera paths, [1] closest-hit queries (Scene::hit), [2] occlusion /
                                visibility queries, [3] reference-style cost (sum FilmSample.cost,
                                renderer.rs:221), [4] kernels launched, [5] deepest path, [6] wave
                                iterations, [7] samples whose radiance was NaN/inf (tone_mapping.rs:42-56),
                                [8] BDPT subpaths cut because the device ran out of vertex storage (the reference
                                caps a subpath at 1024 vertices, bd_path_trace.rs:7; 0 unless the overflow pool is
                                exhausted), [9] shadow rays on which the occlusion BVH and the reference traversal
                                disagreed, plus NEE light tests that passed outside the light's bounding sphere
                                (occlusion mode 2 only; must be 0) */
    double* tile_deltas;     /* optional [ceil(W/16)*ceil(H/16)]: RR threshold used per 16x16 tile */
    double device_ms;        /* CUDA-event time of the render on the context's stream */
} lumo_film_accum;

int32_t lumo_gpu_device_count(int32_t* n);
int32_t lumo_gpu_ctx_create(int32_t device, lumo_ctx** ctx);
int32_t lumo_gpu_ctx_destroy(lumo_ctx* ctx);
/* Run this context's later calls on the caller's CUDA stream (a cudaStream_t); NULL = the context's own, which is
   non-blocking, i.e. NOT ordered against the legacy default stream — a caller working on stream 0 passes cudaStreamLegacy. */
int32_t lumo_gpu_ctx_set_stream(lumo_ctx* ctx, void* cuda_stream);

/* Upload a scene blob (built once on the host from lumo's own kd-tree / BVH build; layout in
 * csrc/common/scene_blob.h).  The caller keeps ownership of `blob`. */
int32_t lumo_gpu_scene_upload(lumo_ctx* ctx, const void* blob, uint64_t len, lumo_scene** scene);
int32_t lumo_gpu_scene_destroy(lumo_scene* scene);

/* Scene::hit on a ray batch (src/tracer/scene.rs:119-147).  origin_xyz / dir_xyz: [n*3] f64 AoS;
 * directions are used as given (Ray::new normalises — pass normalised directions).  t_max may be
 * NULL (= +inf; the reference always uses +inf).  Outputs: obj_id = index in Scene.objects
 * insertion order, lights offset by objects.len(); tri_id = index in that object's
 * KdTree.objects (0 for spheres / loose triangles); both 0xFFFFFFFF on a miss; t = hit distance
 * (+inf on miss); bary_uv[2i..] = first two barycentrics (edges/det) for triangles, (0,0) for
 * spheres.  Bit-exact against the reference traversal. */
int32_t lumo_gpu_trace_closest(lumo_scene* scene, const double* origin_xyz, const double* dir_xyz, const double* t_max,
                               uint64_t n, uint32_t* obj_id, uint32_t* tri_id, double* t, double* bary_uv);

/* The occlusion test of Scene::hit_light (src/tracer/scene.rs:180-186): occluded[i] = 1 iff some
 * object or light is hit in (0, t_max[i]).  Pass t_max = t_light - 1e-10 as the reference does. */
int32_t lumo_gpu_trace_any(lumo_scene* scene, const double* origin_xyz, const double* dir_xyz, const double* t_max,
                           uint64_t n, uint8_t* occluded);

/* Scene::hit_t (src/tracer/scene.rs:150-162): distance of the FIRST hit found in reference
 * traversal order (not the nearest; BDPT's visible() depends on it), +inf on miss. */
int32_t lumo_gpu_trace_first_found(lumo_scene* scene, const double* origin_xyz, const double* dir_xyz, uint64_t n, double* t);

/* The whole hot path: camera rays -> integrator -> film accumulators, for sample indices
 * [spp_begin, spp_end) of every pixel. */
int32_t lumo_gpu_render(lumo_scene* scene, const lumo_render_params* params, lumo_film_accum* out);

/* Same as lumo_gpu_render but the film accumulators are DEVICE buffers of the caller (pixels_dev
 * [W*H*4], splats_dev [W*H*3], overwritten): the multi-GPU path reduces them with one NCCL
 * reduce over NVLink before the host reads anything. */
int32_t lumo_gpu_render_dev(lumo_scene* scene, const lumo_render_params* params, double* pixels_dev, double* splats_dev,
                            uint64_t* counters10, double* device_ms);

/* The same render on n GPUs of this host from ONE process: scenes[g] is the same blob uploaded through
 * its own context (normally one context per GPU).  [spp_begin, spp_end) is cut into n contiguous ranges, one host
 * thread per GPU drives its wave pipeline, and the film accumulators are summed on scenes[0]'s GPU by one kernel that
 * reads the other GPUs' films over NVLink peer memory (staged by cudaMemcpyPeer where peer access is unavailable), in
 * the fixed order 0..n-1.  `out` is filled as by lumo_gpu_render: counters are sums over GPUs ([5], [6]: maxima),
 * device_ms = the slowest GPU's render + the reduce, tile_deltas are GPU 0's.  It replaces ThreadPool's fan-out /
 * fan-in (src/renderer.rs:166-235) for the multi-GPU case; with one process per GPU use lumo_gpu_render_dev + ncclReduce. */
int32_t lumo_gpu_render_multi(lumo_scene** scenes, int32_t n, const lumo_render_params* params, lumo_film_accum* out);
/* The sample range lumo_gpu_render_multi gives GPU g of n out of [begin, end): contiguous, disjoint, covering, sizes differing
 * by at most one.  Host arithmetic only (no device needed). */
int32_t lumo_gpu_sample_range(int32_t g, int32_t n, uint32_t begin, uint32_t end, uint32_t* g_begin, uint32_t* g_end);

/* Film finalisation on the device: Film::rgb_image (src/tracer/film.rs:173-193) = Pixel::value (film.rs:82-90)
 * + splat_scale * splat / filter_integral, then TransferFunction::apply (src/tracer/color/space.rs:8-36;
 * transfer 0 = the sRGB curve used by sRGB and DCI-P3, 1 = the rec. 2020 curve) with Rust's saturating
 * `as u8`.  rgb8 = [n_pixels*3] HOST bytes, row-major like the accumulators.  The _dev variant reads the
 * accumulators where lumo_gpu_render_dev (and the multi-GPU reduce) left them, so the host receives
 * 3 B/pixel instead of 56 B/pixel; the other one takes host accumulators (e.g. after a host-side merge).
 * kernel_ms (optional): CUDA-event time of the kernel on the context's stream. */
int32_t lumo_gpu_film_encode_dev(lumo_ctx* ctx, const double* pixels_dev, const double* splats_dev, uint64_t n_pixels, double splat_scale,
                                 double filter_integral, int32_t transfer, uint8_t* rgb8, float* kernel_ms);
int32_t lumo_gpu_film_encode(lumo_ctx* ctx, const double* pixels, const double* splats, uint64_t n_pixels, double splat_scale,
                             double filter_integral, int32_t transfer, uint8_t* rgb8);

/* Traversal visit counters (N_tlas, N_inst, N_kd, N_idx, N_tri, N_sphere of DESIGN.md's byte
 * formula).  While enabled, the traversal kernels of this context run their counting instantiation;
 * never enabled inside a timed region. */
int32_t lumo_gpu_ctx_count_visits(lumo_ctx* ctx, int32_t enable);
int32_t lumo_gpu_ctx_visits(lumo_ctx* ctx, uint64_t* out12);   /* 6 for the closest-hit kernels, then 6 for the occlusion kernels */
/* Closest hits (Scene::hit) go through a world-space BVH with the reference traversal run on the winning object only, and
 * through the full reference traversal wherever that is not provably the same (csrc/gpu/closest.cuh).  mode 0: that (default);
 * 1: the reference traversal for every ray.  stats (only while visit counting is on; reset by lumo_gpu_ctx_count_visits):
 * [0] BVH nodes, [1] leaf primitives, [2] triangle tests, [3] sphere tests, [4] rays sent to the reference traversal, [5] rays,
 * [6..13] why: stack overflow / another object at or below the nearest hit / a box above the winner fails / the winner's full
 * hit is rejected / its any-hit or full distance is not the nearest hit / the last three again for the nearest light.
 * stats has FOURTEEN entries. */
int32_t lumo_gpu_ctx_closest_mode(lumo_ctx* ctx, int32_t mode);
int32_t lumo_gpu_ctx_closest_stats(lumo_ctx* ctx, uint64_t* stats14);
/* Work of the shading stage in the context's last PathTrace / DirectLight render (bench.py's byte count of the NEE queues):
 * [0] bounces that ran next-event estimation, [1] NEE terms evaluated, [2] shadow rays queued, [3] bounces shaded. */
int32_t lumo_gpu_ctx_shade_stats(lumo_ctx* ctx, uint64_t* stats4);
/* Shadow rays are answered from an order-free occlusion BVH and confirmed by the reference's own per-object traversal
 * (csrc/gpu/occlude.cuh).  mode 0: that (default); 1: the reference's object BVH + kd-tree traversal for shadow rays too;
 * 2: both on every shadow ray of a render, disagreements counted in stats[7].  stats: [0] BVH nodes, [1] leaf primitives,
 * [2] triangle tests, [3] sphere tests, [4] candidates sent to the confirmation pass, [5] confirmed, [6] sent to the faithful
 * kernel, [8] blockers accepted as robust without confirmation (all only while visit counting is on), [7] disagreements.
 * stats has NINE entries.  Reset by lumo_gpu_ctx_count_visits. */
int32_t lumo_gpu_ctx_occlusion_mode(lumo_ctx* ctx, int32_t mode);
int32_t lumo_gpu_ctx_occlusion_stats(lumo_ctx* ctx, uint64_t* stats9);
/* Device time (ms, CUDA events on the launching stream) and launch count per kernel class of the last
 * lumo_gpu_render*: [0] regen (film + refill + compaction), [1] closest-hit trace, [2] shade, [3] occlusion trace. */
int32_t lumo_gpu_ctx_kernel_times(lumo_ctx* ctx, double* ms4, uint64_t* launches4);
/* Per wave iteration of the last render's main pass: out[2i] = closest-hit rays traced, out[2i+1] = shadow rays traced. */
int32_t lumo_gpu_ctx_iter_log(lumo_ctx* ctx, uint32_t* out, uint32_t cap, uint32_t* n);

/* Device-resident variants used by bench.py's kernel-only timing (inputs already in HBM). */
int32_t lumo_gpu_trace_closest_dev(lumo_scene* scene, const double* origin_dev, const double* dir_dev, uint64_t n,
                                   uint32_t* obj_dev, uint32_t* tri_dev, double* t_dev, double* bary_dev, float* kernel_ms);

/* The FP64 issue ceiling of the context's GPU, measured: eight independent DFMA chains per thread over a full grid; TFLOP/s
 * (2 flops per DFMA), best of three launches, and that launch's time.  bench.py reports it as roofline.fp64_peak. */
int32_t lumo_gpu_fp64_peak(lumo_ctx* ctx, double* tflops, double* ms);

/* Parity hook: evaluates csrc/common/lumo_math.h (the FMA-free sin / cos / atan2 / acos / atanh / cosh / exp / log / pow every side
 * of the path shares) on the device.  fn: 0 sin, 1 cos, 2 atan2(x, y), 3 acos, 4 atanh, 5 cosh, 6 exp, 7 log, 8 pow(x, y); y may be
 * NULL for the one-argument functions.  Host pointers. */
int32_t lumo_gpu_math_eval(lumo_ctx* ctx, int32_t fn, const double* x, const double* y, uint64_t n, double* out);

const char* lumo_gpu_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
