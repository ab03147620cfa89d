// Order-free occlusion: the occlusion half of Scene::hit_light (src/tracer/scene.rs:180-186) answered from the 4-wide
// world-space BVH of the scene blob (LSEC_AH_NODES / LSEC_AH_PRIMS, built by csrc/host/ah_bvh.h) instead of replaying
// lumo's unordered object BVH + kd-trees.
//
// Why this is allowed: `hit_light` asks "is ANY primitive hit in (0, t_light - 1e-10)?" and only the boolean is used
// (SURVEY A.8-i); BVH::hit_t's first-found distance (bvh.rs:371-374) is compared with t_max and thrown away.
// Why it is still the reference's answer, bit for bit:
//   1. the f32 boxes only cull, conservatively (rounded outwards at build time; the slab test below carries its own
//      rounding-error bounds), so no primitive the ray can hit is skipped;
//   2. a primitive that survives is tested with the reference's own f64 routine in its object's frame
//      (Triangle::hit_t = tri_hit<false>, Sphere::hit_t) — that only nominates a candidate object;
//   3. k_occl_confirm then runs what the reference would have run for that object: the box tests of the object-BVH nodes
//      above it (bvh.rs:333-335, with tt = t_max as in the any-hit pass) and Object::hit_t through the object's own
//      kd-tree (kdtree.rs:101-169, cell clipping included).  Occluded is reported only if that says so;
//   4. a ray whose candidate is not confirmed (the cell clipping or a box test disagreed by an ulp), or whose traversal
//      stack overflowed, goes to k_occl_fallback = the faithful scene_occluded of trace.cuh.
// tests/test_trace_parity.py checks the booleans against the oracle; LUMO_OCCLUDE_CHECK=1 makes the wave pipeline run the
// faithful kernel next to this one on every shadow ray of a render and count disagreements.
#pragma once
#include "trace.cuh"

namespace lumo_dev {

#define LUMO_AH_STACK 48

struct AhCounters { unsigned long long nodes, prims, tris, spheres, candidates, confirmed, fallback, mismatches; };

// f32 side of a ray for the slab tests.  t = fma(plane, inv, -(o * inv)) with absolute slack s = |o * inv| * 2^-22 folded
// into the two constants (near planes: t - s, far planes: t + s) and a relative slack of 2^-20 at the comparison:
//   |t_computed - t_true| <= |t_true| * 1.5 * 2^-23 + |o * inv| * 2^-23     (o, d and 1/d rounded to f32, one fma rounding)
struct AhRay {
    float ix, iy, iz;        // 1 / d (|d| clamped to >= 1e-18 so that it stays finite)
    float nx, ny, nz;        // -(o * inv) - s   -> entry distances, biased early
    float fx, fy, fz;        // -(o * inv) + s   -> exit distances, biased late
    float tmax;
    int ox, oy, oz;          // which float4 of the node holds the near planes of each axis (0: lo, 3: hi)
};
__device__ __forceinline__ float ah_clamp_dir(double d) {
    float f = (float)d;
    if (fabsf(f) < 1e-18f) f = signbit(d) ? -1e-18f : 1e-18f;
    return f;
}
__device__ __forceinline__ AhRay ah_make_ray(const Ray& r, double t_max) {
    AhRay a;
    const float dx = ah_clamp_dir(r.d.x), dy = ah_clamp_dir(r.d.y), dz = ah_clamp_dir(r.d.z);
    a.ix = __fdiv_rn(1.0f, dx); a.iy = __fdiv_rn(1.0f, dy); a.iz = __fdiv_rn(1.0f, dz);
    const float px = __fmul_rn((float)r.o.x, a.ix), py = __fmul_rn((float)r.o.y, a.iy), pz = __fmul_rn((float)r.o.z, a.iz);
    const float k = 2.384185791015625e-07f;   // 2^-22
    const float sx = __fmaf_rn(fabsf(px), k, 1e-30f), sy = __fmaf_rn(fabsf(py), k, 1e-30f), sz = __fmaf_rn(fabsf(pz), k, 1e-30f);
    a.nx = -px - sx; a.ny = -py - sy; a.nz = -pz - sz;
    a.fx = -px + sx; a.fy = -py + sy; a.fz = -pz + sz;
    float tm = (float)t_max; if ((double)tm < t_max) tm = nextafterf(tm, INFINITY);
    a.tmax = tm;
    a.ox = dx < 0.0f ? 3 : 0; a.oy = dy < 0.0f ? 3 : 0; a.oz = dz < 0.0f ? 3 : 0;
    return a;
}
__device__ __forceinline__ bool ah_slab(float ax, float ay, float az, float bx, float by, float bz, const AhRay& a) {
    const float tn = fmaxf(fmaxf(__fmaf_rn(ax, a.ix, a.nx), __fmaf_rn(ay, a.iy, a.ny)), fmaxf(__fmaf_rn(az, a.iz, a.nz), 0.0f));
    const float tf = fminf(fminf(__fmaf_rn(bx, a.ix, a.fx), __fmaf_rn(by, a.iy, a.fy)), fminf(__fmaf_rn(bz, a.iz, a.fz), a.tmax));
    return tn <= tf * 1.00000095367431640625f;   // 1 + 2^-20
}

// The ray in an instanced object's frame, kept per thread in shared memory (10 doubles + kz): recomputed only when the
// traversal moves to a primitive of another instance.
struct AhLocal { double* slot; int stride; int cur; };
__device__ __forceinline__ void ah_local_store(const AhLocal& L, const RayCtx& c) {
    double* p = L.slot; const int s = L.stride;
    p[0] = c.r.o.x; p[s] = c.r.o.y; p[2 * s] = c.r.o.z; p[3 * s] = c.r.d.x; p[4 * s] = c.r.d.y; p[5 * s] = c.r.d.z;
    p[6 * s] = c.q.sx; p[7 * s] = c.q.sy; p[8 * s] = c.q.sz; p[9 * s] = c.q.wz; p[10 * s] = (double)c.q.kz;
}
__device__ __forceinline__ void ah_local_load(const AhLocal& L, Ray& r, RayTri& q) {
    const double* p = L.slot; const int s = L.stride;
    r.o = d3(p[0], p[s], p[2 * s]); r.d = d3(p[3 * s], p[4 * s], p[5 * s]);
    q.sx = p[6 * s]; q.sy = p[7 * s]; q.sz = p[8 * s]; q.wz = p[9 * s]; q.kz = (int)p[10 * s];
}
#define LUMO_AH_LOCAL_DOUBLES 11

// Walks the occlusion BVH.  Returns 0: nothing the ray could hit, 1: candidate blocker (cand = global object index),
// 2: traversal stack overflow (the caller sends the ray to the faithful kernel).
template <bool CNT>
__device__ __forceinline__ int ah_any_hit(const DevScene& S, const Ray& ray, const RayTri& q, double t_max, AhLocal& L, uint32_t& cand, AhCounters* c) {
    const AhRay a = ah_make_ray(ray, t_max);
    uint32_t stack[LUMO_AH_STACK];
    int sp = 0;
    uint32_t node = 0;                       // the root is always an inner node (ah_bvh.h: collapse)
    for (;;) {
        // ---- inner nodes: test the four child boxes, continue into one hit child, push the others ----
        while (!(node & LUMO_AH_LEAF)) {
            if (CNT) c->nodes++;
            const float4* n4 = reinterpret_cast<const float4*>(S.ah_nodes + node);
            const float4 ax = __ldg(n4 + a.ox), bx = __ldg(n4 + 3 - a.ox);
            const float4 ay = __ldg(n4 + 1 + a.oy), by = __ldg(n4 + 4 - a.oy);
            const float4 az = __ldg(n4 + 2 + a.oz), bz = __ldg(n4 + 5 - a.oz);
            const uint4 ch = __ldg(reinterpret_cast<const uint4*>(n4 + 6));
            // entry distances (biased early) and exit distances (biased late) of the four children
            const float n0 = fmaxf(fmaxf(__fmaf_rn(ax.x, a.ix, a.nx), __fmaf_rn(ay.x, a.iy, a.ny)), fmaxf(__fmaf_rn(az.x, a.iz, a.nz), 0.0f));
            const float n1 = fmaxf(fmaxf(__fmaf_rn(ax.y, a.ix, a.nx), __fmaf_rn(ay.y, a.iy, a.ny)), fmaxf(__fmaf_rn(az.y, a.iz, a.nz), 0.0f));
            const float n2 = fmaxf(fmaxf(__fmaf_rn(ax.z, a.ix, a.nx), __fmaf_rn(ay.z, a.iy, a.ny)), fmaxf(__fmaf_rn(az.z, a.iz, a.nz), 0.0f));
            const float n3 = fmaxf(fmaxf(__fmaf_rn(ax.w, a.ix, a.nx), __fmaf_rn(ay.w, a.iy, a.ny)), fmaxf(__fmaf_rn(az.w, a.iz, a.nz), 0.0f));
            const float f0 = fminf(fminf(__fmaf_rn(bx.x, a.ix, a.fx), __fmaf_rn(by.x, a.iy, a.fy)), fminf(__fmaf_rn(bz.x, a.iz, a.fz), a.tmax));
            const float f1 = fminf(fminf(__fmaf_rn(bx.y, a.ix, a.fx), __fmaf_rn(by.y, a.iy, a.fy)), fminf(__fmaf_rn(bz.y, a.iz, a.fz), a.tmax));
            const float f2 = fminf(fminf(__fmaf_rn(bx.z, a.ix, a.fx), __fmaf_rn(by.z, a.iy, a.fy)), fminf(__fmaf_rn(bz.z, a.iz, a.fz), a.tmax));
            const float f3 = fminf(fminf(__fmaf_rn(bx.w, a.ix, a.fx), __fmaf_rn(by.w, a.iy, a.fy)), fminf(__fmaf_rn(bz.w, a.iz, a.fz), a.tmax));
            const float rel = 1.00000095367431640625f;   // 1 + 2^-20
            const bool h0 = n0 <= f0 * rel, h1 = n1 <= f1 * rel, h2 = n2 <= f2 * rel, h3 = n3 <= f3 * rel;
            // the hit child that starts first is walked next (a blocker is most likely found near the origin), the others wait on the stack
            uint32_t next = LUMO_NONE; float best = 0.0f;
            bool over = false;
#define LUMO_AH_TAKE(h, n, c)                                                                                       \
            if (h) {                                                                                                \
                if (next == LUMO_NONE) { next = c; best = n; }                                                      \
                else if (sp >= LUMO_AH_STACK) over = true;                                                          \
                else if (n < best) { stack[sp++] = next; next = c; best = n; }                                      \
                else stack[sp++] = c;                                                                               \
            }
            LUMO_AH_TAKE(h0, n0, ch.x) LUMO_AH_TAKE(h1, n1, ch.y) LUMO_AH_TAKE(h2, n2, ch.z) LUMO_AH_TAKE(h3, n3, ch.w)
#undef LUMO_AH_TAKE
            if (over) return 2;
            if (next == LUMO_NONE) { if (sp == 0) return 0; next = stack[--sp]; }
            node = next;
        }
        // ---- leaf: the reference's own tests, in the primitive's object frame ----
        const uint32_t first = node & 0x07FFFFFFu, count = ((node >> 27) & 0xFu) + 1u;
        for (uint32_t k = 0; k < count; k++) {
            const uint2 pr = __ldg(reinterpret_cast<const uint2*>(S.ah_prims + first + k));
            if (CNT) c->prims++;
            const uint32_t obj = pr.y & ~LUMO_AH_INSTANCED;
            double t;
            if (pr.x & LUMO_AH_SPHERE) {
                // Sphere::hit_t needs the ray in the sphere's frame but none of the triangle constants.  An enclosing sphere (the
                // environment light) is met by every ray: when origin and end point of the segment are both inside the ball by a
                // wide margin, the exit distance t1 is beyond t_max and the quadratic of sphere.rs:77-96 returns "no hit" — decided
                // here without the square root and the two divisions.
                if (CNT) c->spheres++;
                const double radius = S.spheres[pr.x & ~LUMO_AH_SPHERE].radius;
                const Ray lr = (pr.y & LUMO_AH_INSTANCED) ? to_local<false>(S, S.objects[obj], ray, nullptr) : ray;
                const double qa = dot(lr.d, lr.d), qb = 2.0 * dot(lr.d, lr.o), qo = dot(lr.o, lr.o), r2 = radius * radius;
                const double f_end = (qa * t_max) * t_max + qb * t_max + (qo - r2);
                const double scale = fabs(qa * t_max * t_max) + fabs(qb * t_max) + qo + r2;
                if (qo - r2 < -1e-9 * (qo + r2) && f_end < -1e-9 * scale) t = LUMO_INF;
                else t = sphere_hit_t(radius, lr, 0.0, t_max);
            } else {
                if (CNT) c->tris++;
                Ray lr = ray; RayTri lq = q;
                if (pr.y & LUMO_AH_INSTANCED) {
                    const int inst = S.objects[obj].inst;
                    if (inst != L.cur) {
                        RayCtx lc; make_ctx(to_local<false>(S, S.objects[obj], ray, nullptr), lc);
                        ah_local_store(L, lc); L.cur = inst;
                    }
                    ah_local_load(L, lr, lq);
                }
                TriHit th;
                t = tri_hit<false, false>(S.tri_verts + pr.x, lr, lq, 0.0, t_max, th, nullptr) ? th.t : LUMO_INF;
            }
            if (t < t_max) { cand = obj; return 1; }
        }
        if (sp == 0) return 0;
        node = stack[--sp];
    }
}

// The reference's verdict on ONE object for an any-hit query: every object-BVH node above it must pass the box test of
// bvh.rs:333-335 (t_start <= t_end with tt = t_max), then Object::hit_t (bvh.rs:346) must return a distance below t_max.
template <bool CNT>
__device__ __forceinline__ bool ah_confirm(const DevScene& S, uint32_t obj, const Ray& ray, double t_max, Counters* c) {
    RayCtx w; make_ctx(ray, w);
    const uint32_t p0 = S.obj_path_off[obj], p1 = S.obj_path_off[obj + 1];
    for (uint32_t p = p0; p < p1; p++) {
        const LumoTlasNode* node = S.tlas + S.obj_path[p];
        LUMO_CNT(tlas);
        double t_start, t_end;
        box_intersect(node->lo, node->hi, w.r.o, w.inv, t_start, t_end);
        t_start = fmax(t_start, 0.0); t_end = fmin(t_end, t_max);
        if (!(t_start <= t_end)) return false;
    }
    return object_hit_t<CNT, LUMO_WAVE_KD_ROUND>(S, S.objects[obj], w, 0.0, t_max, c) < t_max;
}

// ---- the three kernels, over any ray source / verdict sink -------------------------------------------------------------
// Source: n(), load(i, Ray&, t_max&).  Sink: verdict(i, occluded).
struct OcclQueues {
    uint32_t* confirm_i; uint32_t* confirm_obj; uint32_t* fallback_i;   // capacity = number of rays of the launch
    uint32_t* counters;   // [0] work cursor of k_occl_bvh, [1] confirm queue size, [2] fallback queue size, [3] confirm cursor, [4] fallback cursor
};

template <bool CNT, class Source, class Sink>
__global__ void __launch_bounds__(128, 4) k_occl_bvh(const __grid_constant__ DevScene S, const Source src, const Sink sink, const OcclQueues Q, AhCounters* gc) {
    __shared__ double local_ctx[LUMO_AH_LOCAL_DOUBLES * 128];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = src.n();
    AhCounters cnt = {0, 0, 0, 0, 0, 0, 0, 0};
    AhLocal L; L.slot = local_ctx + threadIdx.x; L.stride = 128;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&Q.counters[0], 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + lane;
        int verdict = -1; uint32_t cand = 0;
        if (i < n) {
            Ray r; double t_max; src.load(i, r, t_max);
            const RayTri q = ray_tri_setup(r);
            L.cur = -1;
            verdict = ah_any_hit<CNT>(S, r, q, t_max, L, cand, &cnt);
            if (verdict == 0) sink.verdict(i, false);
        }
        // candidates and overflows are compacted into dense queues for the next two kernels
        const uint32_t m1 = __ballot_sync(0xFFFFFFFFu, verdict == 1), m2 = __ballot_sync(0xFFFFFFFFu, verdict == 2);
        if (m1) {
            uint32_t b = 0;
            if (lane == (uint32_t)(__ffs(m1) - 1)) b = atomicAdd(&Q.counters[1], (uint32_t)__popc(m1));
            b = __shfl_sync(0xFFFFFFFFu, b, __ffs(m1) - 1);
            if (verdict == 1) { const uint32_t j = b + __popc(m1 & ((1u << lane) - 1u)); Q.confirm_i[j] = i; Q.confirm_obj[j] = cand; }
        }
        if (m2) {
            uint32_t b = 0;
            if (lane == (uint32_t)(__ffs(m2) - 1)) b = atomicAdd(&Q.counters[2], (uint32_t)__popc(m2));
            b = __shfl_sync(0xFFFFFFFFu, b, __ffs(m2) - 1);
            if (verdict == 2) Q.fallback_i[b + __popc(m2 & ((1u << lane) - 1u))] = i;
        }
        if (CNT) { cnt.candidates += __popc(m1) * (lane == 0); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) sink.done(n);
    if (CNT) { atomicAdd(&gc->nodes, cnt.nodes); atomicAdd(&gc->prims, cnt.prims); atomicAdd(&gc->tris, cnt.tris); atomicAdd(&gc->spheres, cnt.spheres); atomicAdd(&gc->candidates, cnt.candidates); }
}

template <bool CNT, class Source, class Sink>
__global__ void __launch_bounds__(128, LUMO_WAVE_TRACE_BLOCKS) k_occl_confirm(const __grid_constant__ DevScene S, const Source src, const Sink sink, const OcclQueues Q, Counters* vc, AhCounters* gc) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = Q.counters[1];
    Counters cnt = {0, 0, 0, 0, 0, 0};
    unsigned long long n_conf = 0, n_fall = 0;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&Q.counters[3], 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const uint32_t j = base + lane;
        bool again = false; uint32_t i = 0;
        if (j < n) {
            i = Q.confirm_i[j];
            Ray r; double t_max; src.load(i, r, t_max);
            if (ah_confirm<CNT>(S, Q.confirm_obj[j], r, t_max, &cnt)) { sink.verdict(i, true); n_conf++; }
            else { again = true; n_fall++; }
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, again);
        if (m) {
            uint32_t b = 0;
            if (lane == (uint32_t)(__ffs(m) - 1)) b = atomicAdd(&Q.counters[2], (uint32_t)__popc(m));
            b = __shfl_sync(0xFFFFFFFFu, b, __ffs(m) - 1);
            if (again) Q.fallback_i[b + __popc(m & ((1u << lane) - 1u))] = i;
        }
    }
    if (CNT) {
        atomicAdd(&vc->tlas, cnt.tlas); atomicAdd(&vc->inst, cnt.inst); atomicAdd(&vc->kd, cnt.kd); atomicAdd(&vc->leaf, cnt.leaf); atomicAdd(&vc->tri, cnt.tri); atomicAdd(&vc->sphere, cnt.sphere);
        atomicAdd(&gc->confirmed, n_conf); atomicAdd(&gc->fallback, n_fall);
    }
}

template <bool CNT, class Source, class Sink>
__global__ void __launch_bounds__(128, LUMO_WAVE_TRACE_BLOCKS) k_occl_fallback(const __grid_constant__ DevScene S, const Source src, const Sink sink, const OcclQueues Q, Counters* vc) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = Q.counters[2];
    Counters cnt = {0, 0, 0, 0, 0, 0};
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&Q.counters[4], 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const uint32_t j = base + lane;
        if (j < n) {
            const uint32_t i = Q.fallback_i[j];
            Ray r; double t_max; src.load(i, r, t_max);
            sink.verdict(i, scene_occluded<CNT, LUMO_WAVE_KD_ROUND>(S, r, t_max, &cnt));
        }
    }
    if (CNT) { atomicAdd(&vc->tlas, cnt.tlas); atomicAdd(&vc->inst, cnt.inst); atomicAdd(&vc->kd, cnt.kd); atomicAdd(&vc->leaf, cnt.leaf); atomicAdd(&vc->tri, cnt.tri); atomicAdd(&vc->sphere, cnt.sphere); }
}

}  // namespace lumo_dev
