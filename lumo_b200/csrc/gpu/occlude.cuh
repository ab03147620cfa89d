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
//   3. a candidate whose hit point lies well inside its triangle's bounding box is a blocker for the reference as well
//      (ah_hit_is_robust gives the argument); for every other candidate k_occl_confirm runs what the reference would have run
//      for that object: the box tests of the object-BVH nodes
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

struct AhCounters { unsigned long long nodes, prims, tris, spheres, candidates, confirmed, fallback, mismatches, robust; };

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

// One step through an inner node: the four slab tests; the hit child that starts first is walked next, the others go on
// the stack.  Returns false when nothing is left to walk (no hit child and an empty stack).  over: the stack was full.
template <bool CNT>
__device__ __forceinline__ bool ah_node_step(const DevScene& S, const AhRay& a, uint32_t& node, uint32_t* stack, int& sp, bool& over, AhCounters* c) {
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
    // (an empty child has an inverted infinite box, which fails the test by itself — except for garbage rays whose 1/d is 0 or NaN)
    const bool h0 = n0 <= f0 * rel && ch.x != LUMO_NONE, h1 = n1 <= f1 * rel && ch.y != LUMO_NONE, h2 = n2 <= f2 * rel && ch.z != LUMO_NONE, h3 = n3 <= f3 * rel && ch.w != LUMO_NONE;
    // the nearest hit child is walked next, the others go on the stack (in child order)
    const float inf = __int_as_float(0x7F800000);
    float best = h0 ? n0 : inf; uint32_t next = h0 ? ch.x : LUMO_NONE;
    if (h1 && n1 < best) { best = n1; next = ch.y; } else if (h1 && next == LUMO_NONE) next = ch.y;
    if (h2 && n2 < best) { best = n2; next = ch.z; } else if (h2 && next == LUMO_NONE) next = ch.z;
    if (h3 && n3 < best) { best = n3; next = ch.w; } else if (h3 && next == LUMO_NONE) next = ch.w;
    if (next != LUMO_NONE) {
        if (sp + 3 > LUMO_AH_STACK && (int)h0 + (int)h1 + (int)h2 + (int)h3 - 1 + sp > LUMO_AH_STACK) { over = true; return true; }
        if (h0 && ch.x != next) stack[sp++] = ch.x;
        if (h1 && ch.y != next) stack[sp++] = ch.y;
        if (h2 && ch.z != next) stack[sp++] = ch.z;
        if (h3 && ch.w != next) stack[sp++] = ch.w;
    }
    if (next == LUMO_NONE) { if (sp == 0) return false; next = stack[--sp]; }
    node = next;
    return true;
}

// A hit found through the occlusion BVH is *robust* when the hit point lies inside the triangle's own bounding box with a
// margin of 1e-9 (relative) on every axis.  Then the reference's traversal of this object is certain to report a blocker too,
// and the confirmation pass is skipped:
//   * every box above the triangle (its kd-tree's root box, the object-BVH nodes over the object) contains the triangle's
//     box, so the ray is inside each of them around t by the same margin — their slab tests (aabb.rs:33-44, rounding
//     errors of a few 1e-16) pass, and t lies inside the kd root interval;
//   * lumo's kd build lists a triangle in every leaf whose cell its box overlaps (kdtree/node.rs:198-230; checked on the
//     blob by tests/test_host_build.py).  Any split plane the walk's rounding could put on the wrong side of the hit is
//     within ~1e-15 of the hit point, hence cuts the triangle's box, hence both cells list the triangle: the front-to-back
//     walk (kdtree.rs:117-160) meets it in a cell whose exit is beyond t — unless it returns earlier with another hit,
//     which is a blocker just the same.  The distance itself is the same arithmetic (Triangle::hit_t) in both places.
// Axis-aligned triangles (zero-thickness boxes: walls, rectangles) and hits within the margin of a box face are not
// robust; they take the confirmation pass, which replays the reference's traversal of the object.
__device__ __forceinline__ bool ah_hit_is_robust(const LumoTriVerts* tv, const Ray& lr, double t) {
    D3 A, B, C; load_tri(tv, A, B, C);
    const D3 p = lr.o + t * lr.d;
    const D3 lo = d3(fmin(A.x, fmin(B.x, C.x)), fmin(A.y, fmin(B.y, C.y)), fmin(A.z, fmin(B.z, C.z)));
    const D3 hi = d3(fmax(A.x, fmax(B.x, C.x)), fmax(A.y, fmax(B.y, C.y)), fmax(A.z, fmax(B.z, C.z)));
    const double k = 1e-9;
    const double mx = k * (fmax(fabs(lo.x), fabs(hi.x)) + fabs(lr.o.x)) + 1e-300, my = k * (fmax(fabs(lo.y), fabs(hi.y)) + fabs(lr.o.y)) + 1e-300,
                 mz = k * (fmax(fabs(lo.z), fabs(hi.z)) + fabs(lr.o.z)) + 1e-300;
    return p.x > lo.x + mx && p.x < hi.x - mx && p.y > lo.y + my && p.y < hi.y - my && p.z > lo.z + mz && p.z < hi.z - mz && t > k * fabs(t) + 1e-300;
}

// One leaf primitive: the reference's own f64 test in the primitive's object frame.  Returns the hit distance or +inf;
// robust: see ah_hit_is_robust (spheres never are).
template <bool CNT>
__device__ __forceinline__ double ah_prim_test(const DevScene& S, uint32_t prim, const Ray& ray, const RayTri& q, double t_max, AhLocal& L, uint32_t& obj, bool& robust, AhCounters* c) {
    const uint2 pr = __ldg(reinterpret_cast<const uint2*>(S.ah_prims + prim));
    if (CNT) c->prims++;
    robust = false;
    obj = pr.y & ~LUMO_AH_INSTANCED;
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
        if (qo - r2 < -1e-9 * (qo + r2) && f_end < -1e-9 * scale) return LUMO_INF;
        return sphere_hit_t(radius, lr, 0.0, t_max);
    }
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
    if (!tri_hit<false, false>(S.tri_verts + pr.x, lr, lq, 0.0, t_max, th, nullptr)) return LUMO_INF;
    if (th.t < t_max) robust = ah_hit_is_robust(S.tri_verts + pr.x, lr, th.t);
    return th.t;
}

// The reference's verdict on ONE object for an any-hit query: every object-BVH node above it must pass the box test of
// bvh.rs:333-335 (t_start <= t_end with tt = t_max), then Object::hit_t (bvh.rs:346) must return a distance below t_max.
template <bool CNT>
__device__ __forceinline__ bool ah_confirm(const DevScene& S, uint32_t obj, const Ray& ray, double t_max, Counters* c) {
    RayCtx w; make_ctx(ray, w);
    {   // the leaf that lists the object; the nodes above it pass whenever it does (closest.cuh: ch_path_ok gives the argument)
        const LumoTlasNode* node = S.tlas + S.obj_path[S.obj_path_off[obj + 1] - 1u];
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

// The walk itself.  Rays differ a lot in how far they get (a blocked ray stops at its first hit, a free one walks on to the
// light), so a warp that took 32 rays and waited for the slowest ran at ~6 of 32 lanes (ncu, profiles/).  Here a lane
// whose ray is finished takes the next ray of the queue while the others continue (lanes are refilled once LUMO_AH_REFILL
// of them are free), and the lanes of a warp alternate between the two kinds of work together: up to LUMO_AH_NODE_ROUND
// inner-node steps, then one leaf primitive for every lane that holds one.
// (Measured and dropped: a blocker cache for the wave's shadow rays — the classic shadow cache keyed by (light, cell of the origin in a 16^3 grid), 4 M lines,
// the cached primitive tested ahead of the BVH root.  Verdicts are unchanged by construction, but on the street stand-in the
// blockers are small triangles that rarely repeat: occlusion class 151.9 -> 181.1 ms on bistro 4 spp; 8^3 / 32^3 grids and
// 16 M lines 177.9 / 183.1 / 191.1 ms; neutral on bunny, conference and dragon.)
// (Measured and dropped: prefetch.global.L1 of a leaf's primitive records when the walk first sees the leaf, and of the next
// primitive's vertices during a test: occlusion class 131.4 -> 138.5 ms on bistro 4 spp.)
// resident CTAs per SM the two BVH walks are compiled for (register budget 65536 / (128 * blocks)) and launched with.
// Same-box A/B on B200, bistro 4 spp, occlusion class ms: 4 CTAs (114 registers) 176.2, 5 (102) 155.3, 6 (85, spills) 161.4.
#ifndef LUMO_BVH_BLOCKS
#define LUMO_BVH_BLOCKS 5
#endif
#ifndef LUMO_AH_REFILL
#define LUMO_AH_REFILL 8
#endif
// (same A/B: refill at 4 free lanes 181.8 / at 8: 176.2; node steps per round 2: 185.2, 3: 176.2, 4: 173.1)
#ifndef LUMO_AH_NODE_ROUND
#define LUMO_AH_NODE_ROUND 4
#endif
template <bool CNT, class Source, class Sink>
__global__ void __launch_bounds__(128, LUMO_BVH_BLOCKS) k_occl_bvh(const __grid_constant__ DevScene S, const Source src, const Sink sink, const OcclQueues Q, AhCounters* gc) {
    __shared__ double local_ctx[LUMO_AH_LOCAL_DOUBLES * 128];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = src.n();
    AhCounters cnt = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    AhLocal L; L.slot = local_ctx + threadIdx.x; L.stride = 128; L.cur = -1;
    uint32_t stack[LUMO_AH_STACK];
    // per-lane state of the ray in flight
    bool active = false, exhausted = false;
    uint32_t i = 0, node = 0, leaf_pos = 0, leaf_end = 0;
    int sp = 0;
    Ray r; RayTri q; double t_max = 0.0; AhRay a;
    r.o = d3(0, 0, 0); r.d = d3(0, 0, 1); q = ray_tri_setup(r); a = ah_make_ray(r, 0.0);
    for (;;) {
        // ---- refill: free lanes take the next rays of the queue (one atomic per warp) ----
        const uint32_t free_m = __ballot_sync(0xFFFFFFFFu, !active && !exhausted);
        const uint32_t busy_m = __ballot_sync(0xFFFFFFFFu, active);
        if (free_m && ((uint32_t)__popc(free_m) >= LUMO_AH_REFILL || busy_m == 0u)) {
            uint32_t base = 0;
            if (lane == (uint32_t)(__ffs(free_m) - 1)) base = atomicAdd(&Q.counters[0], (uint32_t)__popc(free_m));
            base = __shfl_sync(0xFFFFFFFFu, base, __ffs(free_m) - 1);
            if (!active && !exhausted) {
                i = base + __popc(free_m & ((1u << lane) - 1u));
                if (i < n) {
                    src.load(i, r, t_max);
                    q = ray_tri_setup(r); a = ah_make_ray(r, t_max);
                    L.cur = -1; node = 0; sp = 0; leaf_pos = leaf_end = 0; active = true;      // node 0: the root is always an inner node
                } else exhausted = true;
            }
        } else if (busy_m == 0u) break;            // nothing in flight and nothing left to take
        // ---- inner nodes ----
        int verdict = -1;                            // 0: free of blockers, 1: candidate blocker (to be confirmed), 2: stack overflow, 3: robust blocker
        uint32_t cand = 0;
#pragma unroll 1
        for (int step = 0; step < LUMO_AH_NODE_ROUND; step++) {
            if (active && verdict < 0 && leaf_pos == leaf_end && !(node & LUMO_AH_LEAF)) {
                bool over = false;
                if (!ah_node_step<CNT>(S, a, node, stack, sp, over, &cnt)) verdict = 0;
                if (over) verdict = 2;
            }
        }
        // ---- one leaf primitive ----
        if (active && verdict < 0 && leaf_pos == leaf_end && (node & LUMO_AH_LEAF)) { leaf_pos = node & 0x07FFFFFFu; leaf_end = leaf_pos + ((node >> 27) & 0xFu) + 1u; }
        if (active && verdict < 0 && leaf_pos < leaf_end) {
            uint32_t obj; bool robust;
            const double t = ah_prim_test<CNT>(S, leaf_pos, r, q, t_max, L, obj, robust, &cnt);
            leaf_pos++;
            if (t < t_max) { verdict = robust ? 3 : 1; cand = obj; }
            else if (leaf_pos == leaf_end) { if (sp == 0) verdict = 0; else node = stack[--sp]; }
        }
        // ---- finished rays: verdict, or a place in the queues of the next two kernels ----
        if (verdict == 0) sink.verdict(i, false);
        if (verdict == 3) { sink.verdict(i, true); if (CNT) cnt.robust++; }
        const uint32_t m1 = __ballot_sync(0xFFFFFFFFu, verdict == 1), m2 = __ballot_sync(0xFFFFFFFFu, verdict == 2);
        if (m1) {
            uint32_t b = 0;
            if (lane == (uint32_t)(__ffs(m1) - 1)) b = atomicAdd(&Q.counters[1], (uint32_t)__popc(m1));
            b = __shfl_sync(0xFFFFFFFFu, b, __ffs(m1) - 1);
            if (verdict == 1) { const uint32_t j = b + __popc(m1 & ((1u << lane) - 1u)); Q.confirm_i[j] = i; Q.confirm_obj[j] = cand; }
            if (CNT) cnt.candidates += __popc(m1) * (lane == 0);
        }
        if (m2) {
            uint32_t b = 0;
            if (lane == (uint32_t)(__ffs(m2) - 1)) b = atomicAdd(&Q.counters[2], (uint32_t)__popc(m2));
            b = __shfl_sync(0xFFFFFFFFu, b, __ffs(m2) - 1);
            if (verdict == 2) Q.fallback_i[b + __popc(m2 & ((1u << lane) - 1u))] = i;
        }
        if (verdict >= 0) active = false;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) sink.done(n);
    if (CNT) { atomicAdd(&gc->nodes, cnt.nodes); atomicAdd(&gc->prims, cnt.prims); atomicAdd(&gc->tris, cnt.tris); atomicAdd(&gc->spheres, cnt.spheres); atomicAdd(&gc->candidates, cnt.candidates); atomicAdd(&gc->robust, cnt.robust); }
}

template <bool CNT, class Source, class Sink>
__global__ void __launch_bounds__(128, LUMO_WAVE_TRACE_BLOCKS) k_occl_confirm(const __grid_constant__ DevScene S, const Source src, const Sink sink, const OcclQueues Q, Counters* vc, AhCounters* gc) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = Q.counters[1];
    Counters cnt = {0, 0, 0, 0, 0, 0};
    unsigned long long n_conf = 0, n_fall = 0;
    // short queues are spread over all resident warps (see k_closest_fallback): the kernel then costs one ray's latency
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t per_warp = min(32u, max(1u, (n + warps - 1u) / warps));
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&Q.counters[3], per_warp);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const uint32_t j = base + lane;
        bool again = false; uint32_t i = 0;
        if (lane < per_warp && j < n) {
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
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t per_warp = min(32u, max(1u, (n + warps - 1u) / warps));
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&Q.counters[4], per_warp);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const uint32_t j = base + lane;
        if (lane < per_warp && j < n) {
            const uint32_t i = Q.fallback_i[j];
            Ray r; double t_max; src.load(i, r, t_max);
            sink.verdict(i, scene_occluded<CNT, LUMO_WAVE_KD_ROUND>(S, r, t_max, &cnt));
        }
    }
    if (CNT) { atomicAdd(&vc->tlas, cnt.tlas); atomicAdd(&vc->inst, cnt.inst); atomicAdd(&vc->kd, cnt.kd); atomicAdd(&vc->leaf, cnt.leaf); atomicAdd(&vc->tri, cnt.tri); atomicAdd(&vc->sphere, cnt.sphere); }
}

}  // namespace lumo_dev
