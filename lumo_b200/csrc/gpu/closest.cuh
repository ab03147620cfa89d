// Closest hits (`Scene::hit`, src/tracer/scene.rs:119-147) through the world-space BVH of occlude.cuh, with the reference's
// own traversal run only on the object that wins — and, for every ray where that is not provably the reference's result,
// the reference traversal of trace.cuh as before.
//
// What `Scene::hit` computes: for Scene.objects, then for Scene.lights with t_max = the objects' hit,
// BVH::_hit walks the object BVH in build order and keeps the object whose ANY-hit distance `hit_t(r, t_min, tt)` is
// below the running bound tt (bvh.rs:343-352); the winner is then intersected in full (`hit`, bvh.rs:368).  For a kd-tree
// the any-hit distance is the first triangle the front-to-back walk accepts, not necessarily the nearest one, so in
// general the winner depends on the visit order.  It does not when
//   (1) the nearest hit t1 of the whole group belongs to object G and no other object of the group has a hit at a
//       distance <= t1 (no tie; the kernels even ask for a hair of 1e-9 t1 of clearance), and
//   (2) G's any-hit distance A — the first triangle its own kd walk accepts — equals t1.
// Then every object visited before G reports a distance above t1 (or nothing), G reports A = t1 whatever bound it is handed
// (a bound above A does not change the first accepted triangle), and nothing visited after G can get below it: G wins, and
// the result is G's full intersection.  All distances involved are Triangle::hit_t / Sphere::hit_t values computed with the
// reference's arithmetic in the object's frame, so (1) and (2) are exact comparisons, not estimates.
//
//   k_closest_bvh     lane-refilled, near-child-first walk of the world-space BVH (f32 boxes only cull, conservatively):
//                     per group the nearest hit (t1, G) and the nearest hit of any other object (conservatively), with
//                     the reference's f64 primitive tests; lights are only of interest below the objects' t1.
//   k_closest_finish  for G: the box tests of the object-BVH nodes above it with the tightest bound the reference could
//                     hold there (bvh.rs:333-335), then `Object::hit` = the reference's kd walk, which also yields A;
//                     checks (1), (2) and that the full hit's distance is t1; the same for the nearest light below it.
//                     Any check that fails (a tie, A != t1, a rejected full hit, a traversal stack overflow) sends the ray to
//   k_closest_fallback  `scene_hit` of trace.cuh — the reference traversal, unchanged.
// tests/test_trace_parity.py compares ids, t and barycentrics bit for bit with the oracle through this pipeline.
#pragma once
#include "occlude.cuh"

namespace lumo_dev {

struct ClosestScratch {    // per ray of a launch
    double *t1, *tl;       // nearest hit among Scene.objects / among Scene.lights (below t1)
    uint32_t *o1, *ol;     // their objects (global index), LUMO_NONE if none
    uint32_t* flags;       // bit 0: another object hit at <= t1, bit 1: another light hit at <= tl, bit 2: traversal stack overflow
    uint32_t* fallback_i;  // rays for k_closest_fallback
    uint32_t* counters;    // [0] work cursor of k_closest_bvh, [1] cursor of k_closest_finish, [2] fallback queue size, [3] fallback cursor
};
struct ClosestCounters { unsigned long long nodes, prims, tris, spheres, fallback, rays, why[8]; };   // why: reasons for the reference traversal (see k_closest_finish)

#define LUMO_CH_STACK 48
#ifndef LUMO_CH_REFILL
#define LUMO_CH_REFILL 8
#endif
#ifndef LUMO_CH_NODE_ROUND
#define LUMO_CH_NODE_ROUND 4
#endif

// One inner-node step of the ordered walk: hit children are pushed with their entry distances, the nearest is walked next.
// A stack entry is one 64-bit word — entry distance (f32 bits) above the node — so that a push or a pop is one local-memory access.
__device__ __forceinline__ unsigned long long ch_entry(uint32_t node, float t) { return ((unsigned long long)__float_as_uint(t) << 32) | node; }
template <bool CNT>
__device__ __forceinline__ bool ch_node_step(const DevScene& S, const AhRay& a, uint32_t& node, unsigned long long* stack, int& sp, bool& over, ClosestCounters* c) {
    if (CNT) c->nodes++;
    const float4* n4 = reinterpret_cast<const float4*>(S.ah_nodes + node);
    const float4 ax = __ldg(n4 + a.ox), bx = __ldg(n4 + 3 - a.ox);
    const float4 ay = __ldg(n4 + 1 + a.oy), by = __ldg(n4 + 4 - a.oy);
    const float4 az = __ldg(n4 + 2 + a.oz), bz = __ldg(n4 + 5 - a.oz);
    const uint4 ch = __ldg(reinterpret_cast<const uint4*>(n4 + 6));
    const float n0 = fmaxf(fmaxf(__fmaf_rn(ax.x, a.ix, a.nx), __fmaf_rn(ay.x, a.iy, a.ny)), fmaxf(__fmaf_rn(az.x, a.iz, a.nz), 0.0f));
    const float n1 = fmaxf(fmaxf(__fmaf_rn(ax.y, a.ix, a.nx), __fmaf_rn(ay.y, a.iy, a.ny)), fmaxf(__fmaf_rn(az.y, a.iz, a.nz), 0.0f));
    const float n2 = fmaxf(fmaxf(__fmaf_rn(ax.z, a.ix, a.nx), __fmaf_rn(ay.z, a.iy, a.ny)), fmaxf(__fmaf_rn(az.z, a.iz, a.nz), 0.0f));
    const float n3 = fmaxf(fmaxf(__fmaf_rn(ax.w, a.ix, a.nx), __fmaf_rn(ay.w, a.iy, a.ny)), fmaxf(__fmaf_rn(az.w, a.iz, a.nz), 0.0f));
    const float f0 = fminf(fminf(__fmaf_rn(bx.x, a.ix, a.fx), __fmaf_rn(by.x, a.iy, a.fy)), fminf(__fmaf_rn(bz.x, a.iz, a.fz), a.tmax));
    const float f1 = fminf(fminf(__fmaf_rn(bx.y, a.ix, a.fx), __fmaf_rn(by.y, a.iy, a.fy)), fminf(__fmaf_rn(bz.y, a.iz, a.fz), a.tmax));
    const float f2 = fminf(fminf(__fmaf_rn(bx.z, a.ix, a.fx), __fmaf_rn(by.z, a.iy, a.fy)), fminf(__fmaf_rn(bz.z, a.iz, a.fz), a.tmax));
    const float f3 = fminf(fminf(__fmaf_rn(bx.w, a.ix, a.fx), __fmaf_rn(by.w, a.iy, a.fy)), fminf(__fmaf_rn(bz.w, a.iz, a.fz), a.tmax));
    const float rel = 1.00000095367431640625f;   // 1 + 2^-20
    const bool h0 = n0 <= f0 * rel && ch.x != LUMO_NONE, h1 = n1 <= f1 * rel && ch.y != LUMO_NONE, h2 = n2 <= f2 * rel && ch.z != LUMO_NONE, h3 = n3 <= f3 * rel && ch.w != LUMO_NONE;
    // the nearest hit child is walked next, the others go on the stack with their entry distances (in child order)
    const float inf = __int_as_float(0x7F800000);
    float best = h0 ? n0 : inf; uint32_t next = h0 ? ch.x : LUMO_NONE;
    if (h1 && n1 < best) { best = n1; next = ch.y; } else if (h1 && next == LUMO_NONE) next = ch.y;
    if (h2 && n2 < best) { best = n2; next = ch.z; } else if (h2 && next == LUMO_NONE) next = ch.z;
    if (h3 && n3 < best) { best = n3; next = ch.w; } else if (h3 && next == LUMO_NONE) next = ch.w;
    if (next == LUMO_NONE) return false;
    if (sp + 3 > LUMO_CH_STACK) { over = (int)h0 + (int)h1 + (int)h2 + (int)h3 - 1 + sp > LUMO_CH_STACK; if (over) return true; }
    if (h0 && ch.x != next) stack[sp++] = ch_entry(ch.x, n0);
    if (h1 && ch.y != next) stack[sp++] = ch_entry(ch.y, n1);
    if (h2 && ch.z != next) stack[sp++] = ch_entry(ch.z, n2);
    if (h3 && ch.w != next) stack[sp++] = ch_entry(ch.w, n3);
    node = next;
    return true;
}
// pops the next node that can still hold something at or below the current bound; false: the walk is over
__device__ __forceinline__ bool ch_pop(const AhRay& a, uint32_t& node, const unsigned long long* stack, int& sp) {
    while (sp > 0) {
        const unsigned long long e = stack[--sp];
        if (__uint_as_float((uint32_t)(e >> 32)) <= a.tmax * 1.00000095367431640625f) { node = (uint32_t)e; return true; }
    }
    return false;
}
// one leaf primitive: the reference's own f64 test in the primitive's object frame, distances up to `bound` accepted
template <bool CNT>
__device__ __forceinline__ double ch_prim_test(const DevScene& S, const uint2 pr, const Ray& ray, const RayTri& q, double bound, AhLocal& L, ClosestCounters* c) {
    const uint32_t obj = pr.y & ~LUMO_AH_INSTANCED;
    if (pr.x & LUMO_AH_SPHERE) {
        if (CNT) c->spheres++;
        const Ray lr = (pr.y & LUMO_AH_INSTANCED) ? to_local<false>(S, S.objects[obj], ray, nullptr) : ray;
        return sphere_hit_t(S.spheres[pr.x & ~LUMO_AH_SPHERE].radius, lr, 0.0, bound);
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
    return tri_hit<false, false>(S.tri_verts + pr.x, lr, lq, 0.0, bound, th, nullptr) ? th.t : LUMO_INF;
}

// a hair behind t (see k_closest_bvh)
__device__ __forceinline__ double ch_ext(double t) { return t + fabs(t) * 1e-9 + 1e-300; }

// Source: n(), load(i, Ray&, t_max&).
template <bool CNT, class Source>
__global__ void __launch_bounds__(128, LUMO_BVH_BLOCKS) k_closest_bvh(const __grid_constant__ DevScene S, const Source src, const ClosestScratch Q, ClosestCounters* gc) {
    __shared__ double local_ctx[LUMO_AH_LOCAL_DOUBLES * 128];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = src.n();
    const uint32_t n_objects = S.P.n_objects;
    ClosestCounters cnt = {};
    AhLocal L; L.slot = local_ctx + threadIdx.x; L.stride = 128; L.cur = -1;
    unsigned long long stack[LUMO_CH_STACK];
    bool active = false, exhausted = false;
    uint32_t i = 0, node = 0, leaf_pos = 0, leaf_end = 0, o1 = LUMO_NONE, ol = LUMO_NONE, flags = 0;
    int sp = 0;
    Ray r; RayTri q; AhRay a;
    double t1 = 0.0, tl = 0.0, s1 = 0.0, sl = 0.0;     // nearest per group; nearest of any other object per group (conservative)
    r.o = d3(0, 0, 0); r.d = d3(0, 0, 1); q = ray_tri_setup(r); a = ah_make_ray(r, 0.0);
    for (;;) {
        const uint32_t free_m = __ballot_sync(0xFFFFFFFFu, !active && !exhausted);
        const uint32_t busy_m = __ballot_sync(0xFFFFFFFFu, active);
        if (free_m && ((uint32_t)__popc(free_m) >= LUMO_CH_REFILL || busy_m == 0u)) {
            uint32_t base = 0;
            if (lane == (uint32_t)(__ffs(free_m) - 1)) base = atomicAdd(&Q.counters[0], (uint32_t)__popc(free_m));
            base = __shfl_sync(0xFFFFFFFFu, base, __ffs(free_m) - 1);
            if (!active && !exhausted) {
                i = base + __popc(free_m & ((1u << lane) - 1u));
                if (i < n) {
                    double t_max; src.load(i, r, t_max);
                    q = ray_tri_setup(r); a = ah_make_ray(r, t_max);
                    t1 = t_max; tl = t_max; s1 = LUMO_INF; sl = LUMO_INF; o1 = LUMO_NONE; ol = LUMO_NONE; flags = 0;
                    L.cur = -1; node = 0; sp = 0; leaf_pos = leaf_end = 0; active = true;
                } else exhausted = true;
            }
        } else if (busy_m == 0u) break;
        bool done = false;
#pragma unroll 1
        for (int step = 0; step < LUMO_CH_NODE_ROUND; step++) {
            if (active && !done && leaf_pos == leaf_end && !(node & LUMO_AH_LEAF)) {
                bool over = false;
                if (!ch_node_step<CNT>(S, a, node, stack, sp, over, &cnt)) { if (!ch_pop(a, node, stack, sp)) done = true; }
                if (over) { flags |= 4u; done = true; }
            }
        }
        if (active && !done && leaf_pos == leaf_end && (node & LUMO_AH_LEAF)) { leaf_pos = node & 0x07FFFFFFu; leaf_end = leaf_pos + ((node >> 27) & 0xFu) + 1u; }
        if (active && !done && leaf_pos < leaf_end) {
            const uint2 pr = __ldg(reinterpret_cast<const uint2*>(S.ah_prims + leaf_pos));
            if (CNT) cnt.prims++;
            const uint32_t obj = pr.y & ~LUMO_AH_INSTANCED;
            leaf_pos++;
            // Distances up to a hair (1e-9, relative) behind the nearest hit are looked at too: an object that close behind the
            // winner sends the ray to the reference traversal, and in return k_closest_finish may assume that the bound the
            // reference holds when it reaches the winner's boxes is at least that hair above t1 (flat boxes — walls — have
            // their own entry distance within an ulp of the hit).
            if (obj < n_objects) {
                const double lim = ch_ext(t1);
                const double t = ch_prim_test<CNT>(S, pr, r, q, lim, L, &cnt);
                if (t <= lim && t < LUMO_INF) {                  // (a miss is +inf)
                    if (obj == o1) { if (t < t1) t1 = t; }
                    else if (t < t1 || o1 == LUMO_NONE) { if (o1 != LUMO_NONE) s1 = fmin(s1, t1); t1 = t; o1 = obj; }
                    else s1 = fmin(s1, t);
                    const double e = ch_ext(t1);
                    float tm = (float)e; if ((double)tm < e) tm = nextafterf(tm, INFINITY);
                    a.tmax = tm;                                 // nothing beyond the nearest object hit is of interest (lights: only below it)
                }
            } else {
                const double lim = fmin(t1, ch_ext(tl));
                const double t = ch_prim_test<CNT>(S, pr, r, q, lim, L, &cnt);
                if (t <= lim && t < LUMO_INF) {
                    if (obj == ol) { if (t < tl) tl = t; }
                    else if (t < tl || ol == LUMO_NONE) { if (ol != LUMO_NONE) sl = fmin(sl, tl); tl = t; ol = obj; }
                    else sl = fmin(sl, t);
                }
            }
            if (leaf_pos == leaf_end) { if (!ch_pop(a, node, stack, sp)) done = true; }
        }
        if (done) {
            if (o1 != LUMO_NONE && s1 <= ch_ext(t1)) flags |= 1u;
            if (ol != LUMO_NONE && sl <= ch_ext(tl)) flags |= 2u;
            Q.t1[i] = t1; Q.tl[i] = tl; Q.o1[i] = o1; Q.ol[i] = ol; Q.flags[i] = flags;
            active = false;
        }
    }
    if (CNT) { atomicAdd(&gc->nodes, cnt.nodes); atomicAdd(&gc->prims, cnt.prims); atomicAdd(&gc->tris, cnt.tris); atomicAdd(&gc->spheres, cnt.spheres); }
}

// The box tests of the object-BVH nodes above `obj` (bvh.rs:333-335).  When the reference reaches them its bound tt is at least
// `tt_low` (every other object's hits are beyond it), so passing with tt = tt_low implies passing with the reference's tt.
// Only the LEAF that lists the object has to be tested: an inner node's box is the exact min / max merge of its children's
// (bvh.rs:282-309), and AaBoundingBox::intersect is monotone in the box — subtraction of the origin, multiplication by 1 / d and
// min / max are all monotone under rounding, and a NaN dropped by min / max (0 * inf on a coinciding plane) only removes a
// constraint — so every node above the leaf starts no later and ends no earlier than the leaf does.
template <bool CNT>
__device__ __forceinline__ bool ch_path_ok(const DevScene& S, uint32_t obj, const RayCtx& w, double tt_low, Counters* c) {
    const LumoTlasNode* node = S.tlas + S.obj_path[S.obj_path_off[obj + 1] - 1u];
    LUMO_CNT(tlas);
    double t_start, t_end;
    box_intersect(node->lo, node->hi, w.r.o, w.inv, t_start, t_end);
    t_start = fmax(t_start, 0.0); t_end = fmin(t_end, tt_low);
    return t_start <= t_end;
}
__device__ __forceinline__ bool same_bits(double a, double b) { return __double_as_longlong(a) == __double_as_longlong(b); }

// (Measured and dropped: ordering the rays by the size class of the winning object between the BVH walk and this kernel, so that a
// warp holds either two-triangle walls or large meshes: the permuted, uncoalesced ray and scratch accesses cost more than the
// homogeneous warps save — trace class 88.0 -> 92.9 ms on bistro 4 spp, 54.5 -> 65.4 on dragon, 39.6 -> 43.1 on bunny.)
// Sink: store(i, have, HitRec) — the final hit record of ray i.
template <bool CNT, class Source, class Sink>
__global__ void __launch_bounds__(128, LUMO_WAVE_TRACE_BLOCKS) k_closest_finish(const __grid_constant__ DevScene S, const Source src, const Sink sink, const ClosestScratch Q, Counters* vc, ClosestCounters* gc) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = src.n();
    Counters cnt = {0, 0, 0, 0, 0, 0};
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&Q.counters[1], 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + lane;
        bool again = false;
        if (i < n) {
            Ray ray; double t_max; src.load(i, ray, t_max);
            const uint32_t o1 = Q.o1[i], ol = Q.ol[i], flags = Q.flags[i];
            const double t1 = Q.t1[i], tl = Q.tl[i];
            HitRec h; bool have = false;
            double t_h = t_max;
            int why = -1;      // 0 stack overflow, 1 another object at <= t1, 2 a box above the winner fails, 3 the winner's full hit is rejected, 4 its any-hit or full distance is not t1; 5..7: the same for the light
            if (flags & 4u) why = 0;
            RayCtx w;
            if (why < 0 && (o1 != LUMO_NONE || ol != LUMO_NONE)) make_ctx(ray, w);
            if (why < 0 && o1 != LUMO_NONE) {
                double any_t;
                if (flags & 1u) why = 1;
                else if (!ch_path_ok<CNT>(S, o1, w, fmin(t_max, ch_ext(t1)), &cnt)) why = 2;
                else if (!object_hit<CNT, 64, LUMO_WAVE_KD_ROUND>(S, S.objects[o1], w, 0.0, t_max, h, &cnt, &any_t)) why = 3;     // a rejected full hit empties the whole group
                else if (!same_bits(any_t, t1) || !same_bits(h.t, t1)) why = 4;
                else { h.obj = o1; have = true; t_h = h.t; }
            }
            if (why < 0 && ol != LUMO_NONE && tl < t_h) {      // Scene::hit: lights with t_max = the objects' hit; a light has to be strictly nearer
                HitRec hl; double any_t;
                if (flags & 2u) why = 5;
                else if (!ch_path_ok<CNT>(S, ol, w, fmin(t_h, ch_ext(tl)), &cnt)) why = 6;
                else if (!object_hit<CNT, 64, LUMO_WAVE_KD_ROUND>(S, S.objects[ol], w, 0.0, t_h, hl, &cnt, &any_t)) why = 7;
                else if (!same_bits(any_t, tl)) why = 7;
                else { hl.obj = ol; h = hl; have = true; }
            }
            again = why >= 0;
            if (CNT && again) atomicAdd(&gc->why[why], 1ull);
            if (!again) sink.store(i, have, h);
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, again);
        if (m) {
            uint32_t b = 0;
            if (lane == (uint32_t)(__ffs(m) - 1)) b = atomicAdd(&Q.counters[2], (uint32_t)__popc(m));
            b = __shfl_sync(0xFFFFFFFFu, b, __ffs(m) - 1);
            if (again) Q.fallback_i[b + __popc(m & ((1u << lane) - 1u))] = i;
        }
    }
    if (CNT) { atomicAdd(&vc->tlas, cnt.tlas); atomicAdd(&vc->inst, cnt.inst); atomicAdd(&vc->kd, cnt.kd); atomicAdd(&vc->leaf, cnt.leaf); atomicAdd(&vc->tri, cnt.tri); atomicAdd(&vc->sphere, cnt.sphere); }
}

template <bool CNT, class Source, class Sink>
__global__ void __launch_bounds__(128, LUMO_WAVE_TRACE_BLOCKS) k_closest_fallback(const __grid_constant__ DevScene S, const Source src, const Sink sink, const ClosestScratch Q, Counters* vc, ClosestCounters* gc) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = Q.counters[2];
    Counters cnt = {0, 0, 0, 0, 0, 0};
    // The queue is short (0.2 % of a launch's rays on the street scene) and every ray is a long chain of dependent loads whose
    // lanes diverge: 32 of them in one warp run nearly one after the other.  So the rays are spread over all resident warps
    // first — one ray per warp while there are fewer rays than warps — and the kernel's time is one ray's latency, not 32.
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t per_warp = min(32u, max(1u, (n + warps - 1u) / warps));
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&Q.counters[3], per_warp);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const uint32_t j = base + lane;
        if (lane < per_warp && j < n) {
            const uint32_t i = Q.fallback_i[j];
            Ray ray; double t_max; src.load(i, ray, t_max);
            HitRec h;
            const bool have = scene_hit<CNT, LUMO_WAVE_KD_ROUND>(S, ray, t_max, h, &cnt);
            sink.store(i, have, h);
        }
    }
    if (CNT) {
        atomicAdd(&vc->tlas, cnt.tlas); atomicAdd(&vc->inst, cnt.inst); atomicAdd(&vc->kd, cnt.kd); atomicAdd(&vc->leaf, cnt.leaf); atomicAdd(&vc->tri, cnt.tri); atomicAdd(&vc->sphere, cnt.sphere);
        if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&gc->fallback, (unsigned long long)n); atomicAdd(&gc->rays, (unsigned long long)src.n()); }
    }
}

}  // namespace lumo_dev
