// Node-major traversal for scenes whose object BVHs are small (every BASELINE scene that is "one big mesh in a box").
//
// Why: in those scenes a ray tests a handful of objects in the reference's fixed order — mostly two-triangle walls
// and, for some of the rays, one very large kd-tree.  With one ray per lane (trace.cuh) the whole warp pays for the
// big kd traversal while only the lanes whose ray entered the mesh's box work: ncu showed 6 of 32 lanes busy.
//
// How: the reference's order does not depend on the ray ("left first, then right", bvh.rs:326-360), so a whole batch
// can walk the BVH node by node, in that order, as a sequence of small kernels:
//   k_nm_segment  a run of the traversal for every ray: box tests of the nodes its own traversal visits (the parent's
//                 box was hit — a per-ray bit mask) and, in place, the light objects of hit leaves (walls, spheres,
//                 triangles) with the reference's update (strict '<'; first-found ends an occlusion query).  A run
//                 ends at a heavy object: the rays that reach it are stream-compacted into a queue;
//   k_nm_heavy    that one big kd-tree against the dense queue, lane-refilled (a finished lane pulls the next ray);
//   k_nm_winners  the winners' full intersection (bvh.rs:366-369): light ones in place, heavy ones via k_nm_heavy<GEO>;
// then the same for Scene.lights with the bound Scene::hit / hit_light prescribes.  Per ray the sequence of box tests,
// objects, bounds and comparisons is exactly the one of trace.cuh, so results are bit-identical
// (tests/test_trace_parity.py::test_node_major_bit_exact).
//
// MEASURED (B200, bunny 4 spp, closest + occlusion ms): nested default 32-36, this path 42-44.  Dense queues and lane
// refill do not raise the active lanes of the big kd-tree's kernel (5-7 of 32 with and without): its stalls are 51 %
// long-scoreboard (dependent kd node loads from L2) and 22 % fixed-latency f64 chains, and the extra passes over the
// per-ray state cost more than the idle lanes they remove.  The path therefore stays OPT-IN (LUMO_TRACE_NM=1), like
// trace_flat.cuh, as the record of the experiment; the product path is trace.cuh's nested loops.
#pragma once
#include "trace_flat.cuh"

namespace lumo_dev {

#define LUMO_NM_MAX_NODES 64
#define LUMO_NM_MAX_OBJECTS 32

#define LUMO_NM_MAX_ITEMS (LUMO_NM_MAX_NODES + LUMO_NM_MAX_OBJECTS)
#define LUMO_NM_HEAVY_TRIS 64u           /* kd-trees with more triangles than this get their own, lane-refilled kernel */
struct NmPlan {                          // one object BVH in the reference's traversal order
    uint32_t n_nodes, root, obj_base, n_objects, n_items, pad[3];
    uint32_t node[LUMO_NM_MAX_NODES];    // node index (relative to root) at position k of the order
    int32_t parent[LUMO_NM_MAX_NODES];   // position of the parent, -1 for the root
    // the traversal as a flat list of items in visiting order: box test of node position k (item = k), or the test of an
    // object of the leaf at position k (item = 0x80000000 | k << 8 | local object index)
    uint32_t item[LUMO_NM_MAX_ITEMS];
    uint32_t heavy_bits;                 // bit o set: object o (local index) is heavy
};
struct NmRays {                          // per-ray state of one batch, indexed by position in the batch
    uint32_t cap;
    double* ctx;                         // [14][cap]: o, d, 1/d, sx, sy, sz, wz, kz
    double *tt, *bound, *qtmax;
    uint32_t *idx, *done, *have;
    unsigned long long* mask;
    double *ht, *hb0, *hb1, *hb2; uint32_t *hobj, *htri;
    uint32_t *queue, *sorted;            // compaction targets
    uint32_t* counters;                  // [NM_CNT_TOTAL]
};
#define NM_CNT_SEG 0                       /* queue sizes, one per heavy launch of a BVH pass */
#define NM_CNT_CUR (LUMO_NM_MAX_NODES)     /* their work cursors */
#define NM_CNT_TOTAL (2 * LUMO_NM_MAX_NODES)

__device__ __forceinline__ void nm_store_ctx(const NmRays& R, uint32_t i, const RayCtx& c) {
    double* p = R.ctx + i; const size_t s = R.cap;
    p[0] = c.r.o.x; p[s] = c.r.o.y; p[2 * s] = c.r.o.z; p[3 * s] = c.r.d.x; p[4 * s] = c.r.d.y; p[5 * s] = c.r.d.z;
    p[6 * s] = c.inv.x; p[7 * s] = c.inv.y; p[8 * s] = c.inv.z; p[9 * s] = c.q.sx; p[10 * s] = c.q.sy; p[11 * s] = c.q.sz; p[12 * s] = c.q.wz; p[13 * s] = (double)c.q.kz;
}
__device__ __forceinline__ void nm_load_ctx(const NmRays& R, uint32_t i, RayCtx& c) {
    const double* p = R.ctx + i; const size_t s = R.cap;
    c.r.o = d3(p[0], p[s], p[2 * s]); c.r.d = d3(p[3 * s], p[4 * s], p[5 * s]); c.inv = d3(p[6 * s], p[7 * s], p[8 * s]);
    c.q.sx = p[9 * s]; c.q.sy = p[10 * s]; c.q.sz = p[11 * s]; c.q.wz = p[12 * s]; c.q.kz = (int)p[13 * s];
}

template <class Source>
__global__ void __launch_bounds__(256) k_nm_setup(const __grid_constant__ NmRays R, const uint32_t* n_ptr, uint32_t n_cap, const Source src) {
    const uint32_t n = min(*n_ptr, n_cap);
    if (blockIdx.x == 0 && threadIdx.x < NM_CNT_TOTAL) R.counters[threadIdx.x] = 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Ray r; double tm; src.load(i, r, tm);
        RayCtx c; make_ctx(r, c);
        nm_store_ctx(R, i, c);
        R.tt[i] = tm; R.bound[i] = tm; R.qtmax[i] = tm; R.idx[i] = LUMO_NONE; R.done[i] = 0u; R.have[i] = 0u; R.mask[i] = 0ull;
    }
}
// Scene::hit / hit_light move on to Scene.lights: the bound is the object hit's distance (scene.rs:133-135) or the query's own
template <bool CLOSEST>
__global__ void __launch_bounds__(256) k_nm_begin_lights(const __grid_constant__ NmRays R, const uint32_t* n_ptr, uint32_t n_cap) {
    const uint32_t n = min(*n_ptr, n_cap);
    if (blockIdx.x == 0 && threadIdx.x < NM_CNT_TOTAL) R.counters[threadIdx.x] = 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double b = (CLOSEST && R.have[i]) ? R.ht[i] : R.qtmax[i];
        R.tt[i] = b; R.bound[i] = b; R.idx[i] = LUMO_NONE; R.mask[i] = 0ull;
    }
}
// Items [i0, i1] of the plan for every ray: box tests, and the objects of hit leaves.  Light objects (walls, spheres,
// single triangles) are tested in place with the reference's update; a heavy object can only be the last item of a segment:
// the rays that reach it are stream-compacted into the queue of slot `qslot` for k_nm_heavy.
template <bool CLOSEST>
__global__ void __launch_bounds__(128, 8) k_nm_segment(const __grid_constant__ DevScene S, const __grid_constant__ NmPlan P, const __grid_constant__ NmRays R,
                                                       const uint32_t* n_ptr, uint32_t n_cap, uint32_t i0, uint32_t i1, uint32_t qslot) {
    const uint32_t n = min(*n_ptr, n_cap);
    const uint32_t n_pad = (n + 31u) & ~31u, lane = threadIdx.x & 31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
        bool enq = false;
        if (i < n && !R.done[i]) {
            unsigned long long m = R.mask[i];
            RayCtx c; nm_load_ctx(R, i, c);
            double tt = R.tt[i]; uint32_t idx = R.idx[i]; bool done = false;
            for (uint32_t it = i0; it <= i1 && !done; it++) {
                const uint32_t item = P.item[it];
                if (!(item & 0x80000000u)) {                                // box test of node position k
                    const uint32_t k = item;
                    const int par = P.parent[k];
                    if (par >= 0 && !((m >> par) & 1ull)) continue;
                    const LumoTlasNode* node = S.tlas + P.root + P.node[k];
                    double t_start, t_end;
                    box_intersect(node->lo, node->hi, c.r.o, c.inv, t_start, t_end);
                    t_start = fmax(t_start, 0.0); t_end = fmin(t_end, tt);
                    if (t_start <= t_end) m |= 1ull << k;
                } else {                                                    // an object of the leaf at position k
                    const uint32_t k = (item >> 8) & 0xFFu, local = item & 0xFFu;
                    if (!((m >> k) & 1ull)) continue;
                    if ((P.heavy_bits >> local) & 1u) { enq = true; continue; }   // last item of the segment
                    const double t = object_hit_t<false>(S, S.objects[P.obj_base + local], c, 0.0, tt, nullptr);
                    if (t < tt) { if (CLOSEST) { tt = t; idx = local; } else done = true; }
                }
            }
            R.mask[i] = m; R.tt[i] = tt; R.idx[i] = idx;
            if (done) { R.done[i] = 1u; enq = false; }
        }
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, enq);
        if (b) {
            uint32_t base = 0;
            if (lane == (uint32_t)(__ffs(b) - 1)) base = atomicAdd(&R.counters[NM_CNT_SEG + qslot], (uint32_t)__popc(b));
            base = __shfl_sync(0xFFFFFFFFu, base, __ffs(b) - 1);
            if (enq) R.queue[base + __popc(b & ((1u << lane) - 1u))] = i;
        }
    }
}

// One heavy kd-tree against a dense queue of rays, lane-refilled: a lane that finishes its ray pulls the next queue
// entry at once, so the warp keeps 32 traversals in flight whatever their lengths (with one ray per lane for the whole
// launch, ncu showed 5-7 of 32 lanes busy: a few rays walk hundreds of nodes, most leave after a handful).  The loop is
// kd_hit's (trace.cuh, kdtree.rs:101-169) with its state in registers across rays.
//   GEO = false: Object::hit_t, then the BVH's update of the ray's bound / winner (CLOSEST) or "occluded" (first found);
//   GEO = true : Object::hit of the ray's winner (bvh.rs:366-369), result into the ray's hit record.
template <bool GEO, bool CLOSEST>
__global__ void __launch_bounds__(128, 8) k_nm_heavy(const __grid_constant__ DevScene S, const __grid_constant__ NmRays R, const uint32_t* queue, uint32_t qslot,
                                                     uint32_t obj_global, uint32_t obj_local) {
    const uint32_t n = R.counters[NM_CNT_SEG + qslot];
    uint32_t* cursor = &R.counters[NM_CNT_CUR + qslot];
    const LumoObject o = S.objects[obj_global];
    const LumoKdTree* tree = S.kd_trees + o.geom;
    const LumoTriVerts* tris = S.tri_verts + tree->tri_base;
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned FULL = 0xFFFFFFFFu;
    bool has = false, exhausted = false;
    uint32_t ray = 0;
    D3 lo = d3(0, 0, 0), linv = d3(0, 0, 0); RayTri lq; lq.kz = 2; lq.sx = lq.sy = lq.sz = 0.0; lq.wz = 1.0;
    uint32_t curr = 0, leaf_pos = 0, leaf_end = 0, k_idx = LUMO_NONE; int sp = 0; bool in_leaf = false;
    double t_start = 0.0, t_end = 0.0, t_hit = 0.0, tmax = 0.0;
    KdStackEntry stack[64];
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !has && !exhausted);
        if (idle) {
            uint32_t base = 0;
            const int leader = __ffs(idle) - 1;
            if ((int)lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(idle));
            base = __shfl_sync(FULL, base, leader);
            if (!has && !exhausted) {
                const uint32_t e = base + __popc(idle & ((1u << lane) - 1u));
                if (e >= n) exhausted = true;
                else {
                    ray = queue[e];
                    if (GEO || !R.done[ray]) {
                        RayCtx c; nm_load_ctx(R, ray, c);
                        if (o.inst >= 0) { const LumoInstance* I = S.instances + o.inst; Ray l; l.o = xf_point(I->inv, c.r.o); l.d = xf_dir(I->inv, c.r.d); make_ctx(l, c); }
                        lo = c.r.o; linv = c.inv; lq = c.q;
                        tmax = GEO ? R.bound[ray] : R.tt[ray];
                        sp = 0; t_hit = LUMO_INF; curr = tree->root; k_idx = LUMO_NONE; leaf_pos = leaf_end = 0; in_leaf = false;
                        box_intersect(tree->lo, tree->hi, lo, linv, t_start, t_end);
                        t_start = fmax(t_start, 0.0); t_end = fmin(t_end, tmax);
                        has = true;
                    }
                }
            }
        }
        if (__all_sync(FULL, exhausted && !has)) break;
        uint32_t tri = LUMO_NONE;
        bool finished = false;
#pragma unroll 1
        for (int step = 0; step < 24 && has; step++) {                     // advance to this lane's next triangle (kdtree.rs:117-160)
            if (leaf_pos < leaf_end) { tri = __ldg(S.kd_leaf + leaf_pos); leaf_pos++; break; }
            if (in_leaf) {
                if (sp == 0) { finished = true; break; }
                sp--;
                curr = stack[sp].node; t_start = stack[sp].t_start; t_end = stack[sp].t_end; in_leaf = false;
            }
            if (t_hit < t_start) { finished = true; break; }
            const double2 raw = __ldg(reinterpret_cast<const double2*>(S.kd_nodes + curr));
            const double point = raw.x;
            const uint32_t na = (uint32_t)(__double_as_longlong(raw.y) & 0xFFFFFFFFll);
            const uint32_t nb = (uint32_t)((unsigned long long)__double_as_longlong(raw.y) >> 32);
            if (nb & 0x80000000u) { leaf_pos = na; leaf_end = na + (nb & 0x7FFFFFFFu); in_leaf = true; }
            else {
                const int axis = (int)nb;
                const double o_a = axis == 0 ? lo.x : (axis == 1 ? lo.y : lo.z);
                const double i_a = axis == 0 ? linv.x : (axis == 1 ? linv.y : linv.z);
                const double t_split = (point - o_a) * i_a;
                const bool left_first = o_a < point || (o_a == point && i_a <= 0.0);
                const uint32_t first = left_first ? curr + 1 : na;
                const uint32_t second = left_first ? na : curr + 1;
                if (t_split > t_end || t_split <= 0.0) curr = first;
                else if (t_split < t_start) curr = second;
                else {
                    curr = first;
                    if (sp < 64) { stack[sp].node = second; stack[sp].t_start = t_split; stack[sp].t_end = t_end; sp++; }
                    t_end = t_split;
                }
            }
        }
        double found = LUMO_INF;                                           // any-hit result of this round, if the lane got one
        bool have_found = false;
        if (tri != LUMO_NONE) {
            TriHit th;
            Ray lr; lr.o = lo; lr.d = d3(0, 0, 0);
            const double t = tri_hit<false, false>(tris + tri, lr, lq, 0.0, t_end, th, nullptr) ? th.t : LUMO_INF;
            if (GEO) { if (t < t_end) { t_end = t; t_hit = t; k_idx = tri; } }
            else if (t < t_end) { found = t; have_found = true; finished = true; }
        }
        if (finished && has) {
            if (!GEO) {
                const double tt = R.tt[ray];                                // Object::hit_t returned `found` (INF if nothing was hit)
                if (have_found && found < tt) { if (CLOSEST) { R.tt[ray] = found; R.idx[ray] = obj_local; } else R.done[ray] = 1u; }
            } else if (k_idx != LUMO_NONE) {                                // kdtree.rs:162-168: full test of the closest candidate
                TriHit th;
                Ray lr; lr.o = lo; lr.d = d3(0, 0, 0);
                if (tri_hit<true, false>(tris + k_idx, lr, lq, 0.0, tmax, th, nullptr)) {
                    R.ht[ray] = th.t; R.hb0[ray] = th.bary.x; R.hb1[ray] = th.bary.y; R.hb2[ray] = th.bary.z; R.hobj[ray] = obj_global; R.htri[ray] = k_idx; R.have[ray] = 1u;
                }
            }
            has = false;
        }
    }
}

// The winners' full intersection (bvh.rs:366-369): light objects in place, heavy ones queued per object for k_nm_heavy<GEO>
__global__ void __launch_bounds__(128, 8) k_nm_winners(const __grid_constant__ DevScene S, const __grid_constant__ NmPlan P, const __grid_constant__ NmRays R,
                                                       const uint32_t* n_ptr, uint32_t n_cap, uint32_t heavy_local, uint32_t qslot, uint32_t do_light) {
    const uint32_t n = min(*n_ptr, n_cap);
    const uint32_t n_pad = (n + 31u) & ~31u, lane = threadIdx.x & 31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
        bool enq = false;
        if (i < n) {
            const uint32_t o = R.idx[i];
            if (o != LUMO_NONE) {
                if (o == heavy_local) enq = true;
                else if (do_light && !((P.heavy_bits >> o) & 1u)) {
                    RayCtx c; nm_load_ctx(R, i, c);
                    HitRec h;
                    if (object_hit<false>(S, S.objects[P.obj_base + o], c, 0.0, R.bound[i], h, nullptr)) {
                        R.ht[i] = h.t; R.hb0[i] = h.bary.x; R.hb1[i] = h.bary.y; R.hb2[i] = h.bary.z; R.hobj[i] = P.obj_base + o; R.htri[i] = h.tri; R.have[i] = 1u;
                    }
                }
            }
        }
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, enq);
        if (b) {
            uint32_t base = 0;
            if (lane == (uint32_t)(__ffs(b) - 1)) base = atomicAdd(&R.counters[NM_CNT_SEG + qslot], (uint32_t)__popc(b));
            base = __shfl_sync(0xFFFFFFFFu, base, __ffs(b) - 1);
            if (enq) R.sorted[base + __popc(b & ((1u << lane) - 1u))] = i;
        }
    }
}
template <bool CLOSEST, class Sink>
__global__ void __launch_bounds__(256) k_nm_finish(const __grid_constant__ NmRays R, const uint32_t* n_ptr, uint32_t n_cap, Sink sink, unsigned long long* total) {
    const uint32_t n = min(*n_ptr, n_cap);
    if (total && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(total, (unsigned long long)n);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        FlatResult res; res.t = LUMO_INF;
        if (CLOSEST) {
            res.hit = R.have[i] != 0u;
            res.h.t = R.ht[i]; res.h.bary = d3(R.hb0[i], R.hb1[i], R.hb2[i]); res.h.obj = R.hobj[i]; res.h.tri = R.htri[i];
        } else res.hit = R.done[i] != 0u;
        sink.store(i, res);
    }
}

}  // namespace lumo_dev
