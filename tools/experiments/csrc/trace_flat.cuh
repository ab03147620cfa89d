// Persistent, lane-refilled form of the traversal in trace.cuh.
//
// trace.cuh phrases Scene::hit / hit_light / hit_t as nested routines (object BVH -> kd-tree -> triangle);
// run one ray per lane, a warp then waits for its slowest ray and, inside every loop, for its slowest lane
// (ncu: ~6 of 32 lanes active per issued instruction on incoherent rays).  Here the same per-ray sequence
// of nodes, objects, triangles and comparisons is driven by a per-lane state machine inside ONE loop:
//   * a lane that finishes its ray immediately pulls the next one from the queue (work stealing at lane
//     granularity), so a warp always carries 32 rays;
//   * all lanes first advance — a bounded number of node / stack / object-setup micro-steps — until they
//     hold the next triangle to test, then the lanes that hold one run the Woop test together.
// Per ray nothing changes: node order, leaf order, strict '<' updates, any-hit distances deciding the
// object BVH winner, the winner's second (full) intersection, objects before lights, and
// every floating-point expression is the one in trace.cuh.  The bit-exact parity tests run through this
// code when LUMO_TRACE_FLAT=1 (tests/test_trace_parity.py::test_flat_traversal_is_bit_exact_too).
//
// MEASURED (B200, bunny stand-in, incoherent bounce rays): correct, but 2x SLOWER than the nested while-while
// form of trace.cuh (54.9 vs 26.5 ms per 14.9 M rays) — lanes still average 6.1 of 32 active per issued
// instruction (ncu, profiles/), because with 32 independent rays per warp the lanes spread over the
// states of the machine instead of over the iterations of a loop, at half the occupancy (128 registers).
// Kept as an opt-in experiment, not the default path.
#pragma once
#include "trace.cuh"

namespace lumo_dev {

enum { FQ_CLOSEST = 0, FQ_OCCLUDED = 1, FQ_FIRST_T = 2 };
enum { FS_IDLE = 0, FS_TLAS = 1, FS_KD = 2, FS_EXHAUSTED = 3 };
#define LUMO_FLAT_T_STEPS 6
#define LUMO_FLAT_K_STEPS 16

struct FlatResult { bool hit; HitRec h; double t; };   // CLOSEST: hit + h; OCCLUDED: hit; FIRST_T: t

// Source / Sink are small structs with
//   bool  Source::load(uint64 item, Ray& r, double& t_max)      (item < count)
//   void  Sink::store(uint64 item, const FlatResult& res)
template <int Q, class Source, class Sink>
__device__ __forceinline__ void flat_trace(const DevScene& S, unsigned long long count, unsigned long long* cursor32, uint32_t* cursor, const Source& src, Sink& sink) {
    // ---- per-lane state ----
    int st = FS_IDLE;
    unsigned long long item = 0;
    RayCtx Wc;                                 // world frame (kept for object set-up; lives in local memory)
    D3 lo = d3(0, 0, 0), linv = d3(0, 0, 0); RayTri lq; lq.kz = 2; lq.sx = lq.sy = lq.sz = 0.0; lq.wz = 1.0;   // current object's frame: origin, 1/dir, Woop constants
    // object-BVH pass
    int pass = 0;                              // 0: Scene.objects BVH, 2: Scene.lights BVH (1 / 3: full intersection of the winner)
    uint32_t t_root = 0, t_base = 0, t_curr = 0, t_leafpos = 0, t_leafend = 0, t_idx = LUMO_NONE, t_obj = LUMO_NONE;
    int t_sp = 0; bool t_needpop = false;
    double tt = 0.0, p_bound = 0.0, q_tmax = 0.0;   // running bound of the pass; the bound the pass started with; the query's own t_max
    uint32_t tstack[64];
    // kd pass
    uint32_t k_curr = 0, k_leafpos = 0, k_leafend = 0, k_idx = LUMO_NONE, k_tri = LUMO_NONE;
    int k_sp = 0; bool k_inleaf = false, k_geo = false;
    double k_tstart = 0.0, k_tend = 0.0, k_thit = 0.0, k_tmax = 0.0;
    const LumoTriVerts* k_tris = nullptr;
    KdStackEntry kstack[64];
    // result
    FlatResult res; res.hit = false; res.t = LUMO_INF;
    Ray lray;                                  // local-frame ray (sphere / loose triangle / GEO re-test need the direction)

    const unsigned FULL = 0xFFFFFFFFu;
    const uint32_t lane = threadIdx.x & 31u;
    (void)lane;

    // -- helpers as lambdas over the state --
    auto begin_pass = [&](int p, double bound) {
        pass = p;
        t_root = p == 0 ? 0u : S.P.lights_root; t_base = p == 0 ? 0u : S.P.n_objects;
        t_curr = 0; t_sp = 0; t_idx = LUMO_NONE; t_leafpos = t_leafend = 0; t_needpop = false; tt = bound; p_bound = bound;
        st = FS_TLAS;
    };
    auto finish_ray = [&]() { sink.store(item, res); st = FS_IDLE; };
    // what follows the object-BVH pass `pass` once its winner (if any) has been fully intersected / decided
    auto after_pass = [&]() {
        if (pass == 0 && S.P.n_lights) {
            double bound;
            if (Q == FQ_CLOSEST) bound = res.hit ? res.h.t : q_tmax;          // scene.rs:133-135
            else if (Q == FQ_OCCLUDED) bound = q_tmax;
            else bound = res.t;                                               // scene.rs:158-160
            begin_pass(2, bound);
        } else finish_ray();
    };
    // result of Object::hit_t for the object under test (any-hit distance), bvh.rs:346-351
    auto tlas_update = [&](double t) {
        if (Q == FQ_CLOSEST) { if (t < tt) { tt = t; t_idx = t_obj; } st = FS_TLAS; }
        else if (t < tt) {                                                    // first object found decides (bvh.rs:349-351, :371-374)
            if (Q == FQ_OCCLUDED) { res.hit = true; finish_ray(); }
            else { res.t = fmin(res.t, t); after_pass(); }
        } else st = FS_TLAS;
    };
    // frame + kd / analytic set-up for object `oi` of the current BVH; geo = the winner's full intersection
    auto begin_object = [&](uint32_t oi, bool geo) {
        const LumoObject o = S.objects[t_base + oi];
        const double o_tmax = geo ? p_bound : tt;            // bvh.rs:346 (running bound) vs :368 (the caller's t_max)
        if (o.inst >= 0) {
            const LumoInstance* I = S.instances + o.inst;
            lray.o = xf_point(I->inv, Wc.r.o); lray.d = xf_dir(I->inv, Wc.r.d);
            RayCtx lc; make_ctx(lray, lc);
            lo = lc.r.o; linv = lc.inv; lq = lc.q;
        } else { lray = Wc.r; lo = Wc.r.o; linv = Wc.inv; lq = Wc.q; }
        if (o.kind == LOBJ_KD || o.kind == LOBJ_RECT) {
            const LumoKdTree* tree = S.kd_trees + o.geom;
            k_sp = 0; k_thit = LUMO_INF; k_curr = tree->root; k_idx = LUMO_NONE; k_leafpos = k_leafend = 0; k_inleaf = false; k_geo = geo; k_tmax = o_tmax;
            box_intersect(tree->lo, tree->hi, lo, linv, k_tstart, k_tend);
            k_tstart = fmax(k_tstart, 0.0); k_tend = fmin(k_tend, o_tmax);
            k_tris = S.tri_verts + tree->tri_base;
            st = FS_KD;
            return;
        }
        // analytic objects are decided on the spot
        if (!geo) {
            double t;
            if (o.kind == LOBJ_SPHERE) t = sphere_hit_t(S.spheres[o.geom].radius, lray, 0.0, o_tmax);
            else { TriHit th; t = tri_hit<false, false>(S.tri_verts + o.geom, lray, lq, 0.0, o_tmax, th, nullptr) ? th.t : LUMO_INF; }
            tlas_update(t);
        } else {
            bool ok = false; HitRec h; h.obj = t_base + oi; h.tri = 0; h.bary = d3(0, 0, 0); h.t = LUMO_INF;
            if (o.kind == LOBJ_SPHERE) { const double t = sphere_hit(S.spheres[o.geom].radius, lray, 0.0, o_tmax); if (t < LUMO_INF) { h.t = t; ok = true; } }
            else { TriHit th; if (tri_hit<true, false>(S.tri_verts + o.geom, lray, lq, 0.0, o_tmax, th, nullptr)) { h.t = th.t; h.bary = th.bary; ok = true; } }
            if (ok) { res.hit = true; res.h = h; }
            after_pass();
        }
    };
    // the object-BVH pass has visited everything (closest mode): intersect the winner fully, bvh.rs:366-369
    auto tlas_done = [&]() {
        if (Q == FQ_CLOSEST && t_idx != LUMO_NONE) { t_obj = t_idx; begin_object(t_idx, true); }
        else after_pass();
    };

    for (;;) {
        // ---- refill: idle lanes pull the next ray (one atomic per warp) ----
        {
            const unsigned idle = __ballot_sync(FULL, st == FS_IDLE);
            if (idle) {
                unsigned long long base = 0;
                const int leader = __ffs(idle) - 1;
                if ((int)lane == leader) base = cursor32 ? atomicAdd(cursor32, (unsigned long long)__popc(idle)) : (unsigned long long)atomicAdd(cursor, (uint32_t)__popc(idle));
                base = __shfl_sync(FULL, base, leader);
                if (st == FS_IDLE) {
                    item = base + (unsigned long long)__popc(idle & ((1u << lane) - 1u));
                    if (item < count) {
                        Ray r; double tm;
                        src.load(item, r, tm);
                        make_ctx(r, Wc);
                        q_tmax = tm;
                        res.hit = false; res.t = LUMO_INF; res.h.obj = LUMO_NONE; res.h.tri = LUMO_NONE; res.h.t = LUMO_INF; res.h.bary = d3(0, 0, 0);
                        begin_pass(0, Q == FQ_FIRST_T ? LUMO_INF : tm);
                    } else st = FS_EXHAUSTED;
                }
            }
            if (__all_sync(FULL, st == FS_EXHAUSTED)) break;
        }
        // ---- advance ----
        // Lanes walking the object BVH (box tests, instance transforms, kd set-up: long steps) and lanes walking a
        // kd-tree (short steps) are advanced in separate phases, so that the short steps never wait for the long ones.
        k_tri = LUMO_NONE;
        auto kd_step = [&]() {                                               // kdtree.rs:117-160, one step
                if (k_leafpos < k_leafend) { k_tri = __ldg(S.kd_leaf + k_leafpos); k_leafpos++; return; }
                bool finished = false;
                if (k_inleaf) {
                    if (k_sp == 0) finished = true;
                    else { k_sp--; k_curr = kstack[k_sp].node; k_tstart = kstack[k_sp].t_start; k_tend = kstack[k_sp].t_end; k_inleaf = false; }
                }
                if (!finished && k_thit < k_tstart) finished = true;
                if (!finished) {
                    const double2 raw = __ldg(reinterpret_cast<const double2*>(S.kd_nodes + k_curr));
                    const double point = raw.x;
                    const uint32_t na = (uint32_t)(__double_as_longlong(raw.y) & 0xFFFFFFFFll);
                    const uint32_t nb = (uint32_t)((unsigned long long)__double_as_longlong(raw.y) >> 32);
                    if (nb & 0x80000000u) { k_leafpos = na; k_leafend = na + (nb & 0x7FFFFFFFu); k_inleaf = true; }
                    else {
                        const int axis = (int)nb;
                        const double o_a = axis == 0 ? lo.x : (axis == 1 ? lo.y : lo.z);
                        const double i_a = axis == 0 ? linv.x : (axis == 1 ? linv.y : linv.z);
                        const double t_split = (point - o_a) * i_a;
                        const bool left_first = o_a < point || (o_a == point && i_a <= 0.0);
                        const uint32_t first = left_first ? k_curr + 1 : na;
                        const uint32_t second = left_first ? na : k_curr + 1;
                        if (t_split > k_tend || t_split <= 0.0) k_curr = first;
                        else if (t_split < k_tstart) k_curr = second;
                        else {
                            k_curr = first;
                            if (k_sp < 64) { kstack[k_sp].node = second; kstack[k_sp].t_start = t_split; kstack[k_sp].t_end = k_tend; k_sp++; }
                            k_tend = t_split;
                        }
                    }
                } else if (!k_geo) tlas_update(LUMO_INF);                    // no triangle of this object was hit
                else {                                                        // kdtree.rs:162-168: full test of the closest candidate
                    bool ok = false; HitRec h;
                    if (k_idx != LUMO_NONE) {
                        TriHit th;
                        Ray lr; lr.o = lo; lr.d = lray.d;
                        if (tri_hit<true, false>(k_tris + k_idx, lr, lq, 0.0, k_tmax, th, nullptr)) { h.t = th.t; h.tri = k_idx; h.bary = th.bary; h.obj = t_base + t_obj; ok = true; }
                    }
                    if (ok) { res.hit = true; res.h = h; }
                    after_pass();
                }
        };
        auto tlas_step = [&]() {                                             // bvh.rs:326-360, one step
                if (t_leafpos < t_leafend) { t_obj = S.tlas_leaf[t_leafpos]; t_leafpos++; begin_object(t_obj, false); return; }
                if (t_needpop) {
                    if (t_sp == 0) { tlas_done(); return; }
                    t_curr = tstack[--t_sp]; t_needpop = false;
                }
                const LumoTlasNode* node = S.tlas + t_root + t_curr;
                double t_start, t_end;
                box_intersect(node->lo, node->hi, Wc.r.o, Wc.inv, t_start, t_end);
                t_start = fmax(t_start, 0.0); t_end = fmin(t_end, tt);
                if (t_start <= t_end) {
                    const uint32_t cnt = node->count;
                    if (cnt == 0) {
                        const uint32_t right = node->right;
                        t_curr += 1;
                        if (right != LUMO_NONE && t_sp < 64) tstack[t_sp++] = right;
                        return;
                    }
                    t_leafpos = node->first; t_leafend = t_leafpos + cnt;
                }
                t_needpop = true;
        };
#pragma unroll 1
        for (int step = 0; step < LUMO_FLAT_T_STEPS; step++) {
            if (!__any_sync(FULL, st == FS_TLAS)) break;
            if (st == FS_TLAS) tlas_step();
        }
#pragma unroll 1
        for (int step = 0; step < LUMO_FLAT_K_STEPS; step++) {
            const bool need = st == FS_KD && k_tri == LUMO_NONE;
            if (!__any_sync(FULL, need)) break;
            if (need) kd_step();
        }
        // ---- triangle test for the lanes that hold one (kdtree.rs:124-130) ----
        if (k_tri != LUMO_NONE) {
            TriHit th;
            Ray lr; lr.o = lo; lr.d = lray.d;
            const double t = tri_hit<false, false>(k_tris + k_tri, lr, lq, 0.0, k_tend, th, nullptr) ? th.t : LUMO_INF;
            if (k_geo) { if (t < k_tend) { k_tend = t; k_thit = t; k_idx = k_tri; } }
            else if (t < k_tend) tlas_update(t);                             // any-hit: the first triangle found ends the object's test
        }
    }
}

}  // namespace lumo_dev
