// Device-side traversal and intersection, f64, compiled with -fmad=false.
//
// Faithful two-level traversal: object BVH (left child = index+1, right pushed, no near/far
// ordering; src/tracer/object/bvh.rs:315-362) -> optional instance transform with un-normalised
// direction (instance.rs:80-105, ray.rs:24-30) -> kd-tree front-to-back with a 64-entry
// (node, t_start, t_end) stack (kdtree.rs:101-169) -> Woop watertight triangle test
// (triangle.rs:63-187).  Node, leaf and triangle order in the blob are lumo's, and every
// floating-point expression keeps the reference's operation order, so hit ids, t and
// barycentrics are bit-identical to the CPU path (tests/test_trace_parity.py).
//
// Per-ray work that the reference redoes per triangle (kz selection, axis permutation, shear
// constants) is hoisted out of the leaf loop — same values, computed once.
#pragma once
#include "../common/scene_blob.h"
#include <cuda_runtime.h>
#include <math.h>

namespace lumo_dev {

struct DevScene {
    const LumoTlasNode* tlas; const uint32_t* tlas_leaf;
    const LumoObject* objects; const LumoInstance* instances;
    const LumoKdTree* kd_trees; const LumoKdNode* kd_nodes; const uint32_t* kd_leaf;
    const LumoTriVerts* tri_verts; const LumoTriShade* tri_shade;
    const double* normals; const double* uvs;
    const LumoRect* rects; const LumoSphere* spheres;
    const LumoMaterial* materials; const double* tables; const LumoLight* lights;
    const LumoTexture* textures; const float* tex_pixels; const double* tex_f64;
    const LumoAhNode* ah_nodes; const LumoAhPrim* ah_prims; const uint32_t* obj_path_off; const uint32_t* obj_path;   // occlude.cuh
    LumoSceneParams P;
};

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator*(D3 a, D3 b) { return d3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return d3(s * a.x, s * a.y, s * a.z); }
// (Measured and dropped: one refined reciprocal shared by the three quotients — nvcc's own MUFU.RCP64H + 5 DFMA sequence and
// quotient step, bit-identical to `/` on 3.6 M operand pairs incl. subnormals and all-ones significands.  28 fewer instructions
// per vector division, yet the shade class went from 209.3 to 211.8 ms on bistro 4 spp: the divisions are not what it waits for.)
__device__ __forceinline__ D3 operator/(D3 a, double s) { return d3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ D3 operator-(D3 a) { return d3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 cross(D3 a, D3 b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double length(D3 a) { return sqrt(fmax(dot(a, a), 0.0)); }
__device__ __forceinline__ D3 normalize(D3 a) { return a / length(a); }
__device__ __forceinline__ D3 vabs(D3 a) { return d3(fabs(a.x), fabs(a.y), fabs(a.z)); }
__device__ __forceinline__ double min_element(D3 a) { return fmin(a.x, fmin(a.y, a.z)); }
__device__ __forceinline__ double max_element(D3 a) { return fmax(a.x, fmax(a.y, a.z)); }
__device__ __forceinline__ double dist2(D3 a, D3 b) { D3 d = a - b; return dot(d, d); }

#define LUMO_INF (__longlong_as_double(0x7FF0000000000000LL))
#define LUMO_EPS 1e-10
// gamma(n) = n*eps / (1 - n*eps), eps = 2^-52 (efloat.rs:5-8) — evaluated in f64 exactly as the reference does
__device__ __forceinline__ double gamma_n(int n) { const double e = 2.220446049250313e-16; return ((double)n * e) / (1.0 - (double)n * e); }

struct Ray { D3 o, d; };

// Per-ray constants of the Woop test (triangle.rs:66-92)
struct RayTri { int kz; double sx, sy, sz, wz; };
__device__ __forceinline__ D3 permute(D3 v, int kz) { return kz == 0 ? d3(v.y, v.z, v.x) : (kz == 1 ? d3(v.z, v.x, v.y) : v); }
__device__ __forceinline__ RayTri ray_tri_setup(const Ray& r) {
    D3 a = vabs(r.d);
    RayTri q;
    q.kz = (a.x > a.y && a.x > a.z) ? 0 : (a.y > a.z ? 1 : 2);
    D3 wi = permute(r.d, q.kz);
    q.sx = -wi.x / wi.z; q.sy = -wi.y / wi.z; q.sz = 0.0 / wi.z; q.wz = wi.z;
    return q;
}

// Everything the traversal derives from a ray alone: reciprocal direction (bvh.rs:317, kdtree.rs:103 —
// recomputed by the reference per BVH / kd-tree visit) and the Woop constants (triangle.rs:66-92 —
// recomputed per triangle).  Same values, computed once per (ray, coordinate frame).
struct RayCtx { Ray r; D3 inv; RayTri q; };
__device__ __noinline__ void make_ctx(const Ray& r, RayCtx& c) {
    c.r = r;
    c.inv = d3(1.0 / r.d.x, 1.0 / r.d.y, 1.0 / r.d.z);
    c.q = ray_tri_setup(r);
}

struct TriHit { double t; D3 bary; };

// Visit counters of one thread (same sites as the oracle's: SURVEY §8d byte formula).  Compiled in
// only for CNT=true instantiations (the untimed counting pass of bench.py and the parity tests).
struct Counters { unsigned long long tlas, inst, kd, leaf, tri, sphere; };
#define LUMO_CNT(field) do { if (CNT) c->field++; } while (0)

// 80-byte triangle record as five 16-byte loads through the read-only path
__device__ __forceinline__ void load_tri(const LumoTriVerts* tv, D3& a, D3& b, D3& c) {
    const double2* p = reinterpret_cast<const double2*>(tv);
    double2 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3), v4 = __ldg(p + 4);
    a = d3(v0.x, v0.y, v1.x); b = d3(v1.y, v2.x, v2.y); c = d3(v3.x, v3.y, v4.x);
}

// Triangle::_hit<GEO> (triangle.rs:63-187).  GEO=false: distance only; GEO=true adds the
// conservative delta_t rejection and the barycentrics.
template <bool GEO, bool CNT>
__device__ __forceinline__ bool tri_hit(const LumoTriVerts* tv, const Ray& r, const RayTri& q, double t_min, double t_max, TriHit& out, Counters* c) {
    LUMO_CNT(tri);
    D3 A, B, C; load_tri(tv, A, B, C);
    D3 at = permute(A - r.o, q.kz), bt = permute(B - r.o, q.kz), ct = permute(C - r.o, q.kz);
    at = d3(at.x + q.sx * at.z, at.y + q.sy * at.z, at.z + q.sz * at.z);
    bt = d3(bt.x + q.sx * bt.z, bt.y + q.sy * bt.z, bt.z + q.sz * bt.z);
    ct = d3(ct.x + q.sx * ct.z, ct.y + q.sy * ct.z, ct.z + q.sz * ct.z);
    D3 e = d3(bt.x * ct.y - bt.y * ct.x, ct.x * at.y - ct.y * at.x, at.x * bt.y - at.y * bt.x);
    if (min_element(e) < 0.0 && max_element(e) > 0.0) return false;
    double det = e.x * 1.0 + e.y * 1.0 + e.z * 1.0;
    if (det == 0.0) return false;
    double t_scaled = (e.x * at.z + e.y * bt.z + e.z * ct.z) / q.wz;
    bool b1 = det < 0.0 && (t_scaled > t_min * det || t_scaled < t_max * det);
    bool b2 = det > 0.0 && (t_scaled < t_min * det || t_scaled > t_max * det);
    if (b1 || b2) return false;
    double t = t_scaled / det;
    out.t = t;
    if (!GEO) return true;
    double max_z = fmax(fmax(fabs(at.z), fabs(bt.z)), fabs(ct.z));
    double delta_z = gamma_n(3) * max_z;
    double max_y = fmax(fmax(fabs(at.y), fabs(bt.y)), fabs(ct.y));
    double delta_y = gamma_n(5) * (max_y + max_z);
    double max_x = fmax(fmax(fabs(at.x), fabs(bt.x)), fabs(ct.x));
    double delta_x = gamma_n(5) * (max_x + max_z);
    double delta_e = 2.0 * (gamma_n(2) * max_x * max_y + delta_y * max_x + delta_x * max_y);
    double max_e = fmax(fmax(fabs(e.x), fabs(e.y)), fabs(e.z));
    double delta_t = 3.0 * (gamma_n(3) * max_e * max_z + delta_e * max_z + delta_z * max_e) / fabs(det);
    if (t <= t_min + delta_t) return false;
    out.bary = e / det;
    return true;
}

// AaBoundingBox::intersect (aabb.rs:33-44)
__device__ __forceinline__ void box_intersect(const double* lo, const double* hi, D3 o, D3 inv, double& t_start, double& t_end) {
    D3 rmin = d3((lo[0] - o.x) * inv.x, (lo[1] - o.y) * inv.y, (lo[2] - o.z) * inv.z);
    D3 rmax = d3((hi[0] - o.x) * inv.x, (hi[1] - o.y) * inv.y, (hi[2] - o.z) * inv.z);
    D3 ts = d3(fmin(rmin.x, rmax.x), fmin(rmin.y, rmax.y), fmin(rmin.z, rmax.z));
    D3 te = d3(fmax(rmax.x, rmin.x), fmax(rmax.y, rmin.y), fmax(rmax.z, rmin.z));
    t_start = max_element(ts);
    t_end = min_element(te) * (1.0 + 2.0 * gamma_n(3));
}

// KdTree::_hit<GEO> (kdtree.rs:101-169).
// GEO=false: returns the distance of the first triangle found with t < t_end (any-hit), INF if none.
// GEO=true : closest triangle by the reference's rules, then the full test of the winner
//            (kdtree.rs:165) — returns false if that fails.
struct KdStackEntry { uint32_t node; double t_start, t_end; };

// The loop is the reference's, re-phrased for SIMT ("while-while"): each lane first walks inner nodes /
// pops its stack until it holds the next triangle to test, then all lanes of the warp that hold one run
// the triangle test together.  Per lane the sequence of nodes, triangles and comparisons is exactly
// kdtree.rs:117-160.
// STACK: capacity of the (node, t_start, t_end) stack.  The reference's is 64 (kdtree.rs:110); the kd-trees of
// Rectangle lights have two triangles, and the shading kernels that intersect one chosen light use a small one
// so that their per-thread frame stays small (lumo_gpu_scene_upload checks that light kd-trees fit).
#define LUMO_LIGHT_KD_STACK 8
// ROUND = 0: "while-while" proper — every lane walks until it holds its next triangle, then the warp runs the triangle
// test together.  ROUND = K > 0 bounds the walk to K steps per round, so lanes that already hold a triangle idle less
// while the test runs with fewer lanes.  Measured on one B200, same box, 0 vs 2: the register-bounded wave kernels gain
// (dragon 132 -> 146 Mrays/s, bunny +2 %, bistro and conference unchanged), the batch kernel loses (micro closest-hit
// 1135 / 718 -> 1030 / 656 Mrays/s primary / incoherent).  So the wave kernels instantiate LUMO_WAVE_KD_ROUND and
// everything else the default.
#ifndef LUMO_KD_ROUND
#define LUMO_KD_ROUND 0
#endif
#ifndef LUMO_KD_AXIS_LMEM
#define LUMO_KD_AXIS_LMEM 1
#endif
#ifndef LUMO_WAVE_KD_ROUND
#define LUMO_WAVE_KD_ROUND 2
#endif
#ifndef LUMO_WAVE_TRACE_BLOCKS
// Resident CTAs per SM the wave traversal kernels are compiled for (and their persistent grid).  Same-box sweep on B200,
// trace + occlusion ms for bunny 4 spp / dragon 2 spp: 4 CTAs (128 regs) 39.5 / 72.4, 6: 33.8 / 62.4, 8 (64 regs): 32.4 / 57.9,
// 10 (48 regs): 34.0 / 60.8, 12: 37.3 / 66.4, 16 (32 regs): 51.9 / 92.7 — latency-bound, so occupancy wins until spills take over.
#define LUMO_WAVE_TRACE_BLOCKS 8
#endif
template <bool GEO, bool CNT, int STACK = 64, int ROUND = LUMO_KD_ROUND>
__device__ __forceinline__ bool kd_hit(const DevScene& S, const LumoKdTree* tree, const RayCtx& ctx, double t_min, double t_max,
                                    double& t_out, uint32_t& tri_out, D3& bary_out, Counters* c, double* t_first = nullptr) {
    const Ray r = ctx.r;
    const RayTri q = ctx.q;
    const D3 inv = ctx.inv;
    KdStackEntry stack[STACK];
    // Same-box A/B on B200: indexing the origin / reciprocal direction by the split axis through local memory makes the batch
    // kernels (112 registers, unbounded walk) 16 % faster (micro 1134 / 717 -> 1311 / 836 Mrays/s) and the 64-register wave
    // kernels 1 % slower, so it follows the walk form: on where ROUND == 0.
    constexpr bool AXIS_LMEM = LUMO_KD_AXIS_LMEM && ROUND == 0;
    double oa[AXIS_LMEM ? 3 : 1], ia[AXIS_LMEM ? 3 : 1];
    if (AXIS_LMEM) { oa[0] = r.o.x; oa[AXIS_LMEM ? 1 : 0] = r.o.y; oa[AXIS_LMEM ? 2 : 0] = r.o.z; ia[0] = inv.x; ia[AXIS_LMEM ? 1 : 0] = inv.y; ia[AXIS_LMEM ? 2 : 0] = inv.z; }
    int sp = 0;
    double t_hit = LUMO_INF;
    uint32_t curr = tree->root;
    uint32_t idx = LUMO_NONE;
    double t_start, t_end;
    box_intersect(tree->lo, tree->hi, r.o, inv, t_start, t_end);
    t_start = fmax(t_start, t_min); t_end = fmin(t_end, t_max);
    const LumoTriVerts* tris = S.tri_verts + tree->tri_base;
    uint32_t leaf_pos = 0, leaf_end = 0;     // pending entries of the current leaf in LSEC_KD_LEAF
    bool in_leaf = false;                    // the current node was a leaf: pop once its entries are done
    bool finished = false;
    for (;;) {
        uint32_t tri = LUMO_NONE;
#pragma unroll 1
        for (int step = 0; ROUND == 0 || step < ROUND; step++) {   // advance towards the next triangle of this lane (ROUND = 0: all the way)
            if (leaf_pos < leaf_end) { tri = __ldg(S.kd_leaf + leaf_pos); leaf_pos++; LUMO_CNT(leaf); break; }
            if (in_leaf) {
                if (sp == 0) { finished = true; break; }
                sp--;
                curr = stack[sp].node; t_start = stack[sp].t_start; t_end = stack[sp].t_end;
                in_leaf = false;
            }
            if (t_hit < t_start) { finished = true; break; }
            LUMO_CNT(kd);
            const double2 raw = __ldg(reinterpret_cast<const double2*>(S.kd_nodes + curr));   // 16-byte node: one vector load
            const double point = raw.x;
            const uint32_t na = (uint32_t)(__double_as_longlong(raw.y) & 0xFFFFFFFFll);
            const uint32_t nb = (uint32_t)((unsigned long long)__double_as_longlong(raw.y) >> 32);
            if (nb & 0x80000000u) { leaf_pos = na; leaf_end = na + (nb & 0x7FFFFFFFu); in_leaf = true; }
            else {
                const int axis = (int)nb;
                // the ray's component along the split axis: two local-memory loads (AXIS_LMEM) or ~10 selects
                const double o_a = AXIS_LMEM ? oa[AXIS_LMEM ? axis : 0] : (axis == 0 ? r.o.x : (axis == 1 ? r.o.y : r.o.z));
                const double i_a = AXIS_LMEM ? ia[AXIS_LMEM ? axis : 0] : (axis == 0 ? inv.x : (axis == 1 ? inv.y : inv.z));
                const double t_split = (point - o_a) * i_a;
                const bool left_first = o_a < point || (o_a == point && i_a <= 0.0);
                const uint32_t first = left_first ? curr + 1 : na;
                const uint32_t second = left_first ? na : curr + 1;
                if (t_split > t_end || t_split <= 0.0) curr = first;
                else if (t_split < t_start) curr = second;
                else {
                    curr = first;
                    if (sp < STACK) { stack[sp].node = second; stack[sp].t_start = t_split; stack[sp].t_end = t_end; sp++; }
                    t_end = t_split;
                }
            }
        }
        if (tri == LUMO_NONE) { if (ROUND == 0 || finished) break; continue; }
        TriHit th;
        const double t = tri_hit<false, CNT>(tris + tri, r, q, t_min, t_end, th, c) ? th.t : LUMO_INF;
        // t_first: the distance the any-hit form of this walk (GEO = false) would have returned — both forms are the same walk up to the first accepted triangle
        if (GEO) { if (t < t_end) { if (t_first && idx == LUMO_NONE) *t_first = t; t_end = t; t_hit = t; idx = tri; } }
        else { if (t < t_end) { t_out = t; return true; } }
    }
    if (!GEO) { t_out = LUMO_INF; return false; }
    if (idx == LUMO_NONE) return false;
    TriHit th;
    if (!tri_hit<true, CNT>(tris + idx, r, q, t_min, t_max, th, c)) return false;
    t_out = th.t; tri_out = idx; bary_out = th.bary;
    return true;
}

// Ray::transform::<false> (ray.rs:24-30) through the affine rows of the inverse matrix.
// The reference multiplies a full Vec4 and divides by w; w is exactly 1 (points) or 0 (dirs) for
// affine instances, so only the + m[3]*w term is kept to reproduce signed zeros.
__device__ __forceinline__ D3 xf_point(const double* m, D3 p) {
    return d3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3] * 1.0,
              m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7] * 1.0,
              m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11] * 1.0) / 1.0;
}
__device__ __forceinline__ D3 xf_dir(const double* m, D3 v) {
    return d3(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * 0.0,
              m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * 0.0,
              m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * 0.0);
}
template <bool CNT>
__device__ __forceinline__ Ray to_local(const DevScene& S, const LumoObject& o, const Ray& r, Counters* c) {
    if (o.inst < 0) return r;
    LUMO_CNT(inst);
    const LumoInstance* I = S.instances + o.inst;
    Ray l; l.o = xf_point(I->inv, r.o); l.d = xf_dir(I->inv, r.d);
    return l;
}
// the ray in the object's frame: the world context itself for un-instanced objects
#define LUMO_LOCAL_CTX(S, o, wctx, c)                                                      \
    RayCtx lctx_;                                                                          \
    const RayCtx* lp_ = &(wctx);                                                           \
    if ((o).inst >= 0) { make_ctx(to_local<CNT>(S, o, (wctx).r, c), lctx_); lp_ = &lctx_; } \
    const RayCtx& l = *lp_

// util::quadratic (object.rs:57-72)
__device__ __forceinline__ bool quadratic(double a, double b, double c, double& t0, double& t1) {
    double disc = b * b - 4.0 * a * c;
    if (disc < 0.0) return false;
    double dr = sqrt(disc);
    t0 = (-b - dr) / (2.0 * a); t1 = (-b + dr) / (2.0 * a);
    if (t0 > t1) { double s = t0; t0 = t1; t1 = s; }
    return true;
}
// Sphere::hit_t (sphere.rs:77-96)
__device__ __noinline__ double sphere_hit_t(double radius, const Ray& r, double t_min, double t_max) {
    double a = dot(r.d, r.d), b = 2.0 * dot(r.d, r.o), c = dot(r.o, r.o) - radius * radius;
    double t0, t1;
    if (!quadratic(a, b, c, t0, t1)) return LUMO_INF;
    if (t0 >= t_max || t1 <= t_min) return LUMO_INF;
    if (t0 > t_min) return t0;
    if (t1 >= t_max) return LUMO_INF;
    return t1;
}

// efloat.rs:11-38
__device__ __forceinline__ double next_float(double v) {
    if (isinf(v) && v > 0.0) return v;
    if (v == -0.0) v = 0.0;
    long long bits = __double_as_longlong(v);
    bits = (v >= 0.0) ? bits + 1 : bits - 1;
    return __longlong_as_double(bits);
}
__device__ __forceinline__ double previous_float(double v) {
    if (isinf(v) && v < 0.0) return v;
    if (v == 0.0) v = -0.0;
    long long bits = __double_as_longlong(v);
    bits = (v > 0.0) ? bits - 1 : bits + 1;
    return __longlong_as_double(bits);
}
// EFloat interval arithmetic (efloat.rs:41-204), only what Sphere::hit needs
struct EF { double v, lo, hi; };
__device__ __forceinline__ EF ef(double v) { EF r; r.v = v; r.lo = v; r.hi = v; return r; }
__device__ __forceinline__ EF ef3(double v, double lo, double hi) { EF r; r.v = v; r.lo = lo; r.hi = hi; return r; }
__device__ __forceinline__ EF ef_neg(EF a) { return ef3(-a.v, -a.lo, -a.hi); }
__device__ __forceinline__ EF ef_add(EF a, EF b) { return ef3(a.v + b.v, previous_float(a.lo + b.lo), next_float(a.hi + b.hi)); }
__device__ __forceinline__ EF ef_sub(EF a, EF b) { return ef3(a.v - b.v, previous_float(a.lo - b.hi), next_float(a.hi - b.lo)); }
__device__ __forceinline__ EF ef_mul(EF a, EF b) {
    double p0 = a.lo * b.lo, p1 = a.lo * b.hi, p2 = a.hi * b.lo, p3 = a.hi * b.hi;
    double mn = fmin(fmin(fmin(p0, p1), p2), p3), mx = fmax(fmax(fmax(p0, p1), p2), p3);
    return ef3(a.v * b.v, previous_float(mn), next_float(mx));
}
__device__ __forceinline__ EF ef_div(EF a, EF b) {
    if (b.lo < 0.0 && b.hi > 0.0) return ef3(a.v / b.v, -LUMO_INF, LUMO_INF);
    double p0 = a.lo / b.lo, p1 = a.lo / b.hi, p2 = a.hi / b.lo, p3 = a.hi / b.hi;
    double mn = fmin(fmin(fmin(p0, p1), p2), p3), mx = fmax(fmax(fmax(p0, p1), p2), p3);
    return ef3(a.v / b.v, previous_float(mn), next_float(mx));
}
__device__ __forceinline__ EF ef_sqrt(EF a) { return ef3(sqrt(a.v), previous_float(sqrt(a.lo)), next_float(sqrt(a.hi))); }
// Sphere::hit distance selection (sphere.rs:28-60); returns INF when there is no hit
__device__ __noinline__ double sphere_hit(double radius, const Ray& r, double t_min, double t_max) {
    EF dx = ef(r.d.x), dy = ef(r.d.y), dz = ef(r.d.z), ox = ef(r.o.x), oy = ef(r.o.y), oz = ef(r.o.z);
    EF radius2 = ef_mul(ef(radius), ef(radius));
    EF a = ef_add(ef_add(ef_mul(dx, dx), ef_mul(dy, dy)), ef_mul(dz, dz));
    EF b = ef_mul(ef(2.0), ef_add(ef_add(ef_mul(dx, ox), ef_mul(dy, oy)), ef_mul(dz, oz)));
    EF c = ef_sub(ef_add(ef_add(ef_mul(ox, ox), ef_mul(oy, oy)), ef_mul(oz, oz)), radius2);
    double disc = b.v * b.v - 4.0 * a.v * c.v;                                            // efloat.rs:67-83
    if (disc < 0.0) return LUMO_INF;
    EF dr = ef_sqrt(ef(disc));
    EF t0 = ef_div(ef_sub(ef_neg(b), dr), ef_mul(ef(2.0), a));
    EF t1 = ef_div(ef_add(ef_neg(b), dr), ef_mul(ef(2.0), a));
    if (t0.v > t1.v) { EF s = t0; t0 = t1; t1 = s; }
    if (t0.hi >= t_max || t1.lo <= t_min) return LUMO_INF;
    if (t0.lo > t_min) return t0.v;
    if (t1.hi >= t_max) return LUMO_INF;
    return t1.v;
}

// Object::hit_t for one object record (what the BVH leaf loop calls, bvh.rs:346)
template <bool CNT, int ROUND = LUMO_KD_ROUND>
__device__ __forceinline__ double object_hit_t(const DevScene& S, const LumoObject& o, const RayCtx& w, double t_min, double t_max, Counters* c) {
    LUMO_LOCAL_CTX(S, o, w, c);
    switch (o.kind) {
    case LOBJ_KD: case LOBJ_RECT: {
        double t; uint32_t tri; D3 bary;
        return kd_hit<false, CNT, 64, ROUND>(S, S.kd_trees + o.geom, l, t_min, t_max, t, tri, bary, c) ? t : LUMO_INF;
    }
    case LOBJ_SPHERE: LUMO_CNT(sphere); return sphere_hit_t(S.spheres[o.geom].radius, l.r, t_min, t_max);
    default: {
        TriHit th;
        return tri_hit<false, CNT>(S.tri_verts + o.geom, l.r, l.q, t_min, t_max, th, c) ? th.t : LUMO_INF;
    }
    }
}

struct HitRec { double t; D3 bary; uint32_t obj, tri; };

// Object::hit for one object record: distance + which triangle + barycentrics (the rest of `Hit`
// is rebuilt from these by the shading kernels, shade.cuh reconstruct_hit).
// any_t (optional): Object::hit_t of the same object and arguments, i.e. the distance the selection pass of BVH::_hit compares
// (bvh.rs:346) — for a kd-tree the first triangle the walk accepts, for a sphere Sphere::hit_t, for a triangle its hit_t.
template <bool CNT, int STACK = 64, int ROUND = LUMO_KD_ROUND>
__device__ __noinline__ bool object_hit(const DevScene& S, const LumoObject& o, const RayCtx& w, double t_min, double t_max, HitRec& h, Counters* c, double* any_t = nullptr) {
    LUMO_LOCAL_CTX(S, o, w, c);
    if (any_t) *any_t = LUMO_INF;
    switch (o.kind) {
    case LOBJ_KD: case LOBJ_RECT:
        return kd_hit<true, CNT, STACK, ROUND>(S, S.kd_trees + o.geom, l, t_min, t_max, h.t, h.tri, h.bary, c, any_t);
    case LOBJ_SPHERE: {
        LUMO_CNT(sphere);
        if (any_t) *any_t = sphere_hit_t(S.spheres[o.geom].radius, l.r, t_min, t_max);
        double t = sphere_hit(S.spheres[o.geom].radius, l.r, t_min, t_max);
        if (!(t < LUMO_INF)) return false;
        h.t = t; h.tri = 0; h.bary = d3(0, 0, 0);
        return true;
    }
    default: {
        TriHit th;
        if (any_t) { TriHit ta; *any_t = tri_hit<false, false>(S.tri_verts + o.geom, l.r, l.q, t_min, t_max, ta, nullptr) ? ta.t : LUMO_INF; }
        if (!tri_hit<true, CNT>(S.tri_verts + o.geom, l.r, l.q, t_min, t_max, th, c)) return false;
        h.t = th.t; h.tri = 0; h.bary = th.bary;
        return true;
    }
    }
}

// BVH::_hit<GEO> (bvh.rs:315-362) over one of the two object BVHs.  obj_base = first object
// record of this BVH (0 for Scene.objects, n_objects for Scene.lights).
template <bool GEO, bool CNT, int ROUND = LUMO_KD_ROUND>
__device__ __noinline__ uint32_t tlas_hit(const DevScene& S, uint32_t root, uint32_t obj_base, const RayCtx& w, double t_min, double t_max, Counters* c, double* t_found = nullptr) {
    const Ray r = w.r;
    const D3 inv = w.inv;
    uint32_t stack[64];
    int sp = 0;
    uint32_t curr = 0, idx = LUMO_NONE;
    double tt = t_max;
    uint32_t leaf_pos = 0, leaf_end = 0;     // pending objects of the current leaf
    bool need_pop = false;
    // "while-while" form of bvh.rs:326-360: every lane walks nodes until it holds the next object to test, then
    // the lanes of the warp run Object::hit_t together.  Per lane the order of nodes and objects is the reference's.
    for (;;) {
        uint32_t obj = LUMO_NONE;
        for (;;) {
            if (leaf_pos < leaf_end) { obj = S.tlas_leaf[leaf_pos]; leaf_pos++; break; }
            if (need_pop) {
                if (sp == 0) break;
                curr = stack[--sp];
                need_pop = false;
            }
            const LumoTlasNode* node = S.tlas + root + curr;
            LUMO_CNT(tlas);
            double t_start, t_end;
            box_intersect(node->lo, node->hi, r.o, inv, t_start, t_end);
            t_start = fmax(t_start, t_min); t_end = fmin(t_end, tt);
            if (t_start <= t_end) {
                const uint32_t count = node->count;
                if (count == 0) {
                    const uint32_t right = node->right;
                    curr += 1;
                    if (right != LUMO_NONE && sp < 64) stack[sp++] = right;
                    continue;
                }
                leaf_pos = node->first; leaf_end = leaf_pos + count;
            }
            need_pop = true;
        }
        if (obj == LUMO_NONE) break;
        const double t = object_hit_t<CNT, ROUND>(S, S.objects[obj_base + obj], w, t_min, tt, c);
        if (GEO) { if (t < tt) { tt = t; idx = obj; } }
        else { if (t < tt) { if (t_found) *t_found = t; return obj; } }
    }
    return idx;
}

// Object for BVH: hit / hit_t (bvh.rs:365-375)
template <bool CNT, int ROUND = LUMO_KD_ROUND>
__device__ __forceinline__ bool bvh_hit(const DevScene& S, uint32_t root, uint32_t obj_base, const RayCtx& r, double t_min, double t_max, HitRec& h, Counters* c) {
    const uint32_t idx = tlas_hit<true, CNT, ROUND>(S, root, obj_base, r, t_min, t_max, c);
    if (idx == LUMO_NONE) return false;
    if (!object_hit<CNT, 64, ROUND>(S, S.objects[obj_base + idx], r, t_min, t_max, h, c)) return false;
    h.obj = obj_base + idx;
    return true;
}
// BVH::hit_t (bvh.rs:371-374) finds the first object with a hit and then calls hit_t on it a second time with the
// same arguments; that second call is a pure function of them, so its value is the one the traversal just
// computed.  The counting instantiation still performs it, to report the reference's visit counts.
template <bool CNT, int ROUND = LUMO_KD_ROUND>
__device__ __forceinline__ double bvh_hit_t(const DevScene& S, uint32_t root, uint32_t obj_base, const RayCtx& r, double t_min, double t_max, Counters* c) {
    double t = LUMO_INF;
    const uint32_t idx = tlas_hit<false, CNT, ROUND>(S, root, obj_base, r, t_min, t_max, c, &t);
    if (idx == LUMO_NONE) return LUMO_INF;
    if (CNT) return object_hit_t<CNT, ROUND>(S, S.objects[obj_base + idx], r, t_min, t_max, c);
    return t;
}

// Scene::hit (scene.rs:119-147): objects, then lights with t_max = h.t.  The reference always
// starts from t_max = +inf; the C ABI lets the caller pass a finite one.
template <bool CNT, int ROUND = LUMO_KD_ROUND>
__device__ __forceinline__ bool scene_hit(const DevScene& S, const Ray& ray, double t_max, HitRec& h, Counters* c) {
    RayCtx r; make_ctx(ray, r);
    bool have = bvh_hit<CNT, ROUND>(S, 0, 0, r, 0.0, t_max, h, c);
    if (have) t_max = h.t;
    if (S.P.n_lights) {
        HitRec hl;
        if (bvh_hit<CNT, ROUND>(S, S.P.lights_root, S.P.n_objects, r, 0.0, t_max, hl, c)) { h = hl; have = true; }
    }
    return have;
}
// Scene::hit_t (scene.rs:150-162)
template <bool CNT>
__device__ __forceinline__ double scene_hit_t(const DevScene& S, const Ray& ray, Counters* c) {
    RayCtx r; make_ctx(ray, r);
    double t = LUMO_INF;
    t = fmin(t, bvh_hit_t<CNT>(S, 0, 0, r, 0.0, t, c));
    if (S.P.n_lights) t = fmin(t, bvh_hit_t<CNT>(S, S.P.lights_root, S.P.n_objects, r, 0.0, t, c));
    return t;
}
// the two occlusion tests of Scene::hit_light (scene.rs:180-186)
template <bool CNT, int ROUND = LUMO_KD_ROUND>
__device__ __forceinline__ bool scene_occluded(const DevScene& S, const Ray& ray, double t_max, Counters* c) {
    RayCtx r; make_ctx(ray, r);
    if (bvh_hit_t<CNT, ROUND>(S, 0, 0, r, 0.0, t_max, c) < t_max) return true;
    if (S.P.n_lights && bvh_hit_t<CNT, ROUND>(S, S.P.lights_root, S.P.n_objects, r, 0.0, t_max, c) < t_max) return true;
    return false;
}

}  // namespace lumo_dev
