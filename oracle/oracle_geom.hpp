// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).  PARITY UNPINNED.
// CPU restatement of lumo's geometry layer: Ray, Hit, AABB, Triangle (Woop watertight),
// kd-tree (Wald-Havran SAH build + traversal), object BVH (Garanzha-style build + traversal),
// Instance, Rectangle, Sphere.  Citations are to /root/reference/src/... file:line.
#pragma once
#include "oracle_math.hpp"
#include <vector>
#include <memory>
#include <deque>
#include <cassert>
#include <cstdio>

namespace oracle {

struct Material;  // oracle_shading.hpp

// Visit counters (what SURVEY §8d's byte formula multiplies); thread-local so the multithreaded
// CPU baseline does not contend.
struct Counters {
    uint64_t tlas_nodes = 0, inst = 0, kd_nodes = 0, leaf_idx = 0, tri_tests = 0, sphere_tests = 0;
    uint64_t closest = 0, occlusion = 0;
    void add(const Counters& o) {
        tlas_nodes += o.tlas_nodes; inst += o.inst; kd_nodes += o.kd_nodes; leaf_idx += o.leaf_idx;
        tri_tests += o.tri_tests; sphere_tests += o.sphere_tests; closest += o.closest; occlusion += o.occlusion;
    }
};
extern thread_local Counters g_cnt;

// src/tracer/ray.rs
struct Ray {
    Vec3 origin, dir;
    Ray() {}
    static Ray make(Vec3 o, Vec3 d) { Ray r; r.origin = o; r.dir = d.normalize(); return r; }   // ray.rs:14-19
    static Ray raw(Vec3 o, Vec3 d) { Ray r; r.origin = o; r.dir = d; return r; }
    template <bool NORMALIZE> Ray transform(const Transform& t) const {                          // ray.rs:24-30
        Vec3 o = t.transform_pt_inv(origin);
        Vec3 d = t.transform_dir_inv(dir);
        if (NORMALIZE) d = d.normalize();
        return raw(o, d);
    }
    Vec3 at(Float t) const { return origin + t * dir; }
};

// src/tracer/hit.rs
struct Hit {
    Float t = 0;
    const Material* material = nullptr;
    Vec3 p, fp_error, ns, ng;
    Vec2 uv;
    bool backface = false;
    // parity bookkeeping (not in the reference): which primitive produced the hit
    int32_t obj = -1;     // index in Scene.objects (lights offset by objects.len())
    int32_t tri = 0;      // index in the object's KdTree.objects, 0 for analytic / loose triangle
    Vec3 bary;            // edges / det for triangles
};
static inline Vec2 wrap_uv(Vec2 uv) {                                                            // hit.rs:62-68
    Vec2 f(fract(uv.x), fract(uv.y));
    return Vec2(f.x < 0.0 ? f.x + 1.0 : f.x, f.y < 0.0 ? f.y + 1.0 : f.y);
}
static inline Hit hit_new(Float t, const Material* m, Vec3 wo, Vec3 xi, Vec3 fp_error, Vec3 ns, Vec3 ng, Vec2 uv) {  // hit.rs:37-59
    Hit h;
    h.t = t; h.material = m; h.backface = wo.dot(ng) > 0.0; h.p = xi; h.fp_error = fp_error;
    h.ns = ns; h.ng = ng; h.uv = wrap_uv(uv);
    return h;
}
static inline Vec3 hit_ray_origin(const Hit& h, bool outside) {                                  // hit.rs:85-111
    Vec3 ne = h.ng;
    Float scaled_err = h.fp_error.dot(ne.abs());
    Vec3 offset = outside ? ne * scaled_err : (-ne) * scaled_err;
    Vec3 xi = h.p + offset;
    auto move_double = [](Float v, Float n) { return n > 0.0 ? next_float(v) : (n < 0.0 ? previous_float(v) : v); };
    return Vec3(move_double(xi.x, offset.x), move_double(xi.y, offset.y), move_double(xi.z, offset.z));
}
static inline Ray hit_generate_ray(const Hit& h, Vec3 wi) {                                      // hit.rs:115-122
    Vec3 xi = hit_ray_origin(h, wi.dot(h.ng) >= 0.0);
    return Ray::make(xi, wi);
}

// src/tracer/object/aabb.rs
struct AABB {
    Vec3 ax_min, ax_max;
    AABB() : ax_min(Vec3::splat(INF)), ax_max(Vec3::splat(-INF)) {}
    AABB(Vec3 a, Vec3 b) : ax_min(a), ax_max(b) {}
    void intersect(Vec3 origin, Vec3 inv_dir, Float& t_start, Float& t_end) const {              // aabb.rs:33-44
        Vec3 ro_min = (ax_min - origin) * inv_dir;
        Vec3 ro_max = (ax_max - origin) * inv_dir;
        Vec3 ts = ro_min.min(ro_max);
        Vec3 te = ro_max.max(ro_min);
        t_start = ts.max_element();
        t_end = te.min_element() * (1.0 + 2.0 * gamma_(3));
    }
    Vec3 center() const { return ax_min + (ax_max - ax_min) / 2.0; }
    AABB merge(const AABB& o) const { return AABB(ax_min.min(o.ax_min), ax_max.max(o.ax_max)); }
    Float area() const {
        Vec3 d = ax_max - ax_min;
        return 2.0 * (d.x * d.y + d.x * d.z + d.y * d.z);
    }
    bool cuts(int axis, Float point) const { return ax_min.axis(axis) < point && point < ax_max.axis(axis); }
    Vec3 extent() const { return ax_max - ax_min; }
    void split(int axis, Float value, AABB& l, AABB& r) const {                                  // aabb.rs:98-120
        Vec3 mid_max = ax_max, mid_min = ax_min;
        if (axis == 0) { mid_max.x = value; mid_min.x = value; }
        else if (axis == 1) { mid_max.y = value; mid_min.y = value; }
        else { mid_max.z = value; mid_min.z = value; }
        l = AABB(ax_min, mid_max); r = AABB(mid_min, ax_max);
    }
};

// src/tracer/object.rs:78-157 — Object and Sampleable merged into one base class
struct Object {
    virtual ~Object() {}
    virtual bool hit(const Ray& r, Float t_min, Float t_max, Hit& out) const = 0;
    virtual Float hit_t(const Ray& r, Float t_min, Float t_max) const {                          // object.rs:85-88
        Hit h;
        Float t = hit(r, 0.0, INF, h) ? h.t : INF;
        return (t <= t_min || t >= t_max) ? INF : t;
    }
    virtual AABB bounding_box() const = 0;
    virtual size_t num_primitives() const { return 1; }
    // ---- Sampleable ----
    virtual Float area() const { assert(false); return 0; }
    virtual const Material* material() const { assert(false); return nullptr; }
    virtual Hit sample_on(Vec2 rand_sq) const { assert(false); return Hit(); }
    virtual Vec3 sample_towards(Vec3 xo, Vec2 rand_sq) const {                                   // object.rs:136-139
        Vec3 xi = sample_on(rand_sq).p;
        return (xi - xo).normalize();
    }
    virtual Float sample_towards_pdf(const Ray& ri, Vec3 xi, Vec3 ng) const {                    // object.rs:148-156
        Float p_area = 1.0 / area();
        return p_area * ri.origin.distance_squared(xi) / std::fabs(ng.dot(ri.dir));
    }
    void sample_leaving(Vec2 r0, Vec2 r1, Ray& ri, Hit& ho) const {                              // object.rs:107-117
        ho = sample_on(r0);
        Onb uvw(ho.ns);
        Vec3 wi_local = square_to_cos_hemisphere(r1);
        Vec3 wi = uvw.to_world(wi_local);
        ri = hit_generate_ray(ho, wi);
    }
    void sample_leaving_pdf(const Ray& r, Vec3 ng, Float& pdf_origin, Float& pdf_dir) const {    // object.rs:120-127
        pdf_origin = 1.0 / area();
        pdf_dir = ng.dot(r.dir) / PI;
    }
};

// src/tracer/object/triangle_mesh.rs
struct TriangleMesh {
    std::vector<Vec3> vertices, normals;
    std::vector<Vec2> uvs;
};

// src/tracer/object/triangle.rs
struct Triangle : Object {
    const TriangleMesh* mesh;
    uint32_t v[3];
    const Material* mat;
    bool has_n = false, has_t = false;
    uint32_t n[3] = {0, 0, 0}, tx[3] = {0, 0, 0};
    Vec3 a() const { return mesh->vertices[v[0]]; }
    Vec3 b() const { return mesh->vertices[v[1]]; }
    Vec3 c() const { return mesh->vertices[v[2]]; }
    Vec3 shading_normal(Vec3 bary, Vec3 ng) const {                                              // triangle.rs:49-60
        if (!has_n) return ng;
        Vec3 na = mesh->normals[n[0]], nb = mesh->normals[n[1]], nc = mesh->normals[n[2]];
        return (bary.x * na + bary.y * nb + bary.z * nc).normalize();
    }
    template <bool GEO> bool _hit(const Ray& r, Float t_min, Float t_max, Hit& out) const {      // triangle.rs:63-187
        g_cnt.tri_tests++;
        Vec3 xo = r.origin;
        Vec3 wi_abs = r.dir.abs();
        int kz = (wi_abs.x > wi_abs.y && wi_abs.x > wi_abs.z) ? 0 : (wi_abs.y > wi_abs.z ? 1 : 2);
        auto permute = [kz](Vec3 q) { return kz == 0 ? Vec3(q.y, q.z, q.x) : (kz == 1 ? Vec3(q.z, q.x, q.y) : q); };
        Vec3 wi = permute(r.dir);
        Vec3 at = permute(a() - xo), bt = permute(b() - xo), ct = permute(c() - xo);
        Vec3 shear = Vec3(-wi.x, -wi.y, 0.0) / wi.z;
        at = at + shear * at.z; bt = bt + shear * bt.z; ct = ct + shear * ct.z;
        Vec3 edges(bt.x * ct.y - bt.y * ct.x, ct.x * at.y - ct.y * at.x, at.x * bt.y - at.y * bt.x);
        if (edges.min_element() < 0.0 && edges.max_element() > 0.0) return false;
        Float det = edges.dot(Vec3(1, 1, 1));
        if (det == 0.0) return false;
        Float t_scaled = edges.dot(Vec3(at.z, bt.z, ct.z)) / wi.z;
        bool b1 = det < 0.0 && (t_scaled > t_min * det || t_scaled < t_max * det);
        bool b2 = det > 0.0 && (t_scaled < t_min * det || t_scaled > t_max * det);
        if (b1 || b2) return false;
        Float t = t_scaled / det;
        if (!GEO) { out.t = t; return true; }
        Float max_z_v = fmax_(fmax_(std::fabs(at.z), std::fabs(bt.z)), std::fabs(ct.z));
        Float delta_z = gamma_(3) * max_z_v;
        Float max_y_v = fmax_(fmax_(std::fabs(at.y), std::fabs(bt.y)), std::fabs(ct.y));
        Float delta_y = gamma_(5) * (max_y_v + max_z_v);
        Float max_x_v = fmax_(fmax_(std::fabs(at.x), std::fabs(bt.x)), std::fabs(ct.x));
        Float delta_x = gamma_(5) * (max_x_v + max_z_v);
        Float delta_e = 2.0 * (gamma_(2) * max_x_v * max_y_v + delta_y * max_x_v + delta_x * max_y_v);
        Float max_e = fmax_(fmax_(std::fabs(edges.x), std::fabs(edges.y)), std::fabs(edges.z));
        Float delta_t = 3.0 * (gamma_(3) * max_e * max_z_v + delta_e * max_z_v + delta_z * max_e) / std::fabs(det);
        if (t <= t_min + delta_t) return false;
        Vec3 bary = edges / det;
        Float alpha = bary.x, beta = bary.y, gam = bary.z;
        Vec3 ng = (b() - a()).cross(c() - a()).normalize();
        Vec3 ns = shading_normal(bary, ng);
        Vec3 xi = alpha * a() + beta * b() + gam * c();
        Vec2 ta(0, 0), tb(1, 0), tc(1, 1);
        if (has_t) { ta = mesh->uvs[tx[0]]; tb = mesh->uvs[tx[1]]; tc = mesh->uvs[tx[2]]; }
        Vec2 uv = alpha * ta + beta * tb + gam * tc;
        Vec3 err = gamma_(7) * Vec3((bary * Vec3(a().x, b().x, c().x)).abs().dot(Vec3(1, 1, 1)),
                                    (bary * Vec3(a().y, b().y, c().y)).abs().dot(Vec3(1, 1, 1)),
                                    (bary * Vec3(a().z, b().z, c().z)).abs().dot(Vec3(1, 1, 1)));
        out = hit_new(t, mat, r.dir, xi, err, ns, ng, uv);
        out.bary = bary; out.tri = 0;
        return true;
    }
    bool hit(const Ray& r, Float t_min, Float t_max, Hit& out) const override { return _hit<true>(r, t_min, t_max, out); }
    Float hit_t(const Ray& r, Float t_min, Float t_max) const override {
        Hit h; return _hit<false>(r, t_min, t_max, h) ? h.t : INF;
    }
    AABB bounding_box() const override { return AABB(a().min(b().min(c())), a().max(b().max(c()))); }
    Float area() const override { return (b() - a()).cross(c() - a()).length() / 2.0; }
    const Material* material() const override { return mat; }
    Hit sample_on(Vec2 rs) const override {                                                      // triangle.rs:215-241
        Float gam = 1.0 - std::sqrt(1.0 - rs.x);
        Float beta = rs.y * (1.0 - gam);
        Float alpha = 1.0 - gam - beta;
        Vec3 bary(alpha, beta, gam);
        Vec3 b_m_a = b() - a(), c_m_a = c() - a();
        Vec3 ng = b_m_a.cross(c_m_a).normalize();
        Vec3 ns = shading_normal(bary, ng);
        Vec3 xo = a() + beta * b_m_a + gam * c_m_a;
        Vec3 xo_abs = a().abs() + (beta * b_m_a).abs() + (gam * c_m_a).abs();
        Vec3 err = gamma_(6) * xo_abs;
        return hit_new(0.0, mat, -ng, xo, err, ns, ng, Vec2(0, 0));
    }
};

static inline bool degenerate_triangle(Vec3 a, Vec3 b, Vec3 c) { return (b - a).cross(c - a).length() == 0.0; }  // triangle_mesh.rs:93-96

// ---- kd-tree (src/tracer/object/kdtree.rs, kdtree/node.rs, kdtree/event.rs) ---------------------
struct KdNode { int axis; Float point; std::vector<uint32_t> indices; size_t right; bool leaf; };
static const size_t IDX_NAN = (size_t)-1;
enum { EV_END = 0, EV_PLANAR = 1, EV_START = 2 };
struct KdEvent { Float p; int a; int t; uint32_t idx; };

struct KdTree : Object {
    std::vector<Triangle> objects;
    std::vector<KdNode> nodes;
    AABB boundary;
    std::vector<int8_t> side_;   // scratch for partition (stands in for the FxHashMap, node.rs:198-230)

    static constexpr Float COST_TRAVERSE = 15.0, COST_INTERSECT = 20.0, EMPTY_BONUS = 0.2;       // node.rs:7-9

    explicit KdTree(std::vector<Triangle>&& tris) : objects(std::move(tris)) { build(); }

    void build() {                                                                               // kdtree.rs:43-89
        std::vector<AABB> bounds(objects.size());
        for (size_t i = 0; i < objects.size(); i++) bounds[i] = objects[i].bounding_box();
        boundary = AABB();
        for (auto& b : bounds) boundary = boundary.merge(b);
        std::vector<KdEvent> events; events.reserve(6 * objects.size());
        for (size_t i = 0; i < objects.size(); i++) {
            for (int ax = 0; ax < 3; ax++) {
                Float mi = bounds[i].ax_min.axis(ax), mx = bounds[i].ax_max.axis(ax);
                if (mi == mx) events.push_back({mi, ax, EV_PLANAR, (uint32_t)i});
                else { events.push_back({mi, ax, EV_START, (uint32_t)i}); events.push_back({mx, ax, EV_END, (uint32_t)i}); }
            }
        }
        std::stable_sort(events.begin(), events.end(), [](const KdEvent& x, const KdEvent& y) {     // event.rs:25-46
            if (x.p < y.p) return true; if (x.p > y.p) return false;
            if (x.a < y.a) return true; if (x.a > y.a) return false;
            return x.t < y.t;
        });
        side_.assign(objects.size(), 0);
        construct(std::move(events), objects.size(), boundary, IDX_NAN);
        side_.clear(); side_.shrink_to_fit();
    }

    static Float cost(const AABB& boundary, int axis, Float point, size_t nl, size_t np, size_t nr, int& side) {  // node.rs:87-122
        if (!boundary.cuts(axis, point)) { side = 0; return INF; }
        AABB left, right; boundary.split(axis, point, left, right);
        Float area_left = left.area() / boundary.area();
        Float area_right = right.area() / boundary.area();
        auto cut = [&](size_t num_left, size_t num_right) {
            Float c = COST_TRAVERSE + COST_INTERSECT * ((Float)num_left * area_left + (Float)num_right * area_right);
            return (num_left == 0 || num_right == 0) ? (1.0 - EMPTY_BONUS) * c : c;
        };
        Float cl = cut(nl + np, nr), cr = cut(nl, np + nr);
        if (cl < cr) { side = -1; return cl; }
        side = 1; return cr;
    }

    static void find_best_split(const std::vector<KdEvent>& ev, const AABB& boundary, size_t prims,
                                int& best_axis, Float& best_point, Float& best_cost, int& best_side) {  // node.rs:125-194
        best_cost = INF; best_point = INF; best_axis = 0; best_side = 0;
        size_t num_left[3] = {0, 0, 0}, num_planar[3] = {0, 0, 0}, num_right[3] = {prims, prims, prims};
        size_t i = 0;
        while (i < ev.size()) {
            size_t s = 0, p = 0, e = 0;
            const KdEvent event = ev[i];
            while (i < ev.size() && ev[i].a == event.a && ev[i].p == event.p && ev[i].t == EV_END) { e++; i++; }
            while (i < ev.size() && ev[i].a == event.a && ev[i].p == event.p && ev[i].t == EV_PLANAR) { p++; i++; }
            while (i < ev.size() && ev[i].a == event.a && ev[i].p == event.p && ev[i].t == EV_START) { s++; i++; }
            int axis = event.a;
            num_planar[axis] = p;
            num_right[axis] -= p;
            num_right[axis] -= e;
            int cut_side;
            Float c = cost(boundary, event.a, event.p, num_left[axis], num_planar[axis], num_right[axis], cut_side);
            if (c < best_cost) { best_cost = c; best_point = event.p; best_axis = event.a; best_side = cut_side; }
            num_left[axis] += s;
            num_left[axis] += p;
            num_planar[axis] = 0;
        }
    }

    // Emits nodes in pre-order, which is what KdNodeBuilder::build (node.rs:63-84) produces: left
    // child at index+1, `right` patched when the right subtree root is pushed.
    void construct(std::vector<KdEvent>&& events, size_t prims, const AABB& bnd, size_t parent) {  // node.rs:235-337
        int axis, side; Float point, c;
        find_best_split(events, bnd, prims, axis, point, c, side);
        Float cost_leaf = COST_INTERSECT * (Float)prims;
        size_t pos = nodes.size();
        if (parent != IDX_NAN) nodes[parent].right = pos;
        if (c > cost_leaf) {
            KdNode leaf{0, INF, {}, IDX_NAN, true};
            leaf.indices.reserve(prims);
            // first-occurrence order in the event list (node.rs:247-256); side_ reused as "have" set
            for (auto& e : events) if (side_[e.idx] != 9) { leaf.indices.push_back(e.idx); side_[e.idx] = 9; }
            for (auto i : leaf.indices) side_[i] = 0;
            nodes.push_back(std::move(leaf));
            return;
        }
        // partition (node.rs:198-230): 1 = Left, 2 = Right, 0 = absent (Both)
        for (auto& ev : events) {
            if (ev.a != axis) continue;
            if (ev.t == EV_END) { if (ev.p <= point) side_[ev.idx] = 1; }
            else if (ev.t == EV_START) { if (ev.p >= point) side_[ev.idx] = 2; }
            else { if (ev.p < point) side_[ev.idx] = 1; else if (ev.p > point) side_[ev.idx] = 2; }
        }
        std::vector<KdEvent> el, er;
        el.reserve(events.size()); er.reserve(events.size());
        for (auto& ev : events) {
            int8_t s = side_[ev.idx];
            if (s == 1) el.push_back(ev);
            else if (s == 2) er.push_back(ev);
            else { el.push_back(ev); er.push_back(ev); }
        }
        for (auto& ev : events) side_[ev.idx] = 0;
        events.clear(); events.shrink_to_fit();
        auto filt = [](const KdEvent& e) { return e.a == 0 && (e.t == EV_PLANAR || e.t == EV_START); };
        size_t n_l = 0, n_r = 0;
        for (auto& e : el) n_l += filt(e);
        for (auto& e : er) n_r += filt(e);
        AABB bl, br; bnd.split(axis, point, bl, br);
        nodes.push_back(KdNode{axis, point, {}, IDX_NAN, false});
        construct(std::move(el), n_l, bl, IDX_NAN);
        construct(std::move(er), n_r, br, pos);
    }

    template <bool GEO> bool _hit(const Ray& r, Float t_min, Float t_max, Hit& out) const {       // kdtree.rs:101-169
        Float origin[3] = {r.origin.x, r.origin.y, r.origin.z};
        Float inv_dir[3] = {1.0 / r.dir.x, 1.0 / r.dir.y, 1.0 / r.dir.z};
        struct E { size_t n; Float a, b; };
        E stack[64];
        size_t sp = 0;
        Float t_hit = INF;
        size_t curr = 0;
        int64_t idx = -1;
        Float t_start, t_end;
        boundary.intersect(r.origin, 1.0 / r.dir, t_start, t_end);
        t_start = fmax_(t_start, t_min); t_end = fmin_(t_end, t_max);
        while (true) {
            if (t_hit < t_start) break;
            const KdNode& node = nodes[curr];
            g_cnt.kd_nodes++;
            if (node.leaf) {
                for (uint32_t i : node.indices) {
                    g_cnt.leaf_idx++;
                    Float t = objects[i].hit_t(r, t_min, t_end);
                    if (GEO) { if (t < t_end) { t_end = t; t_hit = t; idx = i; } }
                    else { if (t < t_end) { out.t = t; return true; } }
                }
                if (sp == 0) break;
                sp--;
                curr = stack[sp].n; t_start = stack[sp].a; t_end = stack[sp].b;
            } else {
                Float point = node.point; int axis = node.axis;
                Float t_split = (point - origin[axis]) * inv_dir[axis];
                bool left_first = origin[axis] < point || (origin[axis] == point && inv_dir[axis] <= 0.0);
                size_t first = left_first ? curr + 1 : node.right;
                size_t second = left_first ? node.right : curr + 1;
                if (t_split > t_end || t_split <= 0.0) curr = first;
                else if (t_split < t_start) curr = second;
                else {
                    curr = first;
                    assert(sp < 64);
                    stack[sp] = {second, t_split, t_end};
                    t_end = t_split;
                    sp++;
                }
            }
        }
        if (idx < 0) return false;
        if (GEO) {
            bool ok = objects[idx].hit(r, t_min, t_max, out);
            if (ok) out.tri = (int32_t)idx;
            return ok;
        }
        out.t = INF; return true;   // Hit::from_t(INF), kdtree.rs:167
    }
    bool hit(const Ray& r, Float a, Float b, Hit& out) const override { return _hit<true>(r, a, b, out); }
    Float hit_t(const Ray& r, Float a, Float b) const override { Hit h; return _hit<false>(r, a, b, h) ? h.t : INF; }
    AABB bounding_box() const override { return boundary; }
    size_t num_primitives() const override { return objects.size(); }
    const Material* material() const override { return objects[0].material(); }
};

// ---- Rectangle (src/tracer/object/rectangle.rs) ------------------------------------------------
struct Rectangle : Object {
    std::unique_ptr<TriangleMesh> tm;
    std::unique_ptr<KdTree> mesh;
    Vec3 origin, b0, b1;
    const Material* mat;
    Rectangle(Vec3 a, Vec3 b, Vec3 c, const Material* m) : mat(m) {                              // rectangle.rs:27-42
        origin = b; b0 = c - origin; b1 = a - origin;
        Vec3 d = origin + b0 + b1;
        tm.reset(new TriangleMesh());
        tm->vertices = {a, b, c, d};
        std::vector<Triangle> tris;
        const uint32_t f[4] = {0, 1, 2, 3};
        for (int i = 1; i < 3; i++) {                                                            // triangle_mesh.rs:58-91
            Triangle t; t.mesh = tm.get(); t.v[0] = f[0]; t.v[1] = f[i]; t.v[2] = f[i + 1]; t.mat = m;
            if (degenerate_triangle(t.a(), t.b(), t.c())) continue;
            tris.push_back(t);
        }
        mesh.reset(new KdTree(std::move(tris)));
    }
    bool hit(const Ray& r, Float a, Float b, Hit& out) const override {                          // rectangle.rs:73-85
        if (!mesh->hit(r, a, b, out)) return false;
        out.uv = wrap_uv(Vec2(b0.dot(out.p), b1.dot(out.p)));
        return true;
    }
    Float hit_t(const Ray& r, Float a, Float b) const override { return mesh->hit_t(r, a, b); }
    AABB bounding_box() const override {                                                         // rectangle.rs:91-101
        Vec3 a = b1 + origin, b = origin, c = b0 + origin, d = origin + b0 + b1;
        return AABB(a.min(b).min(c).min(d), a.max(b).max(c).max(d));
    }
    size_t num_primitives() const override { return 2; }
    Float area() const override { return std::fabs(b0.cross(b1).length()); }
    const Material* material() const override { return mat; }
    Hit sample_on(Vec2 rs) const override {                                                      // rectangle.rs:113-133
        Vec3 xo = origin + rs.x * b0 + rs.y * b1;
        Vec3 ng = b0.cross(b1).normalize();
        Vec3 xo_abs = origin.abs() + (rs.x * b0).abs() + (rs.y * b1).abs();
        Vec3 err = gamma_(4) * xo_abs;
        return hit_new(0.0, mat, -ng, xo, err, ng, ng, Vec2(0, 0));
    }
};

// ---- Sphere (src/tracer/object/sphere.rs) ------------------------------------------------------
static inline bool quadratic(Float a, Float b, Float c, Float& t0, Float& t1) {                  // object.rs:57-72
    Float disc = b * b - 4.0 * a * c;
    if (disc < 0.0) return false;
    Float dr = std::sqrt(disc);
    t0 = (-b - dr) / (2.0 * a); t1 = (-b + dr) / (2.0 * a);
    if (t0 > t1) std::swap(t0, t1);
    return true;
}
struct Sphere : Object {
    Float radius; const Material* mat;
    Sphere(Float r, const Material* m) : radius(r), mat(m) {}
    bool hit(const Ray& r, Float t_min, Float t_max, Hit& out) const override {                  // sphere.rs:28-75
        g_cnt.sphere_tests++;
        Vec3 xo = r.origin, wi = r.dir;
        EFloat dx(wi.x), dy(wi.y), dz(wi.z), ox(xo.x), oy(xo.y), oz(xo.z);
        EFloat radius2 = EFloat(radius) * EFloat(radius);
        EFloat a = dx * dx + dy * dy + dz * dz;
        EFloat b = EFloat(2.0) * (dx * ox + dy * oy + dz * oz);
        EFloat c = ox * ox + oy * oy + oz * oz - radius2;
        EFloat t0(0), t1(0);
        if (!efloat_quadratic(a, b, c, t0, t1)) return false;
        if (t0.high >= t_max || t1.low <= t_min) return false;
        EFloat t = t0;
        if (!(t0.low > t_min)) { if (t1.high >= t_max) return false; t = t1; }
        Vec3 xi = r.at(t.value);
        xi = xi * radius / xi.length();
        Vec3 err = gamma_(5) * xi.abs();
        Vec3 ni = xi / radius;
        Float u = (lm_atan2(-ni.z, ni.x) + PI) / (2.0 * PI);
        Float v = lm_acos(-ni.y) / PI;
        out = hit_new(t.value, mat, r.dir, xi, err, ni, ni, Vec2(u, v));
        out.tri = 0; out.bary = Vec3(0, 0, 0);   // the C ABI reports barycentrics for triangles only
        return true;
    }
    Float hit_t(const Ray& r, Float t_min, Float t_max) const override {                         // sphere.rs:77-96
        g_cnt.sphere_tests++;
        Vec3 xo = r.origin, wi = r.dir;
        Float a = wi.dot(wi), b = 2.0 * wi.dot(xo), c = xo.dot(xo) - radius * radius;
        Float t0, t1;
        if (!quadratic(a, b, c, t0, t1)) return INF;
        if (t0 >= t_max || t1 <= t_min) return INF;
        if (t0 > t_min) return t0;
        if (t1 >= t_max) return INF;
        return t1;
    }
    AABB bounding_box() const override { return AABB(-Vec3::splat(radius), Vec3::splat(radius)); }
    Float area() const override { return 4.0 * PI * radius * radius; }
    const Material* material() const override { return mat; }
    Hit sample_on(Vec2 rs) const override {                                                      // sphere.rs:111-131
        Vec3 sph = square_to_sphere(rs);
        Vec3 xo = radius * sph;
        xo = xo * radius / xo.length();
        Vec3 err = xo.abs() * gamma_(5);
        Vec3 ng = xo / radius;
        return hit_new(0.0, mat, -ng, xo, err, ng, ng, Vec2(0, 0));
    }
    Vec3 sample_towards(Vec3 xo, Vec2 rs) const override {                                       // sphere.rs:136-186
        Float d2 = xo.length_squared(), r2 = radius * radius;
        Vec3 xi;
        if (d2 < r2) xi = sample_on(rs).p;
        else {
            Onb uvw(-xo.normalize());
            Float d = std::sqrt(d2);
            Float sin2_max = r2 / d2;
            Float cos_max = std::sqrt(fmax_(1.0 - sin2_max, 0.0));
            Float cos_t = (1.0 - rs.x) + rs.x * cos_max;
            Float sin_t = std::sqrt(fmax_(1.0 - cos_t * cos_t, 0.0));
            Float phi = 2.0 * PI * rs.y;
            Float ds = d * cos_t - std::sqrt(fmax_(r2 - d2 * sin_t * sin_t, 0.0));
            Float cos_a = (d2 + r2 - ds * ds) / (2.0 * d * radius);
            Float sin_a = std::sqrt(fmax_(1.0 - cos_a * cos_a, 0.0));
            Vec3 ng_local(lm_cos(phi) * sin_a, lm_sin(phi) * sin_a, cos_a);
            Vec3 ng = uvw.to_world(-ng_local).normalize();
            xi = ng * radius;
        }
        return (xi - xo).normalize();
    }
    Float sample_towards_pdf(const Ray& ri, Vec3 xi, Vec3 ng) const override {                   // sphere.rs:190-207
        Vec3 xo = ri.origin;
        Float r2 = radius * radius, d2 = xo.length_squared();
        if (d2 < r2) {
            Float p_area = 1.0 / area();
            return p_area * xo.distance_squared(xi) / std::fabs(ng.dot(ri.dir));
        }
        Float sin2_max = r2 / d2;
        Float cos_max = std::sqrt(fmax_(1.0 - sin2_max, 0.0));
        return 1.0 / (2.0 * PI * (1.0 - cos_max));
    }
};

// ---- Instance (src/tracer/object/instance.rs) --------------------------------------------------
struct Instance : Object {
    std::shared_ptr<Object> object;
    Transform transform;
    Mat3 normal_transform;
    const Material* mat = nullptr;   // Option<Material>
    explicit Instance(std::shared_ptr<Object> o) : object(o) { normal_transform = transform.to_normal(); }
    Vec3 propagate_fp_err(Vec3 xo, Vec3 fp_error) const {                                        // instance.rs:40-50
        Vec3 e3 = fp_error.abs(), p3 = xo.abs();
        Transform ta = transform.abs();
        if (e3.x == 0.0 && e3.y == 0.0 && e3.z == 0.0) return gamma_(3) * ta.transform_pt(p3);
        return gamma_(3) * ta.transform_pt(p3) + (gamma_(3) + 1.0) * ta.transform_dir(e3);
    }
    void apply(const Transform& t) { transform = t.mul(transform); normal_transform = transform.to_normal(); }  // :257-299
    void to_origin() {                                                                           // instance.rs:55-60
        AABB bb = bounding_box();
        Vec3 mid = -(bb.ax_min + bb.ax_max) / 2.0;
        apply(Transform::translation(mid.x, mid.y, mid.z));
    }
    void set_axis(int ax, Float v) {                                                             // instance.rs:63-78
        Float mn = bounding_box().ax_min.axis(ax);
        Float d = v - mn;
        apply(Transform::translation(ax == 0 ? d : 0.0, ax == 1 ? d : 0.0, ax == 2 ? d : 0.0));
    }
    bool hit(const Ray& r, Float t_min, Float t_max, Hit& h) const override {                    // instance.rs:81-100
        g_cnt.inst++;
        Ray rl = r.transform<false>(transform);
        if (!object->hit(rl, t_min, t_max, h)) return false;
        h.ns = normal_transform.mul_vec3(h.ns).normalize();
        h.ng = normal_transform.mul_vec3(h.ng).normalize();
        h.fp_error = propagate_fp_err(h.p, h.fp_error);
        if (mat) h.material = mat;
        h.p = transform.transform_pt(h.p);
        return true;
    }
    Float hit_t(const Ray& r, Float t_min, Float t_max) const override {                         // instance.rs:102-105
        g_cnt.inst++;
        Ray rl = r.transform<false>(transform);
        return object->hit_t(rl, t_min, t_max);
    }
    AABB bounding_box() const override {                                                         // instance.rs:107-128
        Vec3 mn = transform.to_translation(), mx = transform.to_translation();
        AABB bb = object->bounding_box();
        for (int ax = 0; ax < 3; ax++) {
            Vec3 ri = transform.row(ax).truncate();
            Vec3 a0 = ri * bb.ax_min, a1 = ri * bb.ax_max;
            Float mi = a0.min(a1).dot(Vec3(1, 1, 1));
            Float ma = a0.max(a1).dot(Vec3(1, 1, 1));
            if (ax == 0) { mn.x += mi; mx.x += ma; } else if (ax == 1) { mn.y += mi; mx.y += ma; } else { mn.z += mi; mx.z += ma; }
        }
        return AABB(mn, mx);
    }
    size_t num_primitives() const override { return object->num_primitives(); }
    Float area() const override {                                                                // instance.rs:134-144
        Vec3 s = transform.to_scale();
        return s.x * s.y * object->area();
    }
    const Material* material() const override { return object->material(); }
    Hit sample_on(Vec2 rs) const override {                                                      // instance.rs:148-161
        Hit ho = object->sample_on(rs);
        ho.ng = normal_transform.mul_vec3(ho.ng).normalize();
        ho.ns = normal_transform.mul_vec3(ho.ns).normalize();
        ho.p = transform.transform_pt(ho.p);
        ho.fp_error = propagate_fp_err(ho.p, ho.fp_error);
        if (mat) ho.material = mat;
        return ho;
    }
    Vec3 sample_towards(Vec3 xo, Vec2 rs) const override {                                       // instance.rs:163-168
        Vec3 xl = transform.transform_pt_inv(xo);
        Vec3 dl = object->sample_towards(xl, rs);
        return transform.transform_dir(dl).normalize();
    }
    Float sample_towards_pdf(const Ray& ri, Vec3 xi, Vec3 ng) const override {                   // instance.rs:170-199
        Mat3 nti = normal_transform.inv().transpose();
        Vec3 ng_local = nti.mul_vec3(ng).normalize();
        Vec3 xi_local = transform.transform_pt_inv(xi);
        Ray ri_local = ri.transform<true>(transform);
        Vec3 wi = ri.dir, wi_local = ri_local.dir, xo = ri.origin, xo_local = ri_local.origin;
        Float pdf_local = object->sample_towards_pdf(ri_local, xi_local, ng_local);
        Float height = std::fabs(ng.dot(transform.transform_dir(ng_local)));
        Float volume = std::fabs(transform.to_mat3().det());
        Float jacobian = volume / height;
        Float sa_conv = xo.distance_squared(xi) * std::fabs(wi_local.dot(ng_local))
            / (xo_local.distance_squared(xi_local) * std::fabs(wi.dot(ng)));
        return pdf_local * sa_conv / jacobian;
    }
};

// ---- object BVH (src/tracer/object/bvh.rs, bvh/node.rs) ----------------------------------------
struct BVHNode { size_t right = IDX_NAN; std::vector<size_t> objects; std::vector<uint64_t> codes; AABB bounds; };

struct BVH {
    std::vector<std::shared_ptr<Object>> objects;
    std::vector<BVHNode> nodes;
    AABB boundary;
    size_t num_prims = 0;
    std::vector<std::pair<Float, size_t>> alias_table;
    std::vector<Float> alias_pdf;

    static constexpr size_t MAX_LEAF_SIZE = 4, MORTON_ORDER = 10, MORTON_BITS = 30, SAH_MAX_DEPTH = 15;  // bvh.rs:10-15
    static constexpr uint64_t MORTON_MAX = 1u << 10;
    static constexpr Float COST_INTERSECT = 15.0, COST_TRAVERSE = 20.0, EMPTY_BONUS = 0.2;               // node.rs:4-6

    void add(std::shared_ptr<Object> o) {                                                        // bvh.rs:196-200
        boundary = boundary.merge(o->bounding_box());
        num_prims += o->num_primitives();
        objects.push_back(o);
    }
    uint64_t morton_code(Vec3 center) const {                                                    // bvh.rs:208-228
        Vec3 diff = center - boundary.ax_min;
        Vec3 dim = boundary.ax_max - boundary.ax_min;
        Vec3 idx = ((Float)MORTON_MAX * diff / dim).floor();
        uint64_t x = sat_u64(idx.x), y = sat_u64(idx.y), z = sat_u64(idx.z);
        auto interleave = [](uint64_t i) {
            if (i >= MORTON_MAX) i = MORTON_MAX - 1;
            i = (i | (i << 16)) & 0b00011000000000000000011111111ull;
            i = (i | (i << 8)) & 0b00011000000001111000000001111ull;
            i = (i | (i << 4)) & 0b00011000011000011000011000011ull;
            i = (i | (i << 2)) & 0b01001001001001001001001001001ull;
            return i;
        };
        return (interleave(z) << 2) | (interleave(y) << 1) | (interleave(x) << 0);
    }

    // bvh/node.rs:46-72
    static bool morton_split(const BVHNode& n, size_t depth, BVHNode& l, BVHNode& r) {
        size_t rss = MORTON_BITS - depth;
        uint64_t first = (n.codes[0] >> rss) & 1, last = (n.codes.back() >> rss) & 1;
        size_t split;
        if (first == last) {
            if (n.codes.size() > MAX_LEAF_SIZE) split = n.codes.size() / 2; else return false;
        } else {
            // slice::partition_point (binary search; see SURVEY App. D caveat for unpartitioned input)
            size_t left = 0, right = n.codes.size(), size = n.codes.size();
            while (left < right) {
                size_t mid = left + size / 2;
                if (((n.codes[mid] >> rss) & 1) == first) left = mid + 1; else right = mid;
                size = right - left;
            }
            split = left;
        }
        l.objects.assign(n.objects.begin(), n.objects.begin() + split);
        l.codes.assign(n.codes.begin(), n.codes.begin() + split);
        r.objects.assign(n.objects.begin() + split, n.objects.end());
        r.codes.assign(n.codes.begin() + split, n.codes.end());
        return true;
    }
    static Float sah_cost(const std::vector<Float>& al, size_t nl, size_t nm, const std::vector<Float>& ar, size_t nr,
                          Float total, int& side) {                                              // node.rs:181-210
        auto get = [&](size_t num_left, size_t num_right) {
            Float a_l = al[num_left], a_r = ar[num_right];
            Float c = COST_TRAVERSE + COST_INTERSECT * ((Float)num_left * a_l + (Float)num_right * a_r) / total;
            return (num_left == 0 || num_right == 0) ? c * (1.0 - EMPTY_BONUS) : c;
        };
        Float cl = get(nl + nm, nr), cr = get(nl, nm + nr);
        if (cl < cr) { side = -1; return cl; }
        side = 1; return cr;
    }
    bool sah_split(const BVHNode& n, BVHNode& l, BVHNode& r) const {                             // node.rs:74-179
        Float best_cost = INF, best_center = INF; int best_axis = 0, best_side = 0;
        for (int axis = 0; axis < 3; axis++) {
            std::vector<size_t> idx = n.objects;
            std::stable_sort(idx.begin(), idx.end(), [&](size_t i, size_t j) {
                Float pi = objects[i]->bounding_box().center().axis(axis);
                Float pj = objects[j]->bounding_box().center().axis(axis);
                return total_cmp_less(pi, pj);
            });
            std::vector<Float> al, ar;
            al.push_back(INF);
            AABB b;
            for (size_t i : idx) { b = b.merge(objects[i]->bounding_box()); al.push_back(b.area()); }
            ar.push_back(INF);
            b = AABB();
            for (size_t k = idx.size(); k-- > 0;) { b = b.merge(objects[idx[k]]->bounding_box()); ar.push_back(b.area()); }
            Float total_area = ar[idx.size()];
            auto get_center = [&](size_t i) { return i == idx.size() ? INF : objects[idx[i]]->bounding_box().center().axis(axis); };
            size_t i = 0;
            while (i < idx.size()) {
                Float center = get_center(i);
                size_t nm = 1;
                while (nm + i <= idx.size() && center == get_center(i + nm)) nm++;
                size_t nl = i, nr = idx.size() - i - nm;
                int side;
                Float c = sah_cost(al, nl, nm, ar, nr, total_area, side);
                if (c < best_cost) { best_cost = c; best_axis = axis; best_center = center; best_side = side; }
                i += nm;
            }
        }
        // sah_partition (node.rs:145-179)
        BVHNode left, right;
        for (size_t i = 0; i < n.codes.size(); i++) {
            Float c = objects[n.objects[i]]->bounding_box().center().axis(best_axis);
            if (c < best_center || (c == best_center && best_side == -1)) { left.codes.push_back(n.codes[i]); left.objects.push_back(n.objects[i]); }
            else if (c > best_center || (c == best_center && best_side == 1)) { right.codes.push_back(n.codes[i]); right.objects.push_back(n.objects[i]); }
            else { assert(false && "unreachable (node.rs:166)"); }
        }
        if (left.codes.empty()) { l = std::move(right); r = std::move(left); }
        else { l = std::move(left); r = std::move(right); }
        return true;
    }
    static bool total_cmp_less(Float a, Float b) {   // f64::total_cmp
        int64_t x = (int64_t)to_bits(a), y = (int64_t)to_bits(b);
        x ^= (int64_t)((uint64_t)(x >> 63) >> 1);
        y ^= (int64_t)((uint64_t)(y >> 63) >> 1);
        return x < y;
    }
    bool split(const BVHNode& n, size_t depth, BVHNode& l, BVHNode& r) const {                   // node.rs:32-44
        if (n.objects.size() <= 1) return false;
        if (depth > SAH_MAX_DEPTH) return morton_split(n, depth, l, r);
        return sah_split(n, l, r);
    }
    void build() {                                                                               // bvh.rs:232-313
        assert(!objects.empty());
        std::vector<std::pair<uint64_t, size_t>> codes;
        for (size_t i = 0; i < objects.size(); i++) codes.push_back({morton_code(objects[i]->bounding_box().center()), i});
        std::sort(codes.begin(), codes.end());
        BVHNode root;
        for (auto& c : codes) { root.codes.push_back(c.first); root.objects.push_back(c.second); }
        struct Q { BVHNode node; size_t idx; bool is_left; size_t depth; };
        std::deque<Q> que;
        que.push_back(Q{std::move(root), IDX_NAN, true, 1});
        while (!que.empty()) {
            Q q = std::move(que.front()); que.pop_front();
            nodes.push_back(std::move(q.node));
            size_t pos = nodes.size() - 1;
            if (q.idx != IDX_NAN && !q.is_left) nodes[q.idx].right = pos;
            BVHNode l, r;
            if (!split(nodes[pos], q.depth, l, r)) { nodes[pos].codes.clear(); continue; }
            bool right_nonempty = !r.objects.empty();
            que.push_front(Q{std::move(l), pos, true, q.depth + 1});
            if (right_nonempty) que.push_back(Q{std::move(r), pos, false, q.depth + 1});
        }
        for (size_t i = nodes.size(); i-- > 0;) {
            bool is_leaf = nodes[i].codes.size() != nodes[i].objects.size();
            if (is_leaf) {
                AABB b;
                for (size_t o : nodes[i].objects) b = b.merge(objects[o]->bounding_box());
                nodes[i].bounds = b;
            } else {
                nodes[i].objects.clear(); nodes[i].codes.clear();
                size_t right = nodes[i].right;
                nodes[i].bounds = right == IDX_NAN ? nodes[i + 1].bounds : nodes[i + 1].bounds.merge(nodes[right].bounds);
            }
        }
    }

    template <bool GEO> int64_t _hit(const Ray& r, Float t_min, Float t_max) const {             // bvh.rs:315-362
        Vec3 origin = r.origin, inv_dir = 1.0 / r.dir;
        size_t stack[64]; size_t sp = 0, curr = 0;
        int64_t idx = -1;
        Float tt = t_max;
        while (true) {
            const BVHNode& node = nodes[curr];
            g_cnt.tlas_nodes++;
            Float t_start, t_end;
            node.bounds.intersect(origin, inv_dir, t_start, t_end);
            t_start = fmax_(t_start, t_min); t_end = fmin_(t_end, tt);
            if (t_start <= t_end) {
                if (node.objects.empty()) {
                    curr += 1;
                    if (node.right != IDX_NAN) { assert(sp < 64); stack[sp++] = node.right; }
                    continue;
                } else {
                    for (size_t i : node.objects) {
                        Float t = objects[i]->hit_t(r, t_min, tt);
                        if (GEO) { if (t < tt) { tt = t; idx = (int64_t)i; } }
                        else { if (t < tt) return (int64_t)i; }
                    }
                }
            }
            if (sp == 0) break;
            curr = stack[--sp];
        }
        return idx;
    }
    bool hit(const Ray& r, Float t_min, Float t_max, Hit& out) const {                           // bvh.rs:366-369
        int64_t idx = _hit<true>(r, t_min, t_max);
        if (idx < 0) return false;
        bool ok = objects[idx]->hit(r, t_min, t_max, out);
        if (ok) out.obj = (int32_t)idx;
        return ok;
    }
    Float hit_t(const Ray& r, Float t_min, Float t_max) const {                                  // bvh.rs:371-374
        int64_t idx = _hit<false>(r, t_min, t_max);
        return idx < 0 ? INF : objects[idx]->hit_t(r, t_min, t_max);
    }
    AABB bounding_box() const { return boundary; }
};

}  // namespace oracle
