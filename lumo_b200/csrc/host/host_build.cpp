// Host scene builder: executes a scene program (lumo_b200/program.py — the serialised
// Scene/Camera construction calls of the reference API) and produces the device scene blob
// (csrc/common/scene_blob.h).  The acceleration structures are lumo's own:
//   * per-mesh kd-tree: Wald-Havran O(n log n) SAH build, src/tracer/object/kdtree.rs:43-89 and
//     kdtree/node.rs:87-337, flattened in DFS pre-order (node.rs:63-84);
//   * object BVH: Garanzha-style Morton + SAH build, src/tracer/object/bvh.rs:208-313 and
//     bvh/node.rs:32-210;
//   * light alias table, bvh.rs:105-166.
// Written iteratively over flat arrays (no boxed node tree); results are node-for-node identical
// to the reference structures (tests/test_host_build.py compares against the oracle).
#include "host_math.h"
#include "spectra_data.h"
#include "../common/scene_blob.h"
#include "../common/lumo_math.h"
#include "ah_bvh.h"
#include "srgb_table.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace lumo_host {

static thread_local std::string g_err;

// ---- program reader -----------------------------------------------------------------------------
enum { TAG_MATERIAL = 1, TAG_MESH = 2, TAG_OBJECT = 3, TAG_ENVMAP = 4, TAG_CAMERA = 5, TAG_TEXTURE = 6 };
enum { OBJ_KDMESH = 0, OBJ_RECT = 1, OBJ_SPHERE = 2, OBJ_LOOSE_TRIS = 3 };
enum { OP_UNIT = 0, OP_ORIGIN, OP_SETX, OP_SETY, OP_SETZ, OP_TRANSLATE, OP_SCALE, OP_ROTX, OP_ROTY, OP_ROTZ };

struct Cursor {
    const uint8_t* p;
    int64_t i() { int64_t v; std::memcpy(&v, p, 8); p += 8; return v; }
    double f() { double v; std::memcpy(&v, p, 8); p += 8; return v; }
};

struct MeshIn {
    const double *v, *n, *t;
    int64_t nv, nn, nt, nf, nc;
    const int64_t *off, *vi, *ni, *ti;
    uint32_t normal_base = 0, uv_base = 0;   // offsets into the global normal / uv sections
};

// ---- spectral helpers (host copies of what the alias table needs) -------------------------------
static double dense_sample(const double* v, double lambda) {                    // dense_spectrum.rs:77-97
    const double STEP = (830.0 - 360.0) / (95.0 - 1.0);
    uint64_t b1 = sat_u64(std::ceil((lambda - 360.0) / STEP));
    double l1 = 360.0 + STEP * (double)b1;
    if (lambda == 0.0) return 0.0;
    if (lambda == l1) return v[b1];
    uint64_t b0 = b1 - 1;
    double l0 = l1 - STEP;
    double x1 = (lambda - l0) / STEP, x0 = 1.0 - x1;
    return v[b0] * x0 + v[b1] * x1;
}
static double spectrum_sample(const float* c, double lambda) {                  // spectrum.rs:108-118 (f32)
    float l = (float)lambda;
    float x = c[0] * l * l + c[1] * l + c[2];
    float sg = 0.5f + x / (2.0f * std::sqrt(1.0f + x * x));
    return (double)(c[3] * sg);
}
static double lambda_sample_one(double u) { return 538.0 - 138.888889 * lm_atanh(0.85691062 - 253.819 * u * 0.0072); }   // wavelength.rs:48-51
static double lambda_pdf_one(double l) {                                        // wavelength.rs:60-66
    if (l < 360.0 || l > 830.0) return 0.0;
    double c = lm_cosh(0.0072 * (l - 538.05));
    return 1.0 / (253.819 * (c * c));
}
static const double* illuminant(int id) {
    switch (id) { case 0: return spectra::A; case 1: return spectra::D50; case 2: return spectra::D65;
                  case 3: return spectra::F2; case 4: return spectra::F7; default: return spectra::CORNELL; }
}
static V3 dense_to_xyz(const double* v) {                                        // dense_spectrum.rs:99-117
    double s[3] = {0, 0, 0};
    const double* cmf[3] = {spectra::X, spectra::Y, spectra::Z};
    for (int k = 0; k < 3; k++) { double sum = 0.0; for (int i = 0; i < 95; i++) sum += v[i] * cmf[k][i]; s[k] = sum / 106.856895; }
    return v3(s[0], s[1], s[2]);
}
static V3 from_xyY(double x, double y, double Y) { if (y == 0.0) return v3(0, 0, 0); return v3(x * Y / y, Y, (1.0 - x - y) * Y / y); }   // xyz.rs:8-18
static void to_xy(V3 c, double& x, double& y) { x = c.x / (c.x + c.y + c.z); y = c.y / (c.x + c.y + c.z); }

// ---- geometry records being accumulated ---------------------------------------------------------
struct Builder {
    std::vector<LumoTlasNode> tlas; std::vector<uint32_t> tlas_leaf;
    std::vector<LumoObject> objects, light_objects;
    std::vector<Box> object_boxes, light_boxes;
    std::vector<LumoInstance> instances;
    std::vector<LumoKdTree> kd_trees; std::vector<LumoKdNode> kd_nodes; std::vector<uint32_t> kd_leaf;
    struct KdJob { uint32_t tri_base, n, tree; };
    std::vector<KdJob> kd_jobs;                      // trees requested while the program is read; built together by finish_kd_trees
    std::vector<LumoTriVerts> tri_verts; std::vector<LumoTriShade> tri_shade;
    std::vector<double> normals, uvs;
    std::vector<LumoRect> rects; std::vector<LumoSphere> spheres;
    std::vector<LumoMaterial> materials; std::vector<double> tables;
    std::vector<LumoTexture> textures; std::vector<float> tex_pixels; std::vector<double> tex_f64;
    std::vector<LumoLight> lights;
    std::vector<double> light_area;
    std::vector<int32_t> light_power_mat;   // Instance::material() is the inner object's (instance.rs:146)
    std::vector<MeshIn> meshes;
    std::map<std::tuple<int64_t, int64_t, int64_t>, uint32_t> kd_cache;
    LumoSceneParams params;
    bool have_env = false; float env_spec[4]; double env_scale = 0; int64_t env_tex = -1;

    uint32_t add_table(const double* src) {
        // dedupe identical tables (constant eta/k are common)
        size_t n = tables.size() / 96;
        for (size_t i = LTAB_FIRST_FREE; i < n; i++) if (std::memcmp(&tables[96 * i], src, 95 * 8) == 0) return (uint32_t)i;
        tables.insert(tables.end(), src, src + 95); tables.push_back(0.0);
        return (uint32_t)n;
    }
};

// ---- kd-tree build ------------------------------------------------------------------------------
enum { EV_END = 0, EV_PLANAR = 1, EV_START = 2 };
struct Ev { double p; uint32_t idx; uint8_t ax, type; };
static const double KD_TRAVERSE = 15.0, KD_INTERSECT = 20.0, KD_EMPTY_BONUS = 0.2;      // kdtree/node.rs:7-9

static double kd_cost(const Box& box, int ax, double point, size_t nl, size_t np, size_t nr) {   // node.rs:87-122
    if (!cuts(box, ax, point)) return kInf;
    Box l, r; split_box(box, ax, point, l, r);
    double al = area(l) / area(box), ar = area(r) / area(box);
    double c1 = KD_TRAVERSE + KD_INTERSECT * ((double)(nl + np) * al + (double)nr * ar);
    if (nl + np == 0 || nr == 0) c1 = (1.0 - KD_EMPTY_BONUS) * c1;
    double c2 = KD_TRAVERSE + KD_INTERSECT * ((double)nl * al + (double)(np + nr) * ar);
    if (nl == 0 || np + nr == 0) c2 = (1.0 - KD_EMPTY_BONUS) * c2;
    return c1 < c2 ? c1 : c2;
}

static double g_kd_seconds = 0.0;   // LUMO_HOST_TIMING
// Registers one tree over triangles [tri_base, tri_base + n) of B.tri_verts: its record (root box, triangle range) exists at
// once — the object that owns it only needs the box — while nodes and leaf lists are built by finish_kd_trees.
static uint32_t build_kd(Builder& B, uint32_t tri_base, uint32_t n) {
    Box root_box = Box::empty();
    for (uint32_t i = 0; i < n; i++) {
        const LumoTriVerts& t = B.tri_verts[tri_base + i];
        V3 a = v3(t.a[0], t.a[1], t.a[2]), b = v3(t.b[0], t.b[1], t.b[2]), c = v3(t.c[0], t.c[1], t.c[2]);
        Box tb; tb.lo = vmin(a, vmin(b, c)); tb.hi = vmax(a, vmax(b, c));                        // triangle.rs:199-204
        root_box = merge(root_box, tb);
    }
    LumoKdTree rec;
    rec.lo[0] = root_box.lo.x; rec.lo[1] = root_box.lo.y; rec.lo[2] = root_box.lo.z;
    rec.hi[0] = root_box.hi.x; rec.hi[1] = root_box.hi.y; rec.hi[2] = root_box.hi.z;
    rec.root = 0; rec.tri_base = tri_base; rec.n_tris = n; rec.pad = 0;
    B.kd_trees.push_back(rec);
    B.kd_jobs.push_back({tri_base, n, (uint32_t)B.kd_trees.size() - 1});
    return (uint32_t)B.kd_trees.size() - 1;
}
// One node of the build (node.rs:125-294): the sweep over the sorted events, then either a leaf (its triangle list) or the split
// plane and the two children's event lists.  `side` is scratch of one byte per triangle of the tree, all zero between calls.
struct KdWork { std::vector<Ev> ev; size_t prims; Box box; };
static bool kd_node(KdWork& w, std::vector<uint8_t>& side, double& point, int& axis_out, KdWork& L, KdWork& R, std::vector<uint32_t>& leaf_list) {
    // sweep (node.rs:125-194)
    double best_cost = kInf, best_point = kInf; int best_axis = 0;
    size_t nl[3] = {0, 0, 0}, nr[3] = {w.prims, w.prims, w.prims};
    const std::vector<Ev>& ev = w.ev;
    for (size_t i = 0; i < ev.size();) {
        const double p = ev[i].p; const int ax = ev[i].ax;
        size_t cnt[3] = {0, 0, 0};
        for (int ty = EV_END; ty <= EV_START; ty++)
            while (i < ev.size() && ev[i].ax == ax && ev[i].p == p && ev[i].type == ty) { cnt[ty]++; i++; }
        nr[ax] -= cnt[EV_PLANAR]; nr[ax] -= cnt[EV_END];
        double c = kd_cost(w.box, ax, p, nl[ax], cnt[EV_PLANAR], nr[ax]);
        if (c < best_cost) { best_cost = c; best_point = p; best_axis = ax; }
        nl[ax] += cnt[EV_START]; nl[ax] += cnt[EV_PLANAR];
    }
    if (best_cost > KD_INTERSECT * (double)w.prims) {                                         // node.rs:245-257
        leaf_list.clear();
        for (const Ev& e : ev) if (!side[e.idx]) { side[e.idx] = 1; leaf_list.push_back(e.idx); }
        for (uint32_t t : leaf_list) side[t] = 0;
        return false;
    }
    // classify (node.rs:198-230): 1 left only, 2 right only, 0 both; later events overwrite
    for (const Ev& e : ev) {
        if (e.ax != best_axis) continue;
        if (e.type == EV_END) { if (e.p <= best_point) side[e.idx] = 1; }
        else if (e.type == EV_START) { if (e.p >= best_point) side[e.idx] = 2; }
        else { if (e.p < best_point) side[e.idx] = 1; else if (e.p > best_point) side[e.idx] = 2; }
    }
    L.ev.clear(); R.ev.clear();
    L.ev.reserve(ev.size()); R.ev.reserve(ev.size());
    L.prims = R.prims = 0;
    for (const Ev& e : ev) {                                                                 // node.rs:269-294
        const uint8_t sd = side[e.idx];
        const bool counts = e.ax == 0 && (e.type == EV_PLANAR || e.type == EV_START);
        if (sd != 2) { L.ev.push_back(e); L.prims += counts; }
        if (sd != 1) { R.ev.push_back(e); R.prims += counts; }
    }
    for (const Ev& e : ev) side[e.idx] = 0;
    split_box(w.box, best_axis, best_point, L.box, R.box);
    point = best_point; axis_out = best_axis;
    return true;
}
// The subtree under `root` in pre-order into nodes / leafs (appended; child links and leaf-list starts are absolute in them).
static void kd_subtree(KdWork&& root, uint32_t n, std::vector<LumoKdNode>& nodes, std::vector<uint32_t>& leafs) {
    struct Item { KdWork w; uint32_t patch; };
    std::vector<Item> stack;
    stack.push_back(Item{std::move(root), LUMO_NONE});
    std::vector<uint8_t> side(n, 0);
    std::vector<uint32_t> leaf_list;
    while (!stack.empty()) {
        Item it = std::move(stack.back()); stack.pop_back();
        const uint32_t self = (uint32_t)nodes.size();
        if (it.patch != LUMO_NONE) nodes[it.patch].a = self;
        double point; int ax; KdWork L, R;
        if (!kd_node(it.w, side, point, ax, L, R, leaf_list)) {
            LumoKdNode leaf; leaf.point = kInf; leaf.a = (uint32_t)leafs.size(); leaf.b = 0x80000000u | (uint32_t)leaf_list.size();
            leafs.insert(leafs.end(), leaf_list.begin(), leaf_list.end());
            nodes.push_back(leaf);
            continue;
        }
        LumoKdNode inner; inner.point = point; inner.a = LUMO_NONE; inner.b = (uint32_t)ax;
        nodes.push_back(inner);
        it.w.ev.clear(); it.w.ev.shrink_to_fit();
        stack.push_back(Item{std::move(R), self});   // right is visited after the whole left subtree -> pre-order
        stack.push_back(Item{std::move(L), LUMO_NONE});
    }
}
// One tree into `nodes` / `leafs` (indices local to the tree: child links from node 0, leaf lists from entry 0).  A large tree
// (one mesh of hundreds of thousands of triangles) is opened from the root down to a few dozen subtrees, which are built on
// `threads` threads and put back in pre-order: the arrays are the ones the one-thread walk produces.
static void build_kd_local(const std::vector<LumoTriVerts>& tri_verts, uint32_t tri_base, uint32_t n, std::vector<LumoKdNode>& nodes, std::vector<uint32_t>& leafs, unsigned threads) {
    std::vector<Box> tb(n);
    Box root_box = Box::empty();
    for (uint32_t i = 0; i < n; i++) {
        const LumoTriVerts& t = tri_verts[tri_base + i];
        V3 a = v3(t.a[0], t.a[1], t.a[2]), b = v3(t.b[0], t.b[1], t.b[2]), c = v3(t.c[0], t.c[1], t.c[2]);
        tb[i].lo = vmin(a, vmin(b, c)); tb[i].hi = vmax(a, vmax(b, c));                       // triangle.rs:199-204
        root_box = merge(root_box, tb[i]);
    }
    std::vector<Ev> events; events.reserve(6 * (size_t)n);
    for (uint32_t i = 0; i < n; i++) for (int ax = 0; ax < 3; ax++) {                            // kdtree.rs:56-69
        double mi = axis(tb[i].lo, ax), mx = axis(tb[i].hi, ax);
        if (mi == mx) events.push_back({mi, i, (uint8_t)ax, EV_PLANAR});
        else { events.push_back({mi, i, (uint8_t)ax, EV_START}); events.push_back({mx, i, (uint8_t)ax, EV_END}); }
    }
    std::stable_sort(events.begin(), events.end(), [](const Ev& x, const Ev& y) {                 // event.rs:25-46
        if (x.p != y.p) return x.p < y.p;
        if (x.ax != y.ax) return x.ax < y.ax;
        return x.type < y.type;
    });
    KdWork root{std::move(events), n, root_box};
    nodes.clear(); leafs.clear();
    if (threads <= 1 || n < 65536u) { kd_subtree(std::move(root), n, nodes, leafs); return; }
    // top of the tree: nodes are opened level by level until there are enough open subtrees
    struct Top { bool leaf = false, open = false; LumoKdNode node; int left = -1, right = -1; std::vector<uint32_t> list; KdWork w; size_t sub = 0; };
    std::vector<Top> top(1);
    top[0].open = true; top[0].w = std::move(root);
    std::vector<uint8_t> side(n, 0);
    for (int level = 0; level < 16; level++) {
        size_t n_open = 0; for (const Top& t : top) n_open += t.open;
        if (n_open >= (size_t)threads * 4u) break;
        bool any = false;
        const size_t cur = top.size();
        for (size_t k = 0; k < cur; k++) {
            if (!top[k].open || top[k].w.ev.size() < 32768u) continue;       // small ranges stay open: they are built in the parallel phase
            double point; int ax; KdWork L, R; std::vector<uint32_t> list;
            KdWork w = std::move(top[k].w);
            top[k].open = false; any = true;
            if (!kd_node(w, side, point, ax, L, R, list)) { top[k].leaf = true; top[k].list = std::move(list); continue; }
            top[k].node.point = point; top[k].node.a = LUMO_NONE; top[k].node.b = (uint32_t)ax;
            Top l, r; l.open = r.open = true; l.w = std::move(L); r.w = std::move(R);
            top[k].left = (int)top.size(); top.push_back(std::move(l));
            top[k].right = (int)top.size(); top.push_back(std::move(r));
        }
        if (!any) break;
    }
    std::vector<size_t> open_ids;
    for (size_t k = 0; k < top.size(); k++) if (top[k].open) { top[k].sub = open_ids.size(); open_ids.push_back(k); }
    std::vector<std::vector<LumoKdNode>> sn(open_ids.size()); std::vector<std::vector<uint32_t>> sl(open_ids.size());
    {
        std::atomic<size_t> next{0};
        auto worker = [&]() { for (;;) { const size_t k = next.fetch_add(1); if (k >= open_ids.size()) break; kd_subtree(std::move(top[open_ids[k]].w), n, sn[k], sl[k]); } };
        const unsigned T = (unsigned)std::min<size_t>(threads, open_ids.size());
        if (T <= 1) worker();
        else { std::vector<std::thread> th; th.reserve(T); for (unsigned t = 0; t < T; t++) th.emplace_back(worker); for (auto& x : th) x.join(); }
    }
    // pre-order assembly
    struct Fr { int id; uint32_t patch; };
    std::vector<Fr> st; st.push_back({0, LUMO_NONE});
    while (!st.empty()) {
        const Fr f = st.back(); st.pop_back();
        const uint32_t self = (uint32_t)nodes.size();
        if (f.patch != LUMO_NONE) nodes[f.patch].a = self;
        Top& t = top[(size_t)f.id];
        if (t.open) {
            const uint32_t nb = (uint32_t)nodes.size(), lb = (uint32_t)leafs.size();
            for (LumoKdNode nd : sn[t.sub]) { nd.a += (nd.b & 0x80000000u) ? lb : nb; nodes.push_back(nd); }
            leafs.insert(leafs.end(), sl[t.sub].begin(), sl[t.sub].end());
            std::vector<LumoKdNode>().swap(sn[t.sub]); std::vector<uint32_t>().swap(sl[t.sub]);
        } else if (t.leaf) {
            LumoKdNode leaf; leaf.point = kInf; leaf.a = (uint32_t)leafs.size(); leaf.b = 0x80000000u | (uint32_t)t.list.size();
            leafs.insert(leafs.end(), t.list.begin(), t.list.end());
            nodes.push_back(leaf);
        } else {
            nodes.push_back(t.node);
            st.push_back({t.right, self});
            st.push_back({t.left, LUMO_NONE});
        }
    }
}
// Builds every registered tree — the trees are independent, so they are spread over the host's threads — and appends them in
// the order they were registered: the node and leaf arrays are exactly what building them one after the other gives.
static void finish_kd_trees(Builder& B) {
    const auto t0 = std::chrono::steady_clock::now();
    const size_t nj = B.kd_jobs.size();
    std::vector<std::vector<LumoKdNode>> nodes(nj); std::vector<std::vector<uint32_t>> leafs(nj);
    unsigned threads = std::thread::hardware_concurrency(); if (threads == 0) threads = 1; if (threads > 32) threads = 32;
    if (const char* e = std::getenv("LUMO_HOST_THREADS")) { const int v = std::atoi(e); if (v >= 1 && v <= 256) threads = (unsigned)v; }
    {
        std::atomic<size_t> next{0};
        auto worker = [&]() { for (;;) { const size_t k = next.fetch_add(1); if (k >= nj) break; build_kd_local(B.tri_verts, B.kd_jobs[k].tri_base, B.kd_jobs[k].n, nodes[k], leafs[k], threads); } };
        const unsigned T = (unsigned)std::min<size_t>(threads, nj);
        if (T <= 1) worker();
        else { std::vector<std::thread> th; th.reserve(T); for (unsigned t = 0; t < T; t++) th.emplace_back(worker); for (auto& x : th) x.join(); }
    }
    for (size_t k = 0; k < nj; k++) {
        const uint32_t node_base = (uint32_t)B.kd_nodes.size(), leaf_base = (uint32_t)B.kd_leaf.size();
        for (LumoKdNode nd : nodes[k]) {
            if (nd.b & 0x80000000u) nd.a += leaf_base;           // leaf: first entry of its list
            else nd.a += node_base;                              // inner: right child
            B.kd_nodes.push_back(nd);
        }
        B.kd_leaf.insert(B.kd_leaf.end(), leafs[k].begin(), leafs[k].end());
        B.kd_trees[B.kd_jobs[k].tree].root = node_base;
        std::vector<LumoKdNode>().swap(nodes[k]); std::vector<uint32_t>().swap(leafs[k]);
    }
    B.kd_jobs.clear();
    g_kd_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// fan triangulation + degenerate drop (triangle_mesh.rs:58-96); returns [first, count) in tri arrays
static void emit_triangles(Builder& B, const MeshIn& m, int64_t f0, int64_t f1, uint32_t& first, uint32_t& count) {
    first = (uint32_t)B.tri_verts.size();
    for (int64_t f = f0; f < f1; f++) {
        const int64_t b = m.off[f], e = m.off[f + 1];
        for (int64_t i = 1; i + 1 < e - b; i++) {
            const int64_t c[3] = {b, b + i, b + i + 1};
            LumoTriVerts tv; tv.pad = 0;
            double* dst[3] = {tv.a, tv.b, tv.c};
            for (int k = 0; k < 3; k++) for (int d = 0; d < 3; d++) dst[k][d] = m.v[3 * m.vi[c[k]] + d];
            V3 a = v3(tv.a[0], tv.a[1], tv.a[2]), bb = v3(tv.b[0], tv.b[1], tv.b[2]), cc = v3(tv.c[0], tv.c[1], tv.c[2]);
            if (length(cross(sub(bb, a), sub(cc, a))) == 0.0) continue;
            LumoTriShade ts; std::memset(&ts, 0, sizeof ts);
            // a face may come without normals / uvs even if the mesh has some (Face::nidx / tidx empty, triangle_mesh.rs:4-23): index -1
            if (m.ni && m.ni[c[0]] >= 0 && m.ni[c[1]] >= 0 && m.ni[c[2]] >= 0) { ts.flags |= 1; for (int k = 0; k < 3; k++) ts.n[k] = m.normal_base + (uint32_t)m.ni[c[k]]; }
            if (m.ti && m.ti[c[0]] >= 0 && m.ti[c[1]] >= 0 && m.ti[c[2]] >= 0) { ts.flags |= 2; for (int k = 0; k < 3; k++) ts.t[k] = m.uv_base + (uint32_t)m.ti[c[k]]; }
            B.tri_verts.push_back(tv); B.tri_shade.push_back(ts);
        }
    }
    count = (uint32_t)B.tri_verts.size() - first;
}

// ---- instance bookkeeping -----------------------------------------------------------------------
static Box xform_box(const Xform& t, const Box& bb) {                                           // instance.rs:107-128
    V3 tr = v3(t.m.a[3], t.m.a[7], t.m.a[11]);
    V3 lo = tr, hi = tr;
    double* l[3] = {&lo.x, &lo.y, &lo.z}; double* h[3] = {&hi.x, &hi.y, &hi.z};
    for (int ax = 0; ax < 3; ax++) {
        V3 ri = v3(t.m.a[4 * ax], t.m.a[4 * ax + 1], t.m.a[4 * ax + 2]);
        V3 a0 = mul(ri, bb.lo), a1 = mul(ri, bb.hi);
        *l[ax] += dot(vmin(a0, a1), v3(1, 1, 1));
        *h[ax] += dot(vmax(a0, a1), v3(1, 1, 1));
    }
    Box r = {lo, hi}; return r;
}
static uint32_t add_instance(Builder& B, const Xform& t) {
    LumoInstance I; std::memset(&I, 0, sizeof I);
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) { I.inv[4 * r + c] = t.inv.a[4 * r + c]; I.m[4 * r + c] = t.m.a[4 * r + c]; }
    M3 nrm = transpose3(to3(t.inv));                                                            // transform.rs:50-52
    std::memcpy(I.nrm, nrm.a, sizeof nrm.a);
    B.instances.push_back(I);
    return (uint32_t)B.instances.size() - 1;
}

// ---- object BVH (bvh.rs:208-313, bvh/node.rs:32-210) --------------------------------------------
static const size_t BVH_MAX_LEAF = 4, MORTON_BITS = 30, SAH_MAX_DEPTH = 15;
static const uint64_t MORTON_MAX = 1u << 10;
static const double BVH_INTERSECT = 15.0, BVH_TRAVERSE = 20.0, BVH_EMPTY_BONUS = 0.2;

static uint64_t morton(const Box& all, V3 c) {
    V3 diff = sub(c, all.lo), dim = sub(all.hi, all.lo);
    double q[3] = {std::floor((double)MORTON_MAX * diff.x / dim.x), std::floor((double)MORTON_MAX * diff.y / dim.y), std::floor((double)MORTON_MAX * diff.z / dim.z)};
    uint64_t out = 0;
    for (int k = 0; k < 3; k++) {
        uint64_t i = sat_u64(q[k]);
        if (i >= MORTON_MAX) i = MORTON_MAX - 1;
        i = (i | (i << 16)) & 0x30000FFull;
        i = (i | (i << 8)) & 0x300F00Full;
        i = (i | (i << 4)) & 0x30C30C3ull;
        i = (i | (i << 2)) & 0x9249249ull;
        out |= i << k;
    }
    return out;
}
static bool total_less(double a, double b) {
    int64_t x, y; std::memcpy(&x, &a, 8); std::memcpy(&y, &b, 8);
    x ^= (int64_t)((uint64_t)(x >> 63) >> 1); y ^= (int64_t)((uint64_t)(y >> 63) >> 1);
    return x < y;
}
struct BvhSet { std::vector<uint32_t> obj; std::vector<uint64_t> code; };

static bool bvh_split(const std::vector<Box>& boxes, const BvhSet& n, size_t depth, BvhSet& l, BvhSet& r) {
    const size_t cnt = n.obj.size();
    if (cnt <= 1) return false;
    if (depth > SAH_MAX_DEPTH) {                                                                // node.rs:46-72
        const size_t rss = MORTON_BITS - depth;
        const uint64_t first = (n.code[0] >> rss) & 1, last = (n.code.back() >> rss) & 1;
        size_t split;
        if (first == last) { if (cnt > BVH_MAX_LEAF) split = cnt / 2; else return false; }
        else {
            size_t lo = 0, hi = cnt, size = cnt;   // slice::partition_point
            while (lo < hi) { size_t mid = lo + size / 2; if (((n.code[mid] >> rss) & 1) == first) lo = mid + 1; else hi = mid; size = hi - lo; }
            split = lo;
        }
        l.obj.assign(n.obj.begin(), n.obj.begin() + split); l.code.assign(n.code.begin(), n.code.begin() + split);
        r.obj.assign(n.obj.begin() + split, n.obj.end()); r.code.assign(n.code.begin() + split, n.code.end());
        return true;
    }
    // SAH over centre-sorted order per axis (node.rs:74-143)
    double best_cost = kInf, best_center = kInf; int best_axis = 0, best_side = 0;
    std::vector<uint32_t> idx; std::vector<double> al(cnt + 1), ar(cnt + 1), ctr(cnt + 1);
    for (int ax = 0; ax < 3; ax++) {
        idx = n.obj;
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t i, uint32_t j) { return total_less(axis(center(boxes[i]), ax), axis(center(boxes[j]), ax)); });
        al[0] = kInf; ar[0] = kInf;
        Box b = Box::empty();
        for (size_t i = 0; i < cnt; i++) { b = merge(b, boxes[idx[i]]); al[i + 1] = area(b); ctr[i] = axis(center(boxes[idx[i]]), ax); }
        ctr[cnt] = kInf;
        b = Box::empty();
        for (size_t i = 0; i < cnt; i++) { b = merge(b, boxes[idx[cnt - 1 - i]]); ar[i + 1] = area(b); }
        const double total = ar[cnt];
        for (size_t i = 0; i < cnt;) {
            const double c = ctr[i];
            size_t nm = 1;
            while (nm + i <= cnt && c == ctr[i + nm]) nm++;
            const size_t nl = i, nr = cnt - i - nm;
            auto cost = [&](size_t a, size_t bb) {                                              // node.rs:181-210
                double v = BVH_TRAVERSE + BVH_INTERSECT * ((double)a * al[a] + (double)bb * ar[bb]) / total;
                return (a == 0 || bb == 0) ? v * (1.0 - BVH_EMPTY_BONUS) : v;
            };
            const double cl = cost(nl + nm, nr), cr = cost(nl, nm + nr);
            const double cbest = cl < cr ? cl : cr; const int side = cl < cr ? -1 : 1;
            if (cbest < best_cost) { best_cost = cbest; best_axis = ax; best_center = c; best_side = side; }
            i += nm;
        }
    }
    BvhSet left, right;                                                                          // node.rs:145-179
    for (size_t i = 0; i < cnt; i++) {
        const double c = axis(center(boxes[n.obj[i]]), best_axis);
        if (c < best_center || (c == best_center && best_side == -1)) { left.obj.push_back(n.obj[i]); left.code.push_back(n.code[i]); }
        else { right.obj.push_back(n.obj[i]); right.code.push_back(n.code[i]); }
    }
    if (left.obj.empty()) { l = std::move(right); r = std::move(left); } else { l = std::move(left); r = std::move(right); }
    return true;
}

// appends one BVH to B.tlas / B.tlas_leaf; returns root index
static uint32_t build_bvh(Builder& B, const std::vector<Box>& boxes) {
    Box all = Box::empty();
    for (const Box& b : boxes) all = merge(all, b);
    std::vector<std::pair<uint64_t, uint32_t>> codes;
    for (uint32_t i = 0; i < boxes.size(); i++) codes.push_back({morton(all, center(boxes[i])), i});
    std::sort(codes.begin(), codes.end());
    struct Q { BvhSet set; uint32_t parent; bool is_left; size_t depth; };
    std::deque<Q> que;
    { Q q; for (auto& c : codes) { q.set.code.push_back(c.first); q.set.obj.push_back(c.second); } q.parent = LUMO_NONE; q.is_left = true; q.depth = 1; que.push_back(std::move(q)); }
    const uint32_t base = (uint32_t)B.tlas.size();
    std::vector<BvhSet> leaf_sets;   // per node (empty for inner)
    while (!que.empty()) {
        Q q = std::move(que.front()); que.pop_front();
        const uint32_t pos = (uint32_t)B.tlas.size();
        LumoTlasNode node; std::memset(&node, 0, sizeof node); node.right = LUMO_NONE;
        B.tlas.push_back(node);
        if (q.parent != LUMO_NONE && !q.is_left) B.tlas[q.parent].right = pos - base;
        BvhSet l, r;
        if (!bvh_split(boxes, q.set, q.depth, l, r)) { leaf_sets.push_back(std::move(q.set)); continue; }
        leaf_sets.push_back(BvhSet());
        const bool right_nonempty = !r.obj.empty();
        que.push_front(Q{std::move(l), pos, true, q.depth + 1});
        if (right_nonempty) que.push_back(Q{std::move(r), pos, false, q.depth + 1});
    }
    const uint32_t n_nodes = (uint32_t)B.tlas.size() - base;
    for (uint32_t k = n_nodes; k-- > 0;) {                                                      // bvh.rs:282-309
        LumoTlasNode& nd = B.tlas[base + k];
        Box b = Box::empty();
        if (!leaf_sets[k].obj.empty()) {
            for (uint32_t o : leaf_sets[k].obj) b = merge(b, boxes[o]);
        } else {
            const LumoTlasNode& lc = B.tlas[base + k + 1];
            b.lo = v3(lc.lo[0], lc.lo[1], lc.lo[2]); b.hi = v3(lc.hi[0], lc.hi[1], lc.hi[2]);
            if (nd.right != LUMO_NONE) {
                const LumoTlasNode& rc = B.tlas[base + nd.right];
                Box rb = {v3(rc.lo[0], rc.lo[1], rc.lo[2]), v3(rc.hi[0], rc.hi[1], rc.hi[2])};
                b = merge(b, rb);
            }
        }
        nd.lo[0] = b.lo.x; nd.lo[1] = b.lo.y; nd.lo[2] = b.lo.z; nd.hi[0] = b.hi.x; nd.hi[1] = b.hi.y; nd.hi[2] = b.hi.z;
    }
    for (uint32_t k = 0; k < n_nodes; k++) {   // leaf lists in node order; node.right stays relative to this BVH's root
        LumoTlasNode& nd = B.tlas[base + k];
        nd.first = (uint32_t)B.tlas_leaf.size(); nd.count = (uint32_t)leaf_sets[k].obj.size();
        for (uint32_t o : leaf_sets[k].obj) B.tlas_leaf.push_back(o);
    }
    return base;
}

// ---- camera / film ------------------------------------------------------------------------------
static void fill_camera(LumoCamera& C, LumoFilm& F, Cursor& r) {
    V3 origin = v3(0, 0, 0), towards = v3(0, 0, -1), up = v3(0, 1, 0);
    origin.x = r.f(); origin.y = r.f(); origin.z = r.f(); towards.x = r.f(); towards.y = r.f(); towards.z = r.f(); up.x = r.f(); up.y = r.f(); up.z = r.f();
    double zoom = r.f(), lens = r.f(), focal = r.f(), vfov = r.f();
    int64_t rx = r.i(), ry = r.i(), ctype = r.i(), fk = r.i();
    double fr = r.f(), fp = r.f();
    int64_t cs = r.i(), il = r.i();
    Xform cts;
    if (ctype == 0) {                                                                           // camera/matrices.rs:4-14
        Xform proj = Xform::perspective(1e-2, 1e3);
        double tvi = 1.0 / std::tan((vfov * (kPi / 180.0)) / 2.0);
        cts = compose(Xform::scale(tvi, tvi, 1.0), proj);
    } else cts = compose(Xform::scale(1.0, 1.0, 1.0 / (1.0 - 0.0)), Xform::translation(0.0, 0.0, -0.0));
    V3 fwd = normalize(sub(towards, origin));                                                   // matrices.rs:24-37
    V3 right = normalize(cross(fwd, up));
    V3 up2 = cross(right, fwd);
    M3 basis = {{right.x, right.y, right.z, up2.x, up2.y, up2.z, fwd.x, fwd.y, fwd.z}};
    Xform wtc = compose(Xform::translation(-dot(origin, right), -dot(origin, up2), -dot(origin, fwd)), Xform::from_m3(basis));
    double ar = (double)rx / (double)ry;                                                        // matrices.rs:39-70
    double smin_x, smin_y, smax_x, smax_y;
    if (ar > 1.0) { smin_x = -ar; smin_y = -1.0; smax_x = ar; smax_y = 1.0; } else { smin_x = -1.0; smin_y = -1.0 / ar; smax_x = 1.0; smax_y = 1.0 / ar; }
    double sdx = smax_x - smin_x, sdy = smax_y - smin_y;
    Xform str = compose(compose(compose(Xform::scale((double)rx, -((double)ry), 1.0), Xform::scale(1.0 / sdx, 1.0 / sdy, 1.0)),
                                Xform::translation(-smin_x, -smax_y, 0.0)), Xform::scale(zoom, zoom, zoom));
    std::memcpy(C.screen_to_raster_m, str.m.a, 128); std::memcpy(C.screen_to_raster_inv, str.inv.a, 128);
    std::memcpy(C.camera_to_screen_m, cts.m.a, 128); std::memcpy(C.camera_to_screen_inv, cts.inv.a, 128);
    std::memcpy(C.world_to_camera_m, wtc.m.a, 128); std::memcpy(C.world_to_camera_inv, wtc.inv.a, 128);
    V3 pmin = apply(cts.inv, apply(str.inv, v3(0, 0, 0), 1.0), 1.0);                            // camera.rs:51-67
    V3 pmax = apply(cts.inv, apply(str.inv, v3((double)rx, (double)ry, 0.0), 1.0), 1.0);
    double zmin = pmin.z == 0.0 ? 1.0 : pmin.z, zmax = pmax.z == 0.0 ? 1.0 : pmax.z;
    double dx = pmax.x / zmax - pmin.x / zmin, dy = pmax.y / zmax - pmin.y / zmin;
    C.image_plane_area = std::fabs(dx * dy);
    C.lens_radius = lens; C.focal_length = focal;
    C.lens_area = lens == 0.0 ? 1.0 : kPi * (lens * lens);
    C.res_x = (uint32_t)rx; C.res_y = (uint32_t)ry; C.ortho = ctype == 1; C.pad = 0;
    // film: colour space + white balance (color/space.rs:50-177)
    double wx, wy; to_xy(dense_to_xyz(spectra::D65), wx, wy);
    V3 W = from_xyY(wx, wy, 1.0);
    static const double prim[3][6] = {{0.64, 0.33, 0.3, 0.6, 0.15, 0.06}, {0.68, 0.32, 0.265, 0.69, 0.15, 0.06}, {0.708, 0.292, 0.170, 0.797, 0.131, 0.046}};
    const double* pr = prim[cs];
    V3 R = from_xyY(pr[0], pr[1], 1.0), G = from_xyY(pr[2], pr[3], 1.0), Bp = from_xyY(pr[4], pr[5], 1.0);
    M3 rgbc = transpose3(M3{{R.x, R.y, R.z, G.x, G.y, G.z, Bp.x, Bp.y, Bp.z}});
    V3 Cc = mulv3(inv3(rgbc), W);
    M3 x2r = inv3(matmul3(rgbc, M3{{Cc.x, 0, 0, 0, Cc.y, 0, 0, 0, Cc.z}}));
    std::memcpy(F.xyz_to_rgb, x2r.a, 72);
    const M3 x2l = {{0.210576, 0.855098, -0.0396983, -0.417076, 1.177260, 0.0786283, 0.0, 0.0, 0.5168350}};
    const M3 l2x = inv3(x2l);
    double ix, iy; to_xy(dense_to_xyz(illuminant((int)il)), ix, iy);
    V3 num = mulv3(x2l, W), den = mulv3(x2l, from_xyY(ix, iy, 1.0));
    M3 dg = {{num.x / den.x, 0, 0, 0, num.y / den.y, 0, 0, 0, num.z / den.z}};
    M3 wb = matmul3(matmul3(l2x, dg), x2l);
    std::memcpy(F.wb, wb.a, 72);
    F.filter_kind = (uint32_t)fk; F.filter_r = fr; F.filter_p = fp;
    F.filter_gr = lm_exp(-(fr * fr) / (2.0 * fp * fp)) / std::sqrt(std::fmax(2.0 * kPi * fp * fp, 0.0));   // filter.rs:118-123
    F.r_disc = (uint32_t)sat_u64(std::ceil(fr - 0.5)); F.color_space = (uint32_t)cs; F.pad = 0;
}

// ---- Perlin::new (src/perlin.rs:31-46) ------------------------------------------------------------
// 256 lattice normals = square_to_sphere of consecutive Xorshift pairs, then three Fisher-Yates permutations from the
// same stream (src/rng.rs:39-116).  Appended to the f64 pool: 768 lattice values, then perm x, y, z as f64.
struct Xorshift {
    uint64_t hi, lo;
    explicit Xorshift(uint64_t seed) { lo = seed > 1 ? seed : 1; hi = lo; step(); step(); step(); }               // rng.rs:40-48
    uint64_t step() { uint64_t l = lo, h = hi; hi = l; h ^= h << 23; h ^= h >> 17; h ^= l; lo = h + l; return h; }  // rng.rs:50-59
    double gen_float() { double v = (double)step() * std::ldexp(1.0, -64); const double lim = 1.0 - std::ldexp(1.0, -52); return v < lim ? v : lim; }   // rng.rs:71-75
};
static uint64_t add_perlin(std::vector<double>& pool, uint64_t seed) {
    const uint64_t at = pool.size();
    Xorshift rng(seed);
    for (int i = 0; i < 256; i++) {
        const double rx = rng.gen_float(), ry = rng.gen_float();
        const double z = 1.0 - 2.0 * ry;                                                        // rng/maps.rs:49-55
        const double rr = std::sqrt(std::fmax(1.0 - z * z, 0.0));
        const double phi = 2.0 * kPi * rx;
        pool.push_back(rr * lm_cos(phi)); pool.push_back(rr * lm_sin(phi)); pool.push_back(z);
    }
    for (int d = 0; d < 3; d++) {
        uint64_t perm[256];
        for (uint64_t i = 0; i < 256; i++) perm[i] = i;
        for (uint64_t i = 0; i + 1 < 256; i++) { const uint64_t rnd = rng.step(); const uint64_t j = i + rnd % (256 - i); std::swap(perm[i], perm[j]); }   // rng.rs:104-116
        for (uint64_t v : perm) pool.push_back((double)v);
    }
    return at;
}

// ---- driver -------------------------------------------------------------------------------------
static bool run(const uint8_t* data, uint64_t len, std::vector<uint8_t>& out) {
    if (len < 16 || std::memcmp(data, "LUMOPRG1", 8) != 0) { g_err = "scene program: bad magic"; return false; }
    uint32_t nrec; std::memcpy(&nrec, data + 12, 4);
    Builder B; std::memset(&B.params, 0, sizeof B.params);
    // fixed tables
    const double* fixed[9] = {spectra::X, spectra::Y, spectra::Z, spectra::A, spectra::D50, spectra::D65, spectra::F2, spectra::F7, spectra::CORNELL};
    for (auto t : fixed) { B.tables.insert(B.tables.end(), t, t + 95); B.tables.push_back(0.0); }
    bool have_camera = false;
    struct LightArea { double area; };
    size_t off = 16;
    auto push_material = [&](int64_t kind, double rough, int64_t ek, double ec, int64_t kk, double kc, const float* kd, const float* ks, const float* tf,
                             const float* ke, int64_t illum, double scale, int64_t two_sided, const int64_t* tex /* kd ks tf ke bump, or null */) {
        LumoMaterial M; std::memset(&M, 0, sizeof M);
        uint32_t* slots[5] = {&M.kd_tex, &M.ks_tex, &M.tf_tex, &M.ke_tex, &M.bump_tex};
        for (int i = 0; i < 5; i++) {
            const int64_t t = tex ? tex[i] : -1;
            if (t >= (int64_t)B.textures.size()) throw std::runtime_error("material refers to an undefined texture");
            if (t >= 0 && ((i == 4) != (B.textures[t].kind == LTEX_BUMP))) throw std::runtime_error("material texture slot / texture kind mismatch");
            *slots[i] = t < 0 ? LUMO_NONE : (uint32_t)t;
        }
        M.kind = (uint32_t)kind; M.roughness = std::fmax(rough, 1e-5);
        auto table_for = [&](int64_t k, double c) -> uint32_t {
            double tmp[95];
            const double* src = tmp;
            switch (k) { case 1: src = spectra::glass_eta; break; case 2: src = spectra::diamond_eta; break; case 3: src = spectra::mirror_eta; break;
                         case 4: src = spectra::mirror_k; break; default: for (double& v : tmp) v = c; }
            return B.add_table(src);
        };
        M.eta_table = table_for(ek, ec); M.k_table = table_for(kk, kc);
        const double* et = &B.tables[96 * M.eta_table];
        bool is_const = true; for (int i = 1; i < 95; i++) if (et[i] != et[0]) is_const = false;       // dense_spectrum.rs:23-27
        if (is_const) M.flags |= LMF_ETA_CONST;
        if (two_sided) M.flags |= LMF_TWO_SIDED;
        std::memcpy(M.kd, kd, 16); std::memcpy(M.ks, ks, 16); std::memcpy(M.tf, tf, 16); std::memcpy(M.ke, ke, 16);
        M.illum_table = LTAB_ILLUM0 + (uint32_t)illum; M.scale = scale;
        B.materials.push_back(M);
    };
    for (uint32_t rec = 0; rec < nrec; rec++) {
        uint32_t tag; uint64_t nbytes;
        if (off + 16 > len) { g_err = "scene program: truncated"; return false; }
        std::memcpy(&tag, data + off, 4); std::memcpy(&nbytes, data + off + 8, 8);
        Cursor r{data + off + 16};
        off += 16 + nbytes;
        if (off > len) { g_err = "scene program: truncated record"; return false; }
        if (tag == TAG_MATERIAL) {
            int64_t kind = r.i(); double rough = r.f(); int64_t ek = r.i(); double ec = r.f(); int64_t kk = r.i(); double kc = r.f();
            float sp[4][4]; for (auto& s : sp) for (float& v : s) v = (float)r.f();
            int64_t illum = r.i(); double scale = r.f(); int64_t ts = r.i();
            int64_t tex[5] = {-1, -1, -1, -1, -1};
            if (nbytes >= 30 * 8) for (int64_t& t : tex) t = r.i();
            push_material(kind, rough, ek, ec, kk, kc, sp[0], sp[1], sp[2], sp[3], illum, scale, ts, tex);
        } else if (tag == TAG_TEXTURE) {
            LumoTexture T; std::memset(&T, 0, sizeof T);
            T.kind = (uint32_t)r.i();
            for (float& v : T.spec) v = (float)r.f();
            const int64_t a = r.i(), b = r.i(); T.scale = r.f();
            const int64_t seed = r.i(), w = r.i(), h = r.i();
            T.a = T.b = LUMO_NONE;
            if (T.kind == LTEX_CHECKER) {
                if (a < 0 || b < 0 || a >= (int64_t)B.textures.size() || b >= (int64_t)B.textures.size()) throw std::runtime_error("checkerboard refers to an undefined texture");
                if (B.textures[a].kind == LTEX_BUMP || B.textures[b].kind == LTEX_BUMP) throw std::runtime_error("checkerboard of a bump map");
                T.a = (uint32_t)a; T.b = (uint32_t)b;
            } else if (T.kind == LTEX_MARBLE) {
                T.data = add_perlin(B.tex_f64, (uint64_t)seed);
            } else if (T.kind == LTEX_IMAGE || T.kind == LTEX_BUMP) {
                const int64_t per = T.kind == LTEX_IMAGE ? 4 : 3;
                if (w <= 0 || h <= 0 || (uint64_t)(11 * 8 + w * h * per * 8) > nbytes) throw std::runtime_error("image texture: bad size");
                T.width = (uint32_t)w; T.height = (uint32_t)h;
                if (T.kind == LTEX_IMAGE) { T.data = B.tex_pixels.size() / 4; for (int64_t i = 0; i < w * h * 4; i++) B.tex_pixels.push_back((float)r.f()); }
                else { T.data = B.tex_f64.size(); for (int64_t i = 0; i < w * h * 3; i++) B.tex_f64.push_back(r.f()); }
            } else if (T.kind != LTEX_SOLID && T.kind != LTEX_MANDELBROT) throw std::runtime_error("unknown texture kind");
            B.textures.push_back(T);
        } else if (tag == TAG_MESH) {
            MeshIn m;
            m.nv = r.i(); m.nn = r.i(); m.nt = r.i(); m.nf = r.i(); m.nc = r.i(); int64_t hn = r.i(), ht = r.i();
            m.v = (const double*)r.p; r.p += 24 * m.nv;
            m.n = (const double*)r.p; r.p += 24 * m.nn;
            m.t = (const double*)r.p; r.p += 16 * m.nt;
            m.off = (const int64_t*)r.p; r.p += 8 * (m.nf + 1);
            m.vi = (const int64_t*)r.p; r.p += 8 * m.nc;
            m.ni = hn ? (const int64_t*)r.p : nullptr; if (hn) r.p += 8 * m.nc;
            m.ti = ht ? (const int64_t*)r.p : nullptr; if (ht) r.p += 8 * m.nc;
            m.normal_base = (uint32_t)(B.normals.size() / 3); m.uv_base = (uint32_t)(B.uvs.size() / 2);
            B.normals.insert(B.normals.end(), m.n, m.n + 3 * m.nn);
            B.uvs.insert(B.uvs.end(), m.t, m.t + 2 * m.nt);
            B.meshes.push_back(m);
        } else if (tag == TAG_OBJECT) {
            int64_t kind = r.i(), is_light = r.i(), mat = r.i(), mesh = r.i(), f0 = r.i(), f1 = r.i();
            double prm[9]; for (double& v : prm) v = r.f();
            int64_t inst_mat = r.i(), n_ops = r.i();
            struct Op { int64_t op; double x, y, z; };
            std::vector<Op> ops((size_t)n_ops);
            for (Op& o : ops) { o.op = r.i(); o.x = r.f(); o.y = r.f(); o.z = r.f(); }
            // base geometry -> list of (object record, local box, area)
            struct Base { LumoObject o; Box box; double area; };
            std::vector<Base> bases;
            auto blank = [&]() { LumoObject o; std::memset(&o, 0, sizeof o); o.inst = -1; o.material = (int32_t)mat; return o; };
            auto tree_box = [&](uint32_t kd) { const LumoKdTree& t = B.kd_trees[kd]; Box b = {v3(t.lo[0], t.lo[1], t.lo[2]), v3(t.hi[0], t.hi[1], t.hi[2])}; return b; };
            if (kind == OBJ_KDMESH) {
                if (mesh < 0 || mesh >= (int64_t)B.meshes.size()) { g_err = "object: bad mesh id"; return false; }
                auto key = std::make_tuple(mesh, f0, f1);
                auto it = B.kd_cache.find(key);
                uint32_t kd;
                if (it != B.kd_cache.end()) kd = it->second;
                else {
                    uint32_t first, count; emit_triangles(B, B.meshes[mesh], f0, f1, first, count);
                    if (count == 0) { g_err = "object: mesh chunk has no non-degenerate triangles"; return false; }
                    kd = build_kd(B, first, count); B.kd_cache[key] = kd;
                }
                Base b; b.o = blank(); b.o.kind = LOBJ_KD; b.o.geom = kd; b.box = tree_box(kd); b.area = 0.0; bases.push_back(b);
            } else if (kind == OBJ_RECT) {                                                       // rectangle.rs:27-42
                V3 a = v3(prm[0], prm[1], prm[2]), bq = v3(prm[3], prm[4], prm[5]), c = v3(prm[6], prm[7], prm[8]);
                V3 origin = bq, b0 = sub(c, origin), b1 = sub(a, origin);
                V3 d = add(add(origin, b0), b1);
                double quad[12] = {a.x, a.y, a.z, bq.x, bq.y, bq.z, c.x, c.y, c.z, d.x, d.y, d.z};
                int64_t qoff[2] = {0, 4}, qvi[4] = {0, 1, 2, 3};
                MeshIn m; m.v = quad; m.n = nullptr; m.t = nullptr; m.nv = 4; m.nn = m.nt = 0; m.nf = 1; m.nc = 4; m.off = qoff; m.vi = qvi; m.ni = nullptr; m.ti = nullptr;
                uint32_t first, count; emit_triangles(B, m, 0, 1, first, count);
                if (count == 0) { g_err = "rectangle is degenerate"; return false; }
                uint32_t kd = build_kd(B, first, count);
                LumoRect R; std::memset(&R, 0, sizeof R);
                R.origin[0] = origin.x; R.origin[1] = origin.y; R.origin[2] = origin.z; R.b0[0] = b0.x; R.b0[1] = b0.y; R.b0[2] = b0.z; R.b1[0] = b1.x; R.b1[1] = b1.y; R.b1[2] = b1.z;
                B.rects.push_back(R);
                Base b; b.o = blank(); b.o.kind = LOBJ_RECT; b.o.geom = kd; b.o.rect = (uint32_t)B.rects.size() - 1;
                V3 pa = add(b1, origin), pb = origin, pc = add(b0, origin), pd = add(add(origin, b0), b1);   // rectangle.rs:91-101
                b.box.lo = vmin(vmin(vmin(pa, pb), pc), pd); b.box.hi = vmax(vmax(vmax(pa, pb), pc), pd);
                b.area = std::fabs(length(cross(b0, b1)));                                       // rectangle.rs:107-109
                bases.push_back(b);
            } else if (kind == OBJ_SPHERE) {
                if (prm[0] == 0.0) { g_err = "sphere radius must be non-zero"; return false; }  // sphere.rs:17
                LumoSphere S = {prm[0], 0.0}; B.spheres.push_back(S);
                Base b; b.o = blank(); b.o.kind = LOBJ_SPHERE; b.o.geom = (uint32_t)B.spheres.size() - 1;
                b.box.lo = v3(-prm[0], -prm[0], -prm[0]); b.box.hi = v3(prm[0], prm[0], prm[0]);
                b.area = 4.0 * kPi * prm[0] * prm[0];
                bases.push_back(b);
            } else if (kind == OBJ_LOOSE_TRIS) {
                if (mesh < 0 || mesh >= (int64_t)B.meshes.size()) { g_err = "object: bad mesh id"; return false; }
                uint32_t first, count; emit_triangles(B, B.meshes[mesh], f0, f1, first, count);
                for (uint32_t i = 0; i < count; i++) {
                    const LumoTriVerts& t = B.tri_verts[first + i];
                    V3 a = v3(t.a[0], t.a[1], t.a[2]), bb = v3(t.b[0], t.b[1], t.b[2]), c = v3(t.c[0], t.c[1], t.c[2]);
                    Base b; b.o = blank(); b.o.kind = LOBJ_TRI; b.o.geom = first + i;
                    b.box.lo = vmin(a, vmin(bb, c)); b.box.hi = vmax(a, vmax(bb, c));
                    b.area = length(cross(sub(bb, a), sub(c, a))) / 2.0;                         // triangle.rs:208-210
                    bases.push_back(b);
                }
            } else { g_err = "object: unknown kind"; return false; }
            for (Base& b : bases) {
                Box box = b.box; double ar = b.area;
                if (n_ops > 0) {
                    Xform t = Xform::identity();
                    for (const Op& o : ops) {
                        Xform step = Xform::identity();
                        switch (o.op) {
                        case OP_UNIT: { double s = 1.0 / max_element(sub(b.box.hi, b.box.lo)); step = Xform::scale(s, s, s); break; }   // kdtree.rs:93-99
                        case OP_ORIGIN: { Box cur = xform_box(t, b.box); V3 mid = divs(neg(add(cur.lo, cur.hi)), 2.0); step = Xform::translation(mid.x, mid.y, mid.z); break; }
                        case OP_SETX: { Box cur = xform_box(t, b.box); step = Xform::translation(o.x - cur.lo.x, 0.0, 0.0); break; }
                        case OP_SETY: { Box cur = xform_box(t, b.box); step = Xform::translation(0.0, o.x - cur.lo.y, 0.0); break; }
                        case OP_SETZ: { Box cur = xform_box(t, b.box); step = Xform::translation(0.0, 0.0, o.x - cur.lo.z); break; }
                        case OP_TRANSLATE: step = Xform::translation(o.x, o.y, o.z); break;
                        case OP_SCALE: if (o.x * o.y * o.z == 0.0) { g_err = "scale by zero"; return false; } step = Xform::scale(o.x, o.y, o.z); break;
                        case OP_ROTX: step = Xform::rot_x(o.x); break;
                        case OP_ROTY: step = Xform::rot_y(o.x); break;
                        case OP_ROTZ: step = Xform::rot_z(o.x); break;
                        default: g_err = "object: unknown transform op"; return false;
                        }
                        t = compose(step, t);                                                    // instance.rs:257-299: T * current
                    }
                    b.o.inst = (int32_t)add_instance(B, t);
                    if (inst_mat >= 0) b.o.material = (int32_t)inst_mat;
                    box = xform_box(t, b.box);
                    M3 mt = transpose3(to3(t.m));                                                // transform.rs:74-83 to_scale
                    double sx = length(v3(mt.a[0], mt.a[1], mt.a[2])), sy = length(v3(mt.a[3], mt.a[4], mt.a[5]));
                    ar = sx * sy * b.area;                                                       // instance.rs:134-144
                }
                if (is_light) { B.light_objects.push_back(b.o); B.light_boxes.push_back(box); B.light_area.push_back(ar); B.light_power_mat.push_back((int32_t)mat); }
                else { B.objects.push_back(b.o); B.object_boxes.push_back(box); }
            }
        } else if (tag == TAG_ENVMAP) {
            for (float& v : B.env_spec) v = (float)r.f();
            B.env_scale = r.f(); B.have_env = true;
            if (nbytes >= 6 * 8) B.env_tex = r.i();
        } else if (tag == TAG_CAMERA) {
            fill_camera(B.params.camera, B.params.film, r); have_camera = true;
        }
    }
    if (!have_camera) { g_err = "scene program has no camera record"; return false; }
    if (B.objects.empty()) { g_err = "scene has no objects"; return false; }
    if (B.light_objects.empty() && !B.have_env) { g_err = "scene has no lights"; return false; }      // renderer.rs:42
    for (const LumoObject& o : B.objects) if (o.material < 0 || o.material >= (int32_t)B.materials.size()) { g_err = "object without material"; return false; }

    finish_kd_trees(B);
    // Scene::build (scene.rs:33-52)
    const uint32_t obj_root = build_bvh(B, B.object_boxes);
    (void)obj_root;
    Box bounds = Box::empty();
    for (const Box& b : B.object_boxes) bounds = merge(bounds, b);
    Box lb = Box::empty();
    for (const Box& b : B.light_boxes) lb = merge(lb, b);
    bounds = merge(bounds, lb);
    if (B.have_env) {
        V3 c = center(bounds);
        double radius = length(sub(c, bounds.lo));   // Vec3::distance
        float z4[4] = {0, 0, 0, 0};
        const int64_t env_tex[5] = {-1, -1, -1, B.env_tex, -1};
        push_material(LMAT_LIGHT, 1.0, 0, 0.0, 0, 0.0, z4, z4, z4, B.env_spec, 2 /* D65 */, B.env_scale, 1, env_tex);   // scene.rs:73-77
        LumoSphere S = {radius, 0.0}; B.spheres.push_back(S);
        Xform t = compose(Xform::translation(c.x, c.y, c.z), Xform::identity());
        LumoObject o; std::memset(&o, 0, sizeof o); o.kind = LOBJ_SPHERE; o.geom = (uint32_t)B.spheres.size() - 1;
        o.inst = (int32_t)add_instance(B, t); o.material = (int32_t)B.materials.size() - 1;
        Box sb = {v3(-radius, -radius, -radius), v3(radius, radius, radius)};
        Box wbx = xform_box(t, sb);
        bounds = merge(bounds, wbx);
        B.light_objects.push_back(o); B.light_boxes.push_back(wbx);
        B.light_area.push_back(1.0 * 1.0 * (4.0 * kPi * radius * radius)); B.light_power_mat.push_back(o.material);
    }
    const uint32_t light_root = build_bvh(B, B.light_boxes);

    // alias table (bvh.rs:105-166); power at ColorWavelength::default() = sample(0.0)
    {
        const size_t n = B.light_objects.size();
        double lam[4], pdfl[4];
        for (int i = 0; i < 4; i++) { double v = 0.0 + (double)i / 4.0; v = v > 1.0 ? v - 1.0 : v; lam[i] = lambda_sample_one(v); pdfl[i] = lambda_pdf_one(lam[i]); }
        std::vector<double> ap(n), pdf(n), prob(n, 1.0); std::vector<uint32_t> alias(n);
        double sum = 0.0;
        for (size_t i = 0; i < n; i++) {
            const LumoMaterial& M = B.materials[B.light_power_mat[i]];
            double acc = 0.0;
            for (int k = 0; k < 4; k++) {
                double phi = 0.0;
                if (M.kind == LMAT_LIGHT) {                                                      // material.rs:234-242
                    const float* ke = M.ke;
                    if (M.ke_tex != LUMO_NONE) {                                                 // Texture::power (texture.rs:95-101)
                        const LumoTexture& T = B.textures[M.ke_tex];
                        if (T.kind != LTEX_SOLID && T.kind != LTEX_IMAGE) { g_err = "light texture has no power (only Solid and Image do, texture.rs:95-101)"; return false; }
                        ke = T.spec;
                    }
                    phi = M.scale * spectrum_sample(ke, lam[k]) * dense_sample(&B.tables[96 * M.illum_table], lam[k]);
                    if (M.flags & LMF_TWO_SIDED) phi = 2.0 * phi;
                }
                double pw = B.light_area[i] * phi;
                acc += pdfl[k] == 0.0 ? 0.0 : pw / pdfl[k];
            }
            ap[i] = acc / 4.0; sum += ap[i]; alias[i] = (uint32_t)i;
        }
        std::vector<size_t> large, small;
        const double pu = 1.0 / (double)n;
        for (size_t i = 0; i < n; i++) { ap[i] /= sum; pdf[i] = ap[i]; (ap[i] > pu ? large : small).push_back(i); }
        size_t is = small.size(), il = large.size();
        while (is > 0 && il > 0) {
            is--; il--;
            const size_t s = small[is], l = large[il];
            prob[s] = pdf[s] * (double)n; alias[s] = (uint32_t)l;
            pdf[l] += pdf[s] - pu;
            if (pdf[l] > pu) { large[il] = l; il++; } else { small[is] = l; is++; }
        }
        while (is > 0) { is--; prob[small[is]] = 1.0; }
        while (il > 0) { il--; prob[large[il]] = 1.0; }
        for (size_t i = 0; i < n; i++) {
            LumoLight L; std::memset(&L, 0, sizeof L);
            L.alias_prob = prob[i]; L.pdf = ap[i]; L.area = B.light_area[i]; L.alias = alias[i];
            const Box& b = B.light_boxes[i];
            const V3 c = center(b);
            const double half = length(sub(b.hi, c));
            const double mag = std::fmax(std::fmax(std::fabs(c.x), std::fabs(c.y)), std::fabs(c.z)) + half;
            L.bound_c[0] = c.x; L.bound_c[1] = c.y; L.bound_c[2] = c.z;
            L.bound_r = half * (1.0 + 1e-4) + 1e-6 * mag + 1e-12;
            B.lights.push_back(L);
        }
    }

    LumoSceneParams& P = B.params;
    P.n_objects = (uint32_t)B.objects.size(); P.n_lights = (uint32_t)B.light_objects.size(); P.lights_root = light_root;
    { uint32_t n = P.n_lights, lg = 0; while ((n >> (lg + 1)) != 0) lg++; P.n_shadow_rays = lg < 1 ? 1 : lg; }   // scene.rs:90-92
    P.n_tlas_nodes = (uint32_t)B.tlas.size(); P.n_kd_trees = (uint32_t)B.kd_trees.size(); P.n_materials = (uint32_t)B.materials.size(); P.n_tris = (uint32_t)B.tri_verts.size();
    P.bounds_lo[0] = bounds.lo.x; P.bounds_lo[1] = bounds.lo.y; P.bounds_lo[2] = bounds.lo.z; P.bounds_hi[0] = bounds.hi.x; P.bounds_hi[1] = bounds.hi.y; P.bounds_hi[2] = bounds.hi.z;

    // make kd inner-node child links and leaf list offsets absolute (they already are: nodes/leaf lists were appended globally)
    std::vector<LumoObject> all_objects = B.objects;
    all_objects.insert(all_objects.end(), B.light_objects.begin(), B.light_objects.end());

    // ---- order-free occlusion structure (ah_bvh.h): every primitive of both object lists, in world space ------------
    AhBuilder ah;
    {
        auto world_vertex = [&](const double* v, int32_t inst) {
            if (inst < 0) return v3(v[0], v[1], v[2]);
            const double* m = B.instances[(size_t)inst].m;
            return v3(m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3], m[4] * v[0] + m[5] * v[1] + m[6] * v[2] + m[7], m[8] * v[0] + m[9] * v[1] + m[10] * v[2] + m[11]);
        };
        auto add_tri = [&](uint32_t ti, uint32_t g, int32_t inst) {
            const LumoTriVerts& T = B.tri_verts[ti];
            const V3 a = world_vertex(T.a, inst), b = world_vertex(T.b, inst), c = world_vertex(T.c, inst);
            Box bb = {vmin(vmin(a, b), c), vmax(vmax(a, b), c)};
            ah.add(bb, ti, g | (inst >= 0 ? LUMO_AH_INSTANCED : 0u));
        };
        for (size_t g = 0; g < all_objects.size(); g++) {
            const LumoObject& o = all_objects[g];
            if (o.kind == LOBJ_KD || o.kind == LOBJ_RECT) {
                const LumoKdTree& T = B.kd_trees[o.geom];
                for (uint32_t k = 0; k < T.n_tris; k++) add_tri(T.tri_base + k, (uint32_t)g, o.inst);
            } else if (o.kind == LOBJ_TRI) add_tri(o.geom, (uint32_t)g, o.inst);
            else {   // sphere: the object's own world box (instance.rs:107-128 / sphere.rs bounding box)
                const Box& wb = g < B.objects.size() ? B.object_boxes[g] : B.light_boxes[g - B.objects.size()];
                ah.add(wb, LUMO_AH_SPHERE | o.geom, (uint32_t)g | (o.inst >= 0 ? LUMO_AH_INSTANCED : 0u));
            }
        }
        if (ah.prims.size() >= (1u << 27)) { g_err = "too many primitives for the occlusion BVH (2^27)"; return false; }
        const bool timing = std::getenv("LUMO_HOST_TIMING") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        ah.build_binary();
        const auto t1 = std::chrono::steady_clock::now();
        ah.collapse();
        const auto t2 = std::chrono::steady_clock::now();
        if (timing) std::fprintf(stderr, "[lumo_host] kd-trees %.2f s; world-space BVH over %zu primitives: binary build %.2f s, collapse %.2f s\n", g_kd_seconds, ah.prims.size(),
                                 std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count());
        g_kd_seconds = 0.0;
    }
    // per object: its chain of object-BVH nodes, root first (the confirming traversal re-runs the reference's box tests on it)
    std::vector<uint32_t> path_off(all_objects.size() + 1, 0), path_nodes;
    {
        std::vector<std::vector<uint32_t>> paths(all_objects.size());
        auto walk = [&](uint32_t root, uint32_t n_nodes, uint32_t obj_base) {
            if (n_nodes == 0) return;
            struct Fr { uint32_t node; std::vector<uint32_t> chain; };
            std::vector<Fr> st; st.push_back({0u, {}});
            while (!st.empty()) {
                Fr f = std::move(st.back()); st.pop_back();
                for (;;) {
                    f.chain.push_back(root + f.node);
                    const LumoTlasNode& nd = B.tlas[root + f.node];
                    if (nd.count > 0) { for (uint32_t k = 0; k < nd.count; k++) paths[obj_base + B.tlas_leaf[nd.first + k]] = f.chain; break; }
                    if (nd.right != LUMO_NONE) st.push_back({nd.right, f.chain});
                    f.node += 1;
                    if (f.node >= n_nodes) break;
                }
            }
        };
        walk(0, light_root, 0);
        walk(light_root, (uint32_t)B.tlas.size() - light_root, (uint32_t)B.objects.size());
        for (size_t g = 0; g < all_objects.size(); g++) { path_off[g] = (uint32_t)path_nodes.size(); path_nodes.insert(path_nodes.end(), paths[g].begin(), paths[g].end()); }
        path_off[all_objects.size()] = (uint32_t)path_nodes.size();
    }

    // assemble
    struct Sec { const void* p; uint64_t bytes, count; };
    Sec secs[LSEC_COUNT] = {
        {B.tlas.data(), B.tlas.size() * sizeof(LumoTlasNode), B.tlas.size()},
        {B.tlas_leaf.data(), B.tlas_leaf.size() * 4, B.tlas_leaf.size()},
        {all_objects.data(), all_objects.size() * sizeof(LumoObject), all_objects.size()},
        {B.instances.data(), B.instances.size() * sizeof(LumoInstance), B.instances.size()},
        {B.kd_trees.data(), B.kd_trees.size() * sizeof(LumoKdTree), B.kd_trees.size()},
        {B.kd_nodes.data(), B.kd_nodes.size() * sizeof(LumoKdNode), B.kd_nodes.size()},
        {B.kd_leaf.data(), B.kd_leaf.size() * 4, B.kd_leaf.size()},
        {B.tri_verts.data(), B.tri_verts.size() * sizeof(LumoTriVerts), B.tri_verts.size()},
        {B.tri_shade.data(), B.tri_shade.size() * sizeof(LumoTriShade), B.tri_shade.size()},
        {B.normals.data(), B.normals.size() * 8, B.normals.size() / 3},
        {B.uvs.data(), B.uvs.size() * 8, B.uvs.size() / 2},
        {B.rects.data(), B.rects.size() * sizeof(LumoRect), B.rects.size()},
        {B.spheres.data(), B.spheres.size() * sizeof(LumoSphere), B.spheres.size()},
        {B.materials.data(), B.materials.size() * sizeof(LumoMaterial), B.materials.size()},
        {B.tables.data(), B.tables.size() * 8, B.tables.size() / 96},
        {B.lights.data(), B.lights.size() * sizeof(LumoLight), B.lights.size()},
        {B.textures.data(), B.textures.size() * sizeof(LumoTexture), B.textures.size()},
        {B.tex_pixels.data(), B.tex_pixels.size() * 4, B.tex_pixels.size() / 4},
        {B.tex_f64.data(), B.tex_f64.size() * 8, B.tex_f64.size()},
        {ah.nodes.data(), ah.nodes.size() * sizeof(LumoAhNode), ah.nodes.size()},
        {ah.out_prims.data(), ah.out_prims.size() * sizeof(LumoAhPrim), ah.out_prims.size()},
        {path_off.data(), path_off.size() * 4, path_off.size()},
        {path_nodes.data(), path_nodes.size() * 4, path_nodes.size()},
    };
    LumoBlobHeader H; std::memset(&H, 0, sizeof H);
    H.magic = LUMO_BLOB_MAGIC; H.version = LUMO_BLOB_VERSION; H.n_sections = LSEC_COUNT; H.params = P;
    uint64_t pos = (sizeof H + 255) & ~(uint64_t)255;
    for (int s = 0; s < LSEC_COUNT; s++) { H.sec[s].offset = pos; H.sec[s].bytes = secs[s].bytes; H.sec[s].count = secs[s].count; pos = (pos + secs[s].bytes + 255) & ~(uint64_t)255; }
    H.total_bytes = pos;
    out.assign(pos, 0);
    std::memcpy(out.data(), &H, sizeof H);
    for (int s = 0; s < LSEC_COUNT; s++) if (secs[s].bytes) std::memcpy(out.data() + H.sec[s].offset, secs[s].p, secs[s].bytes);
    return true;
}

}  // namespace lumo_host

extern "C" {
// Builds the device blob from a scene program.  *blob is malloc'd (free with lumo_host_free).
// Returns 0 on success, -1 on error (message: lumo_host_last_error()).  Never throws.
int32_t lumo_host_build(const void* program, uint64_t len, void** blob, uint64_t* blob_len) {
    try {
        std::vector<uint8_t> out;
        if (!lumo_host::run((const uint8_t*)program, len, out)) return -1;
        void* p = std::malloc(out.size());
        if (!p) { lumo_host::g_err = "out of memory"; return -1; }
        std::memcpy(p, out.data(), out.size());
        *blob = p; *blob_len = out.size();
        return 0;
    } catch (const std::exception& e) { lumo_host::g_err = std::string("host build: ") + e.what(); return -1; }
    catch (...) { lumo_host::g_err = "host build: unknown exception"; return -1; }
}
void lumo_host_free(void* p) { std::free(p); }
// Regenerates the payload of the reference's `srgb.coeff` (spectrum/tables.rs:6-84): scale[64] then data[3 * 64^3 * 3], f32.
// illum_div <= 0: normalise D65 so that white has Y = 1 under the optimiser's own quadrature.  Takes some 10 s per core-minute.
int32_t lumo_host_srgb_table(float* scale64, float* data, int32_t threads, double illum_div, double clamp_max) {
    try { lumo_host::srgb_table::generate(scale64, data, threads, illum_div, clamp_max); return 0; }
    catch (...) { lumo_host::g_err = "srgb table: exception"; return -1; }
}
// lumo_math.h on the host (fn as in lumo_gpu_math_eval): the host-side film finalisation (lumo_b200/color.py: encode) takes its
// transfer-curve pow from here so that it is the device's, bit for bit.
void lumo_host_math_eval(int32_t fn, const double* x, const double* y, uint64_t n, double* out) {
    for (uint64_t i = 0; i < n; i++) {
        const double a = x[i], b = y ? y[i] : 0.0;
        switch (fn) {
        case 0: out[i] = lm_sin(a); break; case 1: out[i] = lm_cos(a); break; case 2: out[i] = lm_atan2(a, b); break; case 3: out[i] = lm_acos(a); break;
        case 4: out[i] = lm_atanh(a); break; case 5: out[i] = lm_cosh(a); break; case 6: out[i] = lm_exp(a); break; case 7: out[i] = lm_log(a); break;
        default: out[i] = lm_pow(a, b); break;
        }
    }
}
const char* lumo_host_last_error(void) { return lumo_host::g_err.c_str(); }
}
