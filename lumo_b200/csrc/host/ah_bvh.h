// Builder of the order-free occlusion structure (LSEC_AH_NODES / LSEC_AH_PRIMS, csrc/common/scene_blob.h).
//
// `Scene::hit_light` (src/tracer/scene.rs:165-189) only asks whether ANY primitive is hit before the light, and that
// boolean does not depend on the order in which lumo's object BVH and kd-trees are walked.  So shadow rays
// do not have to replay the reference's unordered two-level traversal: this file builds a conventional 4-wide
// bounding-volume hierarchy over every primitive of the scene in world space — binned surface-area heuristic on a binary
// tree, then collapsed to four children per node — with f32 boxes rounded outwards.  The boxes only cull; every primitive
// they let through is tested with the reference's own f64 triangle / sphere test in its object's frame, and a candidate
// blocker is confirmed by the reference's own traversal of that one object (csrc/gpu/occlude.cuh).
#pragma once
#include <cstdlib>
#include "host_math.h"
#include "../common/scene_blob.h"
#include <algorithm>
#include <atomic>
#include <cmath>
#include <thread>
#include <vector>

namespace lumo_host {

struct AhBox { float lo[3], hi[3]; };
static inline float f32_down(double v) { float f = (float)v; if ((double)f > v) f = std::nextafterf(f, -INFINITY); return f; }
static inline float f32_up(double v) { float f = (float)v; if ((double)f < v) f = std::nextafterf(f, INFINITY); return f; }
// outward rounding plus two f32 ulps of slack: covers the last-bit differences between a world-space vertex and the
// object-frame arithmetic the exact test runs in
static inline AhBox ah_box(const Box& b) {
    AhBox r;
    const double lo[3] = {b.lo.x, b.lo.y, b.lo.z}, hi[3] = {b.hi.x, b.hi.y, b.hi.z};
    for (int k = 0; k < 3; k++) {
        const double pad = std::fmax(std::fabs(lo[k]), std::fabs(hi[k])) * (1.0 / 4194304.0) + 1e-30;
        r.lo[k] = f32_down(lo[k] - pad); r.hi[k] = f32_up(hi[k] + pad);
    }
    return r;
}
static inline AhBox ah_empty() { AhBox b; for (int k = 0; k < 3; k++) { b.lo[k] = INFINITY; b.hi[k] = -INFINITY; } return b; }
static inline void ah_grow(AhBox& a, const AhBox& b) { for (int k = 0; k < 3; k++) { a.lo[k] = std::fmin(a.lo[k], b.lo[k]); a.hi[k] = std::fmax(a.hi[k], b.hi[k]); } }
static inline double ah_area(const AhBox& b) {
    const double dx = (double)b.hi[0] - b.lo[0], dy = (double)b.hi[1] - b.lo[1], dz = (double)b.hi[2] - b.lo[2];
    if (!(dx >= 0.0 && dy >= 0.0 && dz >= 0.0)) return 0.0;
    return 2.0 * (dx * dy + dx * dz + dy * dz);
}

struct AhBuilder {
    // input: one box and one record per primitive
    std::vector<AhBox> boxes; std::vector<LumoAhPrim> prims;
    // binary tree
    struct N2 { AhBox box; uint32_t left, right, first, count; };   // count > 0: leaf over idx[first, first + count)
    std::vector<N2> n2; std::vector<uint32_t> idx;
    // output
    std::vector<LumoAhNode> nodes; std::vector<LumoAhPrim> out_prims;

    void add(const Box& world, uint32_t tri, uint32_t obj) { boxes.push_back(ah_box(world)); LumoAhPrim p = {tri, obj}; prims.push_back(p); }

    static constexpr int BINS = 16;
    static constexpr double C_TRAV = 1.0;
    double C_ISECT = 1.5;              // cost of one f64 primitive test in units of one four-box node step (LUMO_AH_CISECT overrides: experiments)
    uint32_t max_leaf = LUMO_AH_MAX_LEAF;
    AhBuilder() {
        if (const char* e = std::getenv("LUMO_AH_CISECT")) { const double v = std::atof(e); if (v > 0.0) C_ISECT = v; }
        if (const char* e = std::getenv("LUMO_AH_MAXLEAF")) { const int v = std::atoi(e); if (v >= 1 && v <= (int)LUMO_AH_MAX_LEAF) max_leaf = (uint32_t)v; }
    }

    // One node of the binary tree over idx[begin, end): its box, and — unless it becomes a leaf — the binned-SAH split.
    // Returns false for a leaf; else `mid` (idx[begin, mid) goes left).  `threads` > 1 spreads the two passes over the
    // primitives (bounds, then bins) over that many threads: boxes merge by min / max and counts by integer sums, so the
    // result does not depend on the number of threads.
    bool split(uint32_t begin, uint32_t end, AhBox& box_out, uint32_t& mid, unsigned threads) {
        const uint32_t cnt = end - begin;
        struct Part { AhBox bb; double clo[3], chi[3]; AhBox bin_box[3][BINS]; uint32_t bin_n[3][BINS]; };
        const unsigned T = (threads > 1 && cnt >= (1u << 16)) ? threads : 1u;
        std::vector<Part> parts(T);
        auto chunk = [&](unsigned t, uint32_t& b0, uint32_t& b1) { b0 = begin + (uint32_t)((uint64_t)cnt * t / T); b1 = begin + (uint32_t)((uint64_t)cnt * (t + 1) / T); };
        auto run = [&](auto&& fn) {
            if (T == 1) { fn(0u); return; }
            std::vector<std::thread> th; th.reserve(T);
            for (unsigned t = 0; t < T; t++) th.emplace_back(fn, t);
            for (auto& x : th) x.join();
        };
        run([&](unsigned t) {
            Part& P = parts[t]; P.bb = ah_empty(); for (int k = 0; k < 3; k++) { P.clo[k] = kInf; P.chi[k] = -kInf; }
            uint32_t b0, b1; chunk(t, b0, b1);
            for (uint32_t i = b0; i < b1; i++) {
                const AhBox& b = boxes[idx[i]]; ah_grow(P.bb, b);
                for (int k = 0; k < 3; k++) { const double c = 0.5 * ((double)b.lo[k] + b.hi[k]); P.clo[k] = std::fmin(P.clo[k], c); P.chi[k] = std::fmax(P.chi[k], c); }
            }
        });
        AhBox bb = parts[0].bb; double clo[3], chi[3];
        for (int k = 0; k < 3; k++) { clo[k] = parts[0].clo[k]; chi[k] = parts[0].chi[k]; }
        for (unsigned t = 1; t < T; t++) { ah_grow(bb, parts[t].bb); for (int k = 0; k < 3; k++) { clo[k] = std::fmin(clo[k], parts[t].clo[k]); chi[k] = std::fmax(chi[k], parts[t].chi[k]); } }
        box_out = bb;
        if (cnt <= 1) return false;
        // binned SAH over the three axes
        run([&](unsigned t) {
            Part& P = parts[t];
            for (int ax = 0; ax < 3; ax++) for (int b = 0; b < BINS; b++) { P.bin_box[ax][b] = ah_empty(); P.bin_n[ax][b] = 0; }
            uint32_t b0, b1; chunk(t, b0, b1);
            for (int ax = 0; ax < 3; ax++) {
                const double ext = chi[ax] - clo[ax];
                if (!(ext > 0.0)) continue;
                const double scale = (double)BINS / ext;
                for (uint32_t i = b0; i < b1; i++) {
                    const AhBox& b = boxes[idx[i]];
                    int k = (int)((0.5 * ((double)b.lo[ax] + b.hi[ax]) - clo[ax]) * scale); if (k >= BINS) k = BINS - 1; if (k < 0) k = 0;
                    ah_grow(P.bin_box[ax][k], b); P.bin_n[ax][k]++;
                }
            }
        });
        double best = kInf; int best_ax = -1, best_bin = -1;
        const double parent_area = ah_area(bb);
        for (int ax = 0; ax < 3; ax++) {
            const double ext = chi[ax] - clo[ax];
            if (!(ext > 0.0)) continue;
            AhBox bin_box[BINS]; uint32_t bin_n[BINS];
            for (int b = 0; b < BINS; b++) {
                bin_box[b] = parts[0].bin_box[ax][b]; bin_n[b] = parts[0].bin_n[ax][b];
                for (unsigned t = 1; t < T; t++) { ah_grow(bin_box[b], parts[t].bin_box[ax][b]); bin_n[b] += parts[t].bin_n[ax][b]; }
            }
            double right_area[BINS]; uint32_t right_n[BINS];
            AhBox acc = ah_empty(); uint32_t an = 0;
            for (int b = BINS - 1; b > 0; b--) { ah_grow(acc, bin_box[b]); an += bin_n[b]; right_area[b] = ah_area(acc); right_n[b] = an; }
            acc = ah_empty(); an = 0;
            for (int b = 0; b < BINS - 1; b++) {
                ah_grow(acc, bin_box[b]); an += bin_n[b];
                if (an == 0 || right_n[b + 1] == 0) continue;
                const double cost = C_TRAV + C_ISECT * (ah_area(acc) * an + right_area[b + 1] * right_n[b + 1]) / (parent_area > 0.0 ? parent_area : 1.0);
                if (cost < best) { best = cost; best_ax = ax; best_bin = b; }
            }
        }
        if (cnt <= max_leaf && !(best < C_ISECT * cnt)) return false;
        if (best_ax < 0) mid = begin + cnt / 2;             // all centres coincide: split the list in half
        else {
            const double ext = chi[best_ax] - clo[best_ax], scale = (double)BINS / ext;
            auto it = std::partition(idx.begin() + begin, idx.begin() + end, [&](uint32_t p) {
                const AhBox& b = boxes[p];
                int k = (int)((0.5 * ((double)b.lo[best_ax] + b.hi[best_ax]) - clo[best_ax]) * scale); if (k >= BINS) k = BINS - 1; if (k < 0) k = 0;
                return k <= best_bin; });
            mid = (uint32_t)(it - idx.begin());
            if (mid == begin || mid == end) mid = begin + cnt / 2;
        }
        return true;
    }

    // The subtree over idx[begin, end) into `out` (out[0] is its root, children by index into `out`), depth first.
    void build_subtree(uint32_t begin, uint32_t end, std::vector<N2>& out, unsigned threads) {
        struct Job { uint32_t node, begin, end; };
        std::vector<Job> jobs;
        N2 blank; blank.box = ah_empty(); blank.left = blank.right = LUMO_NONE; blank.first = 0; blank.count = 0;
        out.clear(); out.push_back(blank);
        jobs.push_back({0u, begin, end});
        while (!jobs.empty()) {
            const Job j = jobs.back(); jobs.pop_back();
            uint32_t mid = 0; AhBox bb;
            const bool inner = split(j.begin, j.end, bb, mid, threads);
            out[j.node].box = bb;
            if (!inner) { out[j.node].first = j.begin; out[j.node].count = j.end - j.begin; continue; }
            const uint32_t l = (uint32_t)out.size(); out.push_back(blank); out.push_back(blank);
            out[j.node].left = l; out[j.node].right = l + 1; out[j.node].count = 0;
            jobs.push_back({l, j.begin, mid}); jobs.push_back({l + 1, mid, j.end});
        }
    }

    // Top of the tree first (large nodes, their passes over the primitives spread over the threads) until the open ranges are
    // small enough to be many; then those subtrees in parallel, each over its own slice of idx; then the subtrees appended in
    // range order.  The tree — and with it the collapsed output — is the one a single thread builds.
    void build_binary() {
        const uint32_t n = (uint32_t)boxes.size();
        idx.resize(n); for (uint32_t i = 0; i < n; i++) idx[i] = i;
        n2.clear(); n2.reserve(2 * (size_t)n + 1);
        unsigned threads = std::thread::hardware_concurrency(); if (threads == 0) threads = 1; if (threads > 32) threads = 32;
        if (const char* e = std::getenv("LUMO_HOST_THREADS")) { const int v = std::atoi(e); if (v >= 1 && v <= 256) threads = (unsigned)v; }
        N2 blank; blank.box = ah_empty(); blank.left = blank.right = LUMO_NONE; blank.first = 0; blank.count = 0;
        n2.push_back(blank);
        if (n == 0) return;
        struct Job { uint32_t node, begin, end; };
        const uint32_t grain = std::max<uint32_t>(2048u, n / (threads * 8u));
        std::vector<Job> open, leaves_of_top;
        open.push_back({0u, 0u, n});
        while (!open.empty()) {
            const Job j = open.back(); open.pop_back();
            if (threads == 1 || j.end - j.begin <= grain) { leaves_of_top.push_back(j); continue; }
            uint32_t mid = 0; AhBox bb;
            const bool inner = split(j.begin, j.end, bb, mid, threads);
            n2[j.node].box = bb;
            if (!inner) { n2[j.node].first = j.begin; n2[j.node].count = j.end - j.begin; continue; }
            const uint32_t l = (uint32_t)n2.size(); n2.push_back(blank); n2.push_back(blank);
            n2[j.node].left = l; n2[j.node].right = l + 1; n2[j.node].count = 0;
            open.push_back({l, j.begin, mid}); open.push_back({l + 1, mid, j.end});
        }
        std::sort(leaves_of_top.begin(), leaves_of_top.end(), [](const Job& a, const Job& b) { return a.begin < b.begin; });
        std::vector<std::vector<N2>> sub(leaves_of_top.size());
        {
            std::atomic<size_t> next{0};
            auto worker = [&]() { for (;;) { const size_t k = next.fetch_add(1); if (k >= leaves_of_top.size()) break; build_subtree(leaves_of_top[k].begin, leaves_of_top[k].end, sub[k], 1u); } };
            const unsigned T = (unsigned)std::min<size_t>(threads, leaves_of_top.size());
            if (T <= 1) worker();
            else { std::vector<std::thread> th; th.reserve(T); for (unsigned t = 0; t < T; t++) th.emplace_back(worker); for (auto& x : th) x.join(); }
        }
        for (size_t k = 0; k < leaves_of_top.size(); k++) {
            const std::vector<N2>& L = sub[k];
            const uint32_t base = (uint32_t)n2.size();             // local node m >= 1 becomes base + m - 1; local 0 is the node the top phase left open
            auto reloc = [&](uint32_t m) { return m == LUMO_NONE ? LUMO_NONE : base + m - 1u; };
            N2 r = L[0]; r.left = reloc(r.left); r.right = reloc(r.right);
            n2[leaves_of_top[k].node] = r;
            for (size_t m = 1; m < L.size(); m++) { N2 c = L[m]; c.left = reloc(c.left); c.right = reloc(c.right); n2.push_back(c); }
        }
    }

    // Collapse: a wide node takes the two children of a binary node and keeps replacing the inner child of the largest
    // area by that child's own children until it has four (or only leaves are left).
    void collapse() {
        nodes.clear(); out_prims.clear(); out_prims.reserve(prims.size());
        if (n2.empty() || boxes.empty()) return;
        struct Job { uint32_t wide, bin; };
        std::vector<Job> jobs;
        nodes.emplace_back();
        jobs.push_back({0u, 0u});
        while (!jobs.empty()) {
            const Job j = jobs.back(); jobs.pop_back();
            uint32_t kids[4]; int nk = 0;
            if (n2[j.bin].count > 0) kids[nk++] = j.bin;            // a tree that is one leaf: the root node holds it as its only child
            else { kids[nk++] = n2[j.bin].left; kids[nk++] = n2[j.bin].right; }
            for (;;) {
                if (nk >= 4) break;
                int pick = -1; double pa = -1.0;
                for (int k = 0; k < nk; k++) if (n2[kids[k]].count == 0) { const double a = ah_area(n2[kids[k]].box); if (a > pa) { pa = a; pick = k; } }
                if (pick < 0) break;
                const uint32_t c = kids[pick];
                kids[pick] = n2[c].left; kids[nk++] = n2[c].right;
            }
            LumoAhNode W;
            for (int k = 0; k < 4; k++) { W.lo_x[k] = W.lo_y[k] = W.lo_z[k] = INFINITY; W.hi_x[k] = W.hi_y[k] = W.hi_z[k] = -INFINITY; W.child[k] = LUMO_NONE; W.pad[k] = 0; }
            for (int k = 0; k < nk; k++) {
                const N2& c = n2[kids[k]];
                W.lo_x[k] = c.box.lo[0]; W.lo_y[k] = c.box.lo[1]; W.lo_z[k] = c.box.lo[2]; W.hi_x[k] = c.box.hi[0]; W.hi_y[k] = c.box.hi[1]; W.hi_z[k] = c.box.hi[2];
                if (c.count > 0) {
                    // leaves of the binary tree hold at most LUMO_AH_MAX_LEAF primitives unless all their centres coincide
                    uint32_t first = (uint32_t)out_prims.size(), cnt = c.count;
                    for (uint32_t i = 0; i < c.count; i++) out_prims.push_back(prims[idx[c.first + i]]);
                    if (cnt > 16) cnt = 16;   // cannot happen: build_binary splits every list longer than LUMO_AH_MAX_LEAF
                    W.child[k] = LUMO_AH_LEAF | ((cnt - 1u) << 27) | first;
                } else {
                    const uint32_t w = (uint32_t)nodes.size(); nodes.emplace_back();
                    W.child[k] = w; jobs.push_back({w, kids[k]});
                }
            }
            nodes[j.wide] = W;
        }
    }
};

}  // namespace lumo_host
