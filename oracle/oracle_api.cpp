// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).  PARITY UNPINNED.
// Scene-program loader, the tile/batch renderer (src/renderer.rs, src/renderer/task.rs,
// src/samplers.rs) and a plain C interface for ctypes.  Nothing under lumo_b200/ links this.
#include "oracle_shading.hpp"
#include "oracle_bdpt.hpp"
#include <thread>
#include <atomic>
#include <mutex>
#include <map>
#include <string>
#include <tuple>
#include <functional>

namespace oracle {

thread_local Counters g_cnt;
static std::mutex g_total_mu;
static Counters g_total;
static void flush_counters() { std::lock_guard<std::mutex> l(g_total_mu); g_total.add(g_cnt); g_cnt = Counters(); }

// ---- scene program (format: lumo_b200/program.py) ---------------------------------------------
struct Reader {
    const uint8_t* p; size_t n, off = 0;
    int64_t i64() { int64_t v; std::memcpy(&v, p + off, 8); off += 8; return v; }
    double f64() { double v; std::memcpy(&v, p + off, 8); off += 8; return v; }
};
enum { TAG_MATERIAL = 1, TAG_MESH = 2, TAG_OBJECT = 3, TAG_ENVMAP = 4, TAG_CAMERA = 5, TAG_TEXTURE = 6 };
enum { OBJ_KDMESH = 0, OBJ_RECT = 1, OBJ_SPHERE = 2, OBJ_LOOSE_TRIS = 3 };
enum { OP_UNIT = 0, OP_ORIGIN = 1, OP_SETX = 2, OP_SETY = 3, OP_SETZ = 4, OP_TRANSLATE = 5, OP_SCALE = 6, OP_ROTX = 7, OP_ROTY = 8, OP_ROTZ = 9 };

struct MeshSrc { std::vector<int64_t> face_off, vidx, nidx, tidx; };

struct Loaded {
    Scene scene;
    Camera camera;
    std::vector<MeshSrc> mesh_src;
    std::vector<std::unique_ptr<Texture>> textures;
    std::map<std::tuple<int64_t, int64_t, int64_t, int64_t>, std::shared_ptr<KdTree>> kd_cache;
    std::string err;
};

static Spectrum read_spec(Reader& r) { Spectrum s; s.c0 = (float)r.f64(); s.c1 = (float)r.f64(); s.c2 = (float)r.f64(); s.scale = (float)r.f64(); return s; }

static DenseSpectrum eta_table(int64_t kind, double c) {
    switch (kind) {   // material.rs:37-45 (1.5 -> glass_eta, 2.5 -> diamond_eta when transparent), :136-140 mirror
    case 1: return DenseSpectrum::from(spectra::glass_eta);
    case 2: return DenseSpectrum::from(spectra::diamond_eta);
    case 3: return DenseSpectrum::from(spectra::mirror_eta);
    case 4: return DenseSpectrum::from(spectra::mirror_k);
    default: return DenseSpectrum::constant(c);
    }
}

static std::vector<Triangle> triangles_from_faces(const TriangleMesh* tm, const MeshSrc& ms, int64_t f0, int64_t f1, const Material* mat) {  // triangle_mesh.rs:58-91
    std::vector<Triangle> tris;
    for (int64_t f = f0; f < f1; f++) {
        int64_t b = ms.face_off[f], e = ms.face_off[f + 1];
        for (int64_t i = 1; i + 1 < e - b; i++) {
            int64_t ia = b, ib = b + i, ic = b + i + 1;
            Triangle t; t.mesh = tm; t.mat = mat;
            t.v[0] = (uint32_t)ms.vidx[ia]; t.v[1] = (uint32_t)ms.vidx[ib]; t.v[2] = (uint32_t)ms.vidx[ic];
            if (degenerate_triangle(t.a(), t.b(), t.c())) continue;
            if (!ms.nidx.empty() && ms.nidx[ia] >= 0 && ms.nidx[ib] >= 0 && ms.nidx[ic] >= 0) { t.has_n = true; t.n[0] = (uint32_t)ms.nidx[ia]; t.n[1] = (uint32_t)ms.nidx[ib]; t.n[2] = (uint32_t)ms.nidx[ic]; }
            if (!ms.tidx.empty() && ms.tidx[ia] >= 0 && ms.tidx[ib] >= 0 && ms.tidx[ic] >= 0) { t.has_t = true; t.tx[0] = (uint32_t)ms.tidx[ia]; t.tx[1] = (uint32_t)ms.tidx[ib]; t.tx[2] = (uint32_t)ms.tidx[ic]; }
            tris.push_back(t);
        }
    }
    return tris;
}

static Loaded* load_program(const uint8_t* data, size_t len) {
    auto* L = new Loaded();
    if (len < 16 || std::memcmp(data, "LUMOPRG1", 8) != 0) { L->err = "bad magic"; return L; }
    Reader hd{data, len, 8};
    uint32_t version, nrec; std::memcpy(&version, data + 8, 4); std::memcpy(&nrec, data + 12, 4);
    size_t off = 16;
    Scene& sc = L->scene;
    bool have_camera = false;
    for (uint32_t rec = 0; rec < nrec; rec++) {
        uint32_t tag; uint64_t nbytes;
        std::memcpy(&tag, data + off, 4); std::memcpy(&nbytes, data + off + 8, 8);
        Reader r{data, len, off + 16};
        off += 16 + nbytes;
        if (tag == TAG_MATERIAL) {
            auto m = std::make_unique<Material>();
            m->kind = (int)r.i64();
            Float rough = r.f64();
            m->roughness = Vec2(fmax_(rough, 1e-5), fmax_(rough, 1e-5));   // microfacet.rs:30-36
            int64_t ek = r.i64(); double ec = r.f64(); int64_t kk = r.i64(); double kc = r.f64();
            m->eta = eta_table(ek, ec); m->k = eta_table(kk, kc);
            m->eta_const = m->eta.is_constant();
            m->kd = read_spec(r); m->ks = read_spec(r); m->tf = read_spec(r); m->ke = read_spec(r);
            m->spec = m->kd;
            m->illum = illuminant_table((int)r.i64());
            m->scale = r.f64(); m->two_sided = r.i64() != 0;
            if (nbytes >= 30 * 8) {
                const Texture** slots[5] = {&m->kd_tex, &m->ks_tex, &m->tf_tex, &m->ke_tex, &m->bump_tex};
                for (auto slot : slots) { int64_t t = r.i64(); *slot = t < 0 ? nullptr : L->textures.at((size_t)t).get(); }
            }
            sc.materials.push_back(std::move(m));
        } else if (tag == TAG_TEXTURE) {
            auto t = std::make_unique<Texture>();
            t->kind = (int)r.i64(); t->spec = read_spec(r);
            int64_t a = r.i64(), b = r.i64(); t->scale = r.f64();
            int64_t seed = r.i64(), w = r.i64(), h = r.i64();
            if (t->kind == TEX_CHECKER) { t->t1 = L->textures.at((size_t)a).get(); t->t2 = L->textures.at((size_t)b).get(); }
            else if (t->kind == TEX_MARBLE) t->pn = std::make_unique<Perlin>((uint64_t)seed);
            else if (t->kind == TEX_IMAGE) {
                t->img.width = (uint32_t)w; t->img.height = (uint32_t)h;
                for (int64_t i = 0; i < w * h; i++) t->img.buffer.push_back(read_spec(r));
            } else if (t->kind == TEX_BUMP) {
                t->bump.width = (uint32_t)w; t->bump.height = (uint32_t)h;
                for (int64_t i = 0; i < w * h; i++) { Float x = r.f64(), y = r.f64(), z = r.f64(); t->bump.buffer.push_back(Vec3(x, y, z)); }
            }
            L->textures.push_back(std::move(t));
        } else if (tag == TAG_MESH) {
            int64_t nv = r.i64(), nn = r.i64(), nt = r.i64(), nf = r.i64(), nc = r.i64(), has_n = r.i64(), has_t = r.i64();
            auto tm = std::make_unique<TriangleMesh>();
            MeshSrc ms;
            for (int64_t i = 0; i < nv; i++) { Float x = r.f64(), y = r.f64(), z = r.f64(); tm->vertices.push_back(Vec3(x, y, z)); }
            for (int64_t i = 0; i < nn; i++) { Float x = r.f64(), y = r.f64(), z = r.f64(); tm->normals.push_back(Vec3(x, y, z)); }
            for (int64_t i = 0; i < nt; i++) { Float x = r.f64(), y = r.f64(); tm->uvs.push_back(Vec2(x, y)); }
            ms.face_off.resize(nf + 1); for (auto& v : ms.face_off) v = r.i64();
            ms.vidx.resize(nc); for (auto& v : ms.vidx) v = r.i64();
            if (has_n) { ms.nidx.resize(nc); for (auto& v : ms.nidx) v = r.i64(); }
            if (has_t) { ms.tidx.resize(nc); for (auto& v : ms.tidx) v = r.i64(); }
            sc.meshes.push_back(std::move(tm)); L->mesh_src.push_back(std::move(ms));
        } else if (tag == TAG_OBJECT) {
            int64_t kind = r.i64(), is_light = r.i64(), mat = r.i64(), mesh = r.i64(), f0 = r.i64(), f1 = r.i64();
            double prm[9]; for (int i = 0; i < 9; i++) prm[i] = r.f64();
            int64_t inst_mat = r.i64(), n_ops = r.i64();
            const Material* m = mat >= 0 ? sc.materials[mat].get() : nullptr;
            std::vector<std::shared_ptr<Object>> made;
            if (kind == OBJ_KDMESH) {
                auto key = std::make_tuple(mesh, f0, f1, mat);
                auto it = L->kd_cache.find(key);
                std::shared_ptr<KdTree> kd;
                if (it != L->kd_cache.end()) kd = it->second;
                else {
                    kd = std::make_shared<KdTree>(triangles_from_faces(sc.meshes[mesh].get(), L->mesh_src[mesh], f0, f1, m));
                    L->kd_cache[key] = kd;
                }
                made.push_back(kd);
            } else if (kind == OBJ_RECT) {
                made.push_back(std::make_shared<Rectangle>(Vec3(prm[0], prm[1], prm[2]), Vec3(prm[3], prm[4], prm[5]), Vec3(prm[6], prm[7], prm[8]), m));
            } else if (kind == OBJ_SPHERE) {
                made.push_back(std::make_shared<Sphere>(prm[0], m));
            } else if (kind == OBJ_LOOSE_TRIS) {
                for (auto& t : triangles_from_faces(sc.meshes[mesh].get(), L->mesh_src[mesh], f0, f1, m)) made.push_back(std::make_shared<Triangle>(t));
            }
            struct Op { int64_t op; double x, y, z; };
            std::vector<Op> ops(n_ops);
            for (auto& o : ops) { o.op = r.i64(); o.x = r.f64(); o.y = r.f64(); o.z = r.f64(); }
            for (auto& base : made) {
                std::shared_ptr<Object> obj = base;
                if (n_ops > 0) {
                    auto inst = std::make_shared<Instance>(base);
                    if (inst_mat >= 0) inst->mat = sc.materials[inst_mat].get();
                    for (auto& o : ops) {
                        switch (o.op) {
                        case OP_UNIT: {                                                        // kdtree.rs:93-99
                            AABB bb = base->bounding_box();
                            Float s = 1.0 / (bb.ax_max - bb.ax_min).max_element();
                            inst->apply(Transform::scale(s, s, s)); break;
                        }
                        case OP_ORIGIN: inst->to_origin(); break;
                        case OP_SETX: inst->set_axis(0, o.x); break;
                        case OP_SETY: inst->set_axis(1, o.x); break;
                        case OP_SETZ: inst->set_axis(2, o.x); break;
                        case OP_TRANSLATE: inst->apply(Transform::translation(o.x, o.y, o.z)); break;
                        case OP_SCALE: inst->apply(Transform::scale(o.x, o.y, o.z)); break;
                        case OP_ROTX: inst->apply(Transform::rotate_x(o.x)); break;
                        case OP_ROTY: inst->apply(Transform::rotate_y(o.x)); break;
                        case OP_ROTZ: inst->apply(Transform::rotate_z(o.x)); break;
                        }
                    }
                    obj = inst;
                }
                if (is_light) sc.lights.add(obj); else sc.objects.add(obj);
            }
        } else if (tag == TAG_ENVMAP) {
            auto m = std::make_unique<Material>();
            m->kind = M_LIGHT; m->ke = read_spec(r); m->scale = r.f64(); m->illum = spectra::D65; m->two_sided = true;  // scene.rs:73-77
            if (nbytes >= 6 * 8) { int64_t t = r.i64(); if (t >= 0) m->ke_tex = L->textures.at((size_t)t).get(); }
            sc.env = m.get();
            sc.materials.push_back(std::move(m));
        } else if (tag == TAG_CAMERA) {
            Vec3 o, t, u;
            o.x = r.f64(); o.y = r.f64(); o.z = r.f64(); t.x = r.f64(); t.y = r.f64(); t.z = r.f64(); u.x = r.f64(); u.y = r.f64(); u.z = r.f64();
            Float zoom = r.f64(), lens = r.f64(), focal = r.f64(), vfov = r.f64();
            int64_t rx = r.i64(), ry = r.i64(), ctype = r.i64(), fk = r.i64();
            Float fr = r.f64(), fp = r.f64();
            int64_t cs = r.i64(), il = r.i64();
            PixelFilter pf; pf.kind = (int)fk; pf.r = fr; pf.p = fp;
            L->camera = Camera::build(o, t, u, zoom, lens, focal, vfov, (uint64_t)rx, (uint64_t)ry, ctype == 1, pf, (int)cs, (int)il);
            have_camera = true;
        }
    }
    if (!have_camera) L->camera = Camera::build(Vec3(), Vec3(0, 0, -1), Vec3(0, 1, 0), 1.0, 0, 0, 90.0, 1024, 768, false, PixelFilter(), 1, 2);
    if (sc.objects.objects.empty()) { L->err = "scene has no objects"; return L; }
    if (sc.lights.objects.empty() && !sc.env) { L->err = "scene has no lights"; return L; }   // renderer.rs:42
    sc.build();
    return L;
}

// ---- samplers (src/samplers.rs) ----------------------------------------------------------------
static const uint64_t SAMPLES_INCREMENT = 256, TILE_SIZE = 16;   // renderer.rs:15-17
static std::vector<size_t> gen_perm(Rng& rng, size_t n) {                                       // rng.rs:104-116
    std::vector<size_t> perm(n);
    for (size_t i = 0; i < n; i++) perm[i] = i;
    for (size_t i = 0; i + 1 < n; i++) {
        size_t rnd = (size_t)rng.gen_u64();
        size_t j = i + (rnd % (n - i));
        std::swap(perm[i], perm[j]);
    }
    return perm;
}
// SobolSampler (samplers.rs:193-247, samplers/sobol_seq.rs:1-39): two Gray-code Sobol dimensions of degree 10, direction
// numbers m << (64 - i - 1); a batch starts from the state after `batch * 256` steps; the pixel's sampler seed is XORed in.
static const uint64_t SOBOL_M1[10] = {1, 1, 7, 15, 5, 19, 69, 51, 121, 695};     // dim 119 (sobol_seq.rs:7)
static const uint64_t SOBOL_M2[10] = {1, 1, 7, 7, 7, 53, 57, 229, 473, 533};     // dim 103 (sobol_seq.rs:9)
static inline uint64_t sobol_v(const uint64_t* m, int i) { return m[i] << (64 - i - 1); }
struct Sampler {   // Uniform / Jittered / MultiJittered / Sobol (samplers.rs:54-247)
    int kind; uint64_t state, end, dim; Vec2 scale0, scale1; std::vector<size_t> px, py; Rng rng;
    uint64_t sob_seed = 0, sob0 = 0, sob1 = 0;
    Sampler(int kind_, uint64_t batch, uint64_t samples, uint64_t seed) : kind(kind_), rng(Rng::xorshift(seed)) {
        uint64_t s0 = batch * SAMPLES_INCREMENT, s1 = std::min((batch + 1) * SAMPLES_INCREMENT, samples);
        state = s0; end = s1;
        if (kind == 3) {                                   // BATCH_STATES[batch]: the sequence advanced s0 times (sobol_seq.rs:13-31)
            sob_seed = seed;
            for (uint64_t k = 1; k <= s0; k++) { const int tz = __builtin_ctzll(k); sob0 ^= sobol_v(SOBOL_M1, tz); sob1 ^= sobol_v(SOBOL_M2, tz); }
            return;
        }
        if (kind == 0) { state = 0; end = s1 - s0; return; }
        dim = sat_u64(std::ceil(std::sqrt((Float)samples)));
        scale0 = Vec2(1.0 / (Float)dim, (Float)dim / (Float)samples);
        if (kind == 2) { px = gen_perm(rng, dim); py = gen_perm(rng, dim); scale1 = scale0 / (Float)dim; }
    }
    bool next(Vec2& out) {
        if (state == end) return false;
        if (kind == 3) {                                   // samplers.rs:218-247
            state++;
            const int tz = __builtin_ctzll(state);
            sob0 ^= sobol_v(SOBOL_M1, tz); sob1 ^= sobol_v(SOBOL_M2, tz);
            out = Vec2((Float)(sob0 ^ sob_seed) * 5.421010862427522170037e-20, (Float)(sob1 ^ sob_seed) * 5.421010862427522170037e-20);
            return true;
        }
        if (kind == 0) { state++; out = rng.gen_vec2(); return true; }
        uint64_t x0 = state % dim, y0 = state / dim;
        Vec2 off0 = scale0 * Vec2((Float)x0, (Float)y0);
        if (kind == 1) { state++; out = scale0 * rng.gen_vec2() + off0; return true; }
        size_t x1 = px[y0], y1 = py[x0];
        Vec2 off1 = scale1 * Vec2((Float)x1, (Float)y1);
        Vec2 rs = scale1 * rng.gen_vec2();
        state++;
        out = off0 + off1 + rs;
        return true;
    }
};

// Correlated multi-jitter with a hashed permutation (Kensler 2013) — NOT in the reference; this is
// the stateless stand-in the GPU path uses for MultiJitteredSampler's Fisher-Yates tables.
static inline uint32_t cmj_permute(uint32_t i, uint32_t l, uint32_t p) {
    uint32_t w = l - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p; i *= 0xe170893d; i ^= p >> 16; i ^= (i & w) >> 4; i ^= p >> 8; i *= 0x0929eb3f;
        i ^= p >> 23; i ^= (i & w) >> 1; i *= 1 | p >> 27; i *= 0x6935fa69; i ^= (i & w) >> 11;
        i *= 0x74dcb303; i ^= (i & w) >> 2; i *= 0x9e501cc3; i ^= (i & w) >> 2; i *= 0xc860a3df;
        i &= w; i ^= i >> 5;
    } while (i >= l);
    return (i + p) % l;
}

struct RenderParams {       // mirrors include/lumo_gpu.h lumo_render_params + two oracle-only fields
    int32_t integrator, sampler, tone_map, rng_mode;
    double tone_map_arg, rr_delta;
    uint64_t seed;
    uint32_t spp_begin, spp_end, total_spp;
    int32_t threads;
    int32_t tile_step, pad;  // reference schedule only: render every tile_step-th 16x16 tile (0 / 1 = all) — bench.py's bounded CPU sample of a full-size workload at the workload's own spp
};

static std::vector<FilmSample> integrate(const Loaded& L, int integrator, Rng& rng, Float delta, Vec2 raster_xy) {  // integrator.rs:45-69
    Ray r = L.camera.generate_ray(raster_xy, rng.gen_vec2());
    Lambda lam = Lambda::sample(rng.gen_float());
    if (integrator == 0) return {path_trace(L.scene, r, rng, lam, delta, raster_xy)};
    if (integrator == 1) return {direct_light(L.scene, r, rng, lam, raster_xy)};
    return bdpt_integrate(L.scene, L.camera, r, rng, lam, delta, raster_xy);
}

struct FilmAccum { uint64_t w, h; std::vector<double> pixels, splats; uint64_t cam = 0, cost = 0; };

static void add_tile(FilmAccum& f, const FilmTile& t) {                                         // film.rs:155-171
    for (uint64_t y = 0; y < t.px_max_y - t.px_min_y; y++) for (uint64_t x = 0; x < t.px_max_x - t.px_min_x; x++) {
        const Pixel& p = t.pixels[x + y * t.width];
        size_t idx = (x + t.px_min_x) + (y + t.px_min_y) * f.w;
        f.pixels[4 * idx + 0] += p.color.x; f.pixels[4 * idx + 1] += p.color.y; f.pixels[4 * idx + 2] += p.color.z; f.pixels[4 * idx + 3] += p.w;
    }
    for (auto& s : t.splats) { size_t idx = s.x + f.w * s.y; f.splats[3 * idx] += s.color.x; f.splats[3 * idx + 1] += s.color.y; f.splats[3 * idx + 2] += s.color.z; }
}

// Reference schedule: renderer.rs:159-244 (tasks = batch x tile, one seed each) and
// task.rs:25-81 (per-tile sequential RNG, ring-buffer RR threshold).
static void render_reference(const Loaded& L, const RenderParams& P, FilmAccum& film) {
    const Camera& cam = L.camera;
    ColorSpace cs = ColorSpace::get(cam.color_space);
    Mat3 wb = cs.wb_matrix(illuminant_table(cam.illuminant));
    struct Task { uint64_t x0, y0, x1, y1, batch, samples, seed; };
    std::vector<Task> tasks;
    Rng master = Rng::xorshift(P.seed);
    uint64_t tiles_x = (cam.res_x + TILE_SIZE - 1) / TILE_SIZE, tiles_y = (cam.res_y + TILE_SIZE - 1) / TILE_SIZE;
    uint64_t taken = 0, total = P.total_spp;
    while (taken < total) {
        uint64_t prev = taken, batch = taken / SAMPLES_INCREMENT;
        taken = std::min(taken + SAMPLES_INCREMENT, total);
        for (uint64_t y = 0; y < tiles_y; y++) for (uint64_t x = 0; x < tiles_x; x++) {
            uint64_t x0 = x * TILE_SIZE, y0 = y * TILE_SIZE;
            const uint64_t seed = master.gen_u64();                 // every tile draws its seed, rendered or not
            if (P.tile_step > 1 && (x + y * tiles_x) % (uint64_t)P.tile_step != 0) continue;
            tasks.push_back({x0, y0, std::min(x0 + TILE_SIZE, cam.res_x), std::min(y0 + TILE_SIZE, cam.res_y), batch, taken - prev, seed});
        }
    }
    std::vector<std::unique_ptr<FilmTile>> done(tasks.size());
    std::vector<uint64_t> costs(tasks.size(), 0);
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        while (true) {
            size_t ti = next.fetch_add(1);
            if (ti >= tasks.size()) break;
            const Task& task = tasks[ti];
            auto tile = std::make_unique<FilmTile>(task.x0, task.y0, task.x1, task.y1, cam.res_x, cam.res_y, &cs, wb, &cam.filter);
            Rng rng = Rng::xorshift(task.seed);
            std::vector<size_t> ns(SAMPLES_INCREMENT, 0); std::vector<Float> fs(SAMPLES_INCREMENT, 0.0);
            size_t ptr = 0; uint64_t num_rays = 0;
            for (uint64_t y = task.y0; y < task.y1; y++) for (uint64_t x = task.x0; x < task.x1; x++) {
                Vec2 xy((Float)x, (Float)y);
                Sampler sampler(P.sampler, task.batch, total, rng.gen_u64());
                Vec2 rs;
                while (sampler.next(rs)) {
                    Vec2 raster_xy = xy + rs;
                    Float f = 0.0, f2 = 0.0;
                    for (uint64_t i = 0; i < task.samples; i++) f += fs[i];
                    for (uint64_t i = 0; i < task.samples; i++) f2 += fs[i] * fs[i];
                    Float var = f2 - f * f / (Float)task.samples;
                    Float delta;
                    if (var <= 0.0) delta = 1e-5;
                    else { size_t cost = 0; for (uint64_t i = 0; i < task.samples; i++) cost += ns[i]; delta = std::sqrt(var / (Float)cost); }
                    if (P.rr_delta > 0.0) delta = P.rr_delta;
                    auto samples = integrate(L, P.integrator, rng, delta, raster_xy);
                    FilmSample& main = samples.back();
                    num_rays += main.cost;
                    ns[ptr] = main.cost; fs[ptr] = color_luminance(main.color, main.lambda);
                    ptr = (ptr + 1) % task.samples;
                    for (auto& s : samples) { s.color = tone_map(P.tone_map, P.tone_map_arg, s); tile->add_sample(s); }
                }
            }
            costs[ti] = num_rays;
            done[ti] = std::move(tile);
        }
        flush_counters();
    };
    int nt = std::max(1, P.threads);
    std::vector<std::thread> th;
    for (int i = 0; i < nt; i++) th.emplace_back(worker);
    for (auto& t : th) t.join();
    film.cam = 0;
    for (size_t i = 0; i < tasks.size(); i++) { add_tile(film, *done[i]); film.cost += costs[i]; done[i].reset(); film.cam += tasks[i].samples * (tasks[i].x1 - tasks[i].x0) * (tasks[i].y1 - tasks[i].y0); }
}

// GPU schedule (what lumo_gpu_render does): every (pixel, sample) owns a Philox
// stream, raster jitter = hashed CMJ, per-tile RR threshold from two pilot rounds.  Shading and
// traversal code is the same restated reference code as above.
static Vec2 counter_jitter(const RenderParams& P, uint32_t pixel, uint32_t s, Rng& rng) {
    if (P.sampler == 0) return rng.gen_vec2();
    if (P.sampler == 3) {   // Sobol point s + 1 in closed form: the Gray-code recurrence unrolled, XOR of the direction numbers of the set bits of gray(n)
        Rng keyr = Rng::counter(P.seed, pixel, 0xFFFFFFFFu, 1);
        const uint64_t k = keyr.gen_u64();
        const uint64_t n = (uint64_t)s + 1, g = n ^ (n >> 1);
        uint64_t a = 0, b = 0;
        for (int i = 0; i < 10; i++) if ((g >> i) & 1) { a ^= sobol_v(SOBOL_M1, i); b ^= sobol_v(SOBOL_M2, i); }
        return Vec2((Float)(a ^ k) * 5.421010862427522170037e-20, (Float)(b ^ k) * 5.421010862427522170037e-20);
    }
    uint64_t total = P.total_spp;
    uint64_t dim = sat_u64(std::ceil(std::sqrt((Float)total)));
    Vec2 scale0(1.0 / (Float)dim, (Float)dim / (Float)total);
    uint64_t x0 = s % dim, y0 = s / dim;
    Vec2 off0 = scale0 * Vec2((Float)x0, (Float)y0);
    if (P.sampler == 1) return scale0 * rng.gen_vec2() + off0;
    Rng keyr = Rng::counter(P.seed, pixel, 0xFFFFFFFFu, 1);
    uint64_t k = keyr.gen_u64();
    uint32_t kx = (uint32_t)k, ky = (uint32_t)(k >> 32);
    Vec2 scale1 = scale0 / (Float)dim;
    uint32_t x1 = cmj_permute((uint32_t)y0, (uint32_t)dim, kx), y1 = cmj_permute((uint32_t)x0, (uint32_t)dim, ky);
    Vec2 off1 = scale1 * Vec2((Float)x1, (Float)y1);
    Vec2 rs = scale1 * rng.gen_vec2();
    return off0 + off1 + rs;
}
static const uint32_t PILOT_N = 64;
static void render_counter(const Loaded& L, const RenderParams& P, FilmAccum& film, std::vector<double>* deltas_out) {
    const Camera& cam = L.camera;
    ColorSpace cs = ColorSpace::get(cam.color_space);
    Mat3 wb = cs.wb_matrix(illuminant_table(cam.illuminant));
    uint64_t tiles_x = (cam.res_x + TILE_SIZE - 1) / TILE_SIZE, tiles_y = (cam.res_y + TILE_SIZE - 1) / TILE_SIZE;
    size_t ntiles = tiles_x * tiles_y;
    std::vector<double> delta(ntiles, P.rr_delta > 0.0 ? P.rr_delta : 1e-5);
    std::vector<std::unique_ptr<FilmTile>> done(ntiles);
    std::vector<uint64_t> costs(ntiles, 0);
    int nt = std::max(1, P.threads);
    auto run = [&](std::function<void(size_t)> fn) {
        std::atomic<size_t> next(0);
        std::vector<std::thread> th;
        for (int i = 0; i < nt; i++) th.emplace_back([&]() { while (true) { size_t t = next.fetch_add(1); if (t >= ntiles) break; fn(t); } flush_counters(); });
        for (auto& t : th) t.join();
    };
    if (P.rr_delta <= 0.0 && P.integrator != 1) {
        for (uint32_t round = 0; round < 2; round++) {
            std::vector<double> nd(ntiles);
            run([&](size_t t) {
                uint64_t x0 = (t % tiles_x) * TILE_SIZE, y0 = (t / tiles_x) * TILE_SIZE;
                Float f = 0.0, f2 = 0.0; uint64_t cost = 0;
                for (uint32_t k = 0; k < PILOT_N; k++) {
                    uint64_t x = std::min(x0 + 2 * (k % 8), cam.res_x - 1), y = std::min(y0 + 2 * (k / 8), cam.res_y - 1);
                    uint32_t pixel = (uint32_t)(x + y * cam.res_x);
                    Rng rng = Rng::counter(P.seed, pixel, 0xFFFFFF00u + round, 0);
                    Vec2 raster_xy = Vec2((Float)x, (Float)y) + rng.gen_vec2();
                    auto samples = integrate(L, P.integrator, rng, delta[t], raster_xy);
                    FilmSample& main = samples.back();
                    Float lum = color_luminance(main.color, main.lambda);
                    f += lum; f2 += lum * lum; cost += main.cost;
                }
                Float var = f2 - f * f / (Float)PILOT_N;
                nd[t] = var <= 0.0 ? 1e-5 : std::sqrt(var / (Float)cost);
            });
            delta = nd;
        }
        { std::lock_guard<std::mutex> l(g_total_mu); g_total = Counters(); }   // reported ray counts cover the main pass only
    }
    if (deltas_out) *deltas_out = delta;
    run([&](size_t t) {
        uint64_t x0 = (t % tiles_x) * TILE_SIZE, y0 = (t / tiles_x) * TILE_SIZE;
        uint64_t x1 = std::min(x0 + TILE_SIZE, cam.res_x), y1 = std::min(y0 + TILE_SIZE, cam.res_y);
        auto tile = std::make_unique<FilmTile>(x0, y0, x1, y1, cam.res_x, cam.res_y, &cs, wb, &cam.filter);
        uint64_t num_rays = 0;
        for (uint64_t y = y0; y < y1; y++) for (uint64_t x = x0; x < x1; x++) {
            uint32_t pixel = (uint32_t)(x + y * cam.res_x);
            for (uint32_t s = P.spp_begin; s < P.spp_end; s++) {
                Rng rng = Rng::counter(P.seed, pixel, s, 0);
                { const char* e = std::getenv("ORACLE_DEBUG_PIXEL"); g_dbg = e && (uint32_t)std::atoll(e) == pixel; }
                Vec2 raster_xy = Vec2((Float)x, (Float)y) + counter_jitter(P, pixel, s, rng);
                auto samples = integrate(L, P.integrator, rng, delta[t], raster_xy);
                num_rays += samples.back().cost;
                for (auto& sm : samples) { sm.color = tone_map(P.tone_map, P.tone_map_arg, sm); tile->add_sample(sm); }
            }
        }
        costs[t] = num_rays; done[t] = std::move(tile);
    });
    for (size_t i = 0; i < ntiles; i++) { add_tile(film, *done[i]); film.cost += costs[i]; }
    film.cam = (uint64_t)(P.spp_end - P.spp_begin) * cam.res_x * cam.res_y;
}

}  // namespace oracle

using namespace oracle;

template <class F> static void parallel_for(uint64_t n, int threads, F fn) {
    int nt = std::max(1, threads);
    std::vector<std::thread> th;
    uint64_t chunk = (n + nt - 1) / nt;
    for (int i = 0; i < nt; i++) {
        uint64_t b = i * chunk, e = std::min(n, b + chunk);
        if (b >= e) break;
        th.emplace_back([=]() { fn(b, e); flush_counters(); });
    }
    for (auto& t : th) t.join();
}

extern "C" {

void* oracle_scene_create(const void* program, uint64_t len) { return load_program((const uint8_t*)program, (size_t)len); }
const char* oracle_scene_error(void* h) { auto* L = (Loaded*)h; return L->err.empty() ? nullptr : L->err.c_str(); }
void oracle_scene_destroy(void* h) { delete (Loaded*)h; }

void oracle_scene_info(void* h, uint64_t* out) {   // n_objects, n_lights, n_shadow_rays, res_x, res_y
    auto* L = (Loaded*)h;
    out[0] = L->scene.objects.objects.size(); out[1] = L->scene.lights.objects.size(); out[2] = L->scene.num_shadow_rays();
    out[3] = L->camera.res_x; out[4] = L->camera.res_y;
}
void oracle_scene_bounds(void* h, double* out) {
    auto* L = (Loaded*)h;
    out[0] = L->scene.bounds.ax_min.x; out[1] = L->scene.bounds.ax_min.y; out[2] = L->scene.bounds.ax_min.z;
    out[3] = L->scene.bounds.ax_max.x; out[4] = L->scene.bounds.ax_max.y; out[5] = L->scene.bounds.ax_max.z;
}

static Ray mk(const double* o, const double* d, uint64_t i) { return Ray::raw(Vec3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), Vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2])); }

// Scene::hit on a caller-supplied batch (directions are used as given: pass normalised ones).
void oracle_trace_closest(void* h, const double* o, const double* d, uint64_t n, uint32_t* obj, uint32_t* tri, double* t, double* bary, int threads) {
    auto* L = (Loaded*)h;
    parallel_for(n, threads, [=](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; i++) {
            Hit hit;
            if (L->scene.hit(mk(o, d, i), hit)) { obj[i] = (uint32_t)hit.obj; tri[i] = (uint32_t)hit.tri; t[i] = hit.t; bary[2 * i] = hit.bary.x; bary[2 * i + 1] = hit.bary.y; }
            else { obj[i] = 0xFFFFFFFFu; tri[i] = 0xFFFFFFFFu; t[i] = INF; bary[2 * i] = 0; bary[2 * i + 1] = 0; }
        }
    });
}
// full Hit record for shading parity: p, ng, ns, uv, fp_error, backface, material
void oracle_trace_closest_full(void* h, const double* o, const double* d, uint64_t n, double* out /* 16 per ray */, int threads) {
    auto* L = (Loaded*)h;
    parallel_for(n, threads, [=](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; i++) {
            Hit hit; double* q = out + 16 * i;
            if (L->scene.hit(mk(o, d, i), hit)) {
                q[0] = hit.t; q[1] = hit.p.x; q[2] = hit.p.y; q[3] = hit.p.z; q[4] = hit.ng.x; q[5] = hit.ng.y; q[6] = hit.ng.z;
                q[7] = hit.ns.x; q[8] = hit.ns.y; q[9] = hit.ns.z; q[10] = hit.uv.x; q[11] = hit.uv.y;
                q[12] = hit.fp_error.x; q[13] = hit.fp_error.y; q[14] = hit.fp_error.z; q[15] = hit.backface ? 1.0 : 0.0;
            } else for (int k = 0; k < 16; k++) q[k] = k == 0 ? INF : 0.0;
        }
    });
}
// the two tests of Scene::hit_light (scene.rs:180-186) with a caller-supplied t_max
void oracle_trace_any(void* h, const double* o, const double* d, const double* t_max, uint64_t n, uint8_t* occluded, int threads) {
    auto* L = (Loaded*)h;
    parallel_for(n, threads, [=](uint64_t b, uint64_t e) { for (uint64_t i = b; i < e; i++) occluded[i] = L->scene.occluded(mk(o, d, i), t_max[i]) ? 1 : 0; });
}
// Scene::hit_t (first-found distance)
void oracle_trace_first_found(void* h, const double* o, const double* d, uint64_t n, double* t, int threads) {
    auto* L = (Loaded*)h;
    parallel_for(n, threads, [=](uint64_t b, uint64_t e) { for (uint64_t i = b; i < e; i++) t[i] = L->scene.hit_t(mk(o, d, i)); });
}
void oracle_camera_rays(void* h, const double* raster_xy, const double* lens_uv, uint64_t n, double* o, double* d) {
    auto* L = (Loaded*)h;
    for (uint64_t i = 0; i < n; i++) {
        Ray r = L->camera.generate_ray(Vec2(raster_xy[2 * i], raster_xy[2 * i + 1]), Vec2(lens_uv[2 * i], lens_uv[2 * i + 1]));
        o[3 * i] = r.origin.x; o[3 * i + 1] = r.origin.y; o[3 * i + 2] = r.origin.z; d[3 * i] = r.dir.x; d[3 * i + 1] = r.dir.y; d[3 * i + 2] = r.dir.z;
    }
}
void oracle_counters(uint64_t* out, int reset) {
    std::lock_guard<std::mutex> l(g_total_mu);
    g_total.add(g_cnt); g_cnt = Counters();
    out[0] = g_total.tlas_nodes; out[1] = g_total.inst; out[2] = g_total.kd_nodes; out[3] = g_total.leaf_idx;
    out[4] = g_total.tri_tests; out[5] = g_total.sphere_tests; out[6] = g_total.closest; out[7] = g_total.occlusion;
    if (reset) g_total = Counters();
}

// pixels[W*H*4] (sum r*w, g*w, b*w, w), splats[W*H*3], counters[4] = camera paths, closest-hit
// queries, occlusion queries, reference-style cost.  Buffers are overwritten.
int oracle_render(void* h, const RenderParams* P, double* pixels, double* splats, uint64_t* counters, double* tile_deltas) {
    auto* L = (Loaded*)h;
    FilmAccum film; film.w = L->camera.res_x; film.h = L->camera.res_y;
    film.pixels.assign(film.w * film.h * 4, 0.0); film.splats.assign(film.w * film.h * 3, 0.0);
    { std::lock_guard<std::mutex> l(g_total_mu); g_total = Counters(); g_cnt = Counters(); }
    std::vector<double> deltas;
    if (P->rng_mode == 0) render_reference(*L, *P, film); else render_counter(*L, *P, film, &deltas);
    std::memcpy(pixels, film.pixels.data(), film.pixels.size() * 8);
    std::memcpy(splats, film.splats.data(), film.splats.size() * 8);
    uint64_t c[8]; oracle_counters(c, 0);
    counters[0] = film.cam; counters[1] = c[6]; counters[2] = c[7]; counters[3] = film.cost;
    if (tile_deltas && !deltas.empty()) std::memcpy(tile_deltas, deltas.data(), deltas.size() * 8);
    return 0;
}

// ---- structure export (tests compare the product's flattened blob with these) -------------------
// which: 0 = objects BVH, 1 = lights BVH.  Returns node count; fills up to cap nodes:
// bounds[6], right (-1 = none), first (offset in leaf list), count.
uint64_t oracle_export_bvh(void* h, int which, uint64_t cap, double* bounds, int64_t* right, int64_t* first, int64_t* count, int64_t* leaf_list, uint64_t leaf_cap) {
    auto* L = (Loaded*)h;
    const BVH& b = which == 0 ? L->scene.objects : L->scene.lights;
    uint64_t nl = 0;
    for (size_t i = 0; i < b.nodes.size(); i++) {
        const BVHNode& n = b.nodes[i];
        if (i < cap) {
            bounds[6 * i] = n.bounds.ax_min.x; bounds[6 * i + 1] = n.bounds.ax_min.y; bounds[6 * i + 2] = n.bounds.ax_min.z;
            bounds[6 * i + 3] = n.bounds.ax_max.x; bounds[6 * i + 4] = n.bounds.ax_max.y; bounds[6 * i + 5] = n.bounds.ax_max.z;
            right[i] = n.right == IDX_NAN ? -1 : (int64_t)n.right; first[i] = (int64_t)nl; count[i] = (int64_t)n.objects.size();
        }
        for (size_t o : n.objects) { if (nl < leaf_cap) leaf_list[nl] = (int64_t)o; nl++; }
    }
    return b.nodes.size();
}
static const KdTree* find_kd(const Object* o) {
    if (auto* k = dynamic_cast<const KdTree*>(o)) return k;
    if (auto* r = dynamic_cast<const Rectangle*>(o)) return r->mesh.get();
    if (auto* i = dynamic_cast<const Instance*>(o)) return find_kd(i->object.get());
    return nullptr;
}
// kd-tree of object `idx` (which: 0 objects, 1 lights).  Returns node count (0 if the object has no
// kd-tree); out arrays sized by caller: axis, point, right(-1), leaf flag, first, count; leaf_list.
uint64_t oracle_export_kd(void* h, int which, uint64_t idx, uint64_t cap, int64_t* axis, double* point, int64_t* right, int64_t* leaf,
                          int64_t* first, int64_t* count, int64_t* leaf_list, uint64_t leaf_cap, uint64_t* n_tris, double* tri_verts, uint64_t tri_cap) {
    auto* L = (Loaded*)h;
    const BVH& b = which == 0 ? L->scene.objects : L->scene.lights;
    const KdTree* kd = find_kd(b.objects[idx].get());
    if (!kd) return 0;
    uint64_t nl = 0;
    for (size_t i = 0; i < kd->nodes.size(); i++) {
        const KdNode& n = kd->nodes[i];
        if (i < cap) { axis[i] = n.axis; point[i] = n.point; right[i] = n.right == IDX_NAN ? -1 : (int64_t)n.right; leaf[i] = n.leaf; first[i] = (int64_t)nl; count[i] = (int64_t)n.indices.size(); }
        for (uint32_t t : n.indices) { if (nl < leaf_cap) leaf_list[nl] = t; nl++; }
    }
    *n_tris = kd->objects.size();
    for (size_t i = 0; i < kd->objects.size() && i < tri_cap; i++) {
        const Triangle& t = kd->objects[i];
        Vec3 a = t.a(), bb = t.b(), c = t.c();
        double* q = tri_verts + 9 * i;
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = bb.x; q[4] = bb.y; q[5] = bb.z; q[6] = c.x; q[7] = c.y; q[8] = c.z;
    }
    return kd->nodes.size();
}
uint64_t oracle_export_kd_leaf_total(void* h, int which, uint64_t idx) {
    auto* L = (Loaded*)h;
    const BVH& b = which == 0 ? L->scene.objects : L->scene.lights;
    const KdTree* kd = find_kd(b.objects[idx].get());
    if (!kd) return 0;
    uint64_t nl = 0; for (auto& n : kd->nodes) nl += n.indices.size();
    return nl;
}
// instance transform of object idx: m[16], inv[16] row-major; returns 0 if not an instance
int oracle_export_instance(void* h, int which, uint64_t idx, double* m, double* inv) {
    auto* L = (Loaded*)h;
    const BVH& b = which == 0 ? L->scene.objects : L->scene.lights;
    auto* i = dynamic_cast<const Instance*>(b.objects[idx].get());
    if (!i) return 0;
    const Mat4* src[2] = {&i->transform.m, &i->transform.inv}; double* dst[2] = {m, inv};
    for (int k = 0; k < 2; k++) {
        const Vec4* rows[4] = {&src[k]->y0, &src[k]->y1, &src[k]->y2, &src[k]->y3};
        for (int r = 0; r < 4; r++) { dst[k][4 * r] = rows[r]->x; dst[k][4 * r + 1] = rows[r]->y; dst[k][4 * r + 2] = rows[r]->z; dst[k][4 * r + 3] = rows[r]->w; }
    }
    return 1;
}
void oracle_export_alias(void* h, double* prob, int64_t* alias, double* pdf) {
    auto* L = (Loaded*)h;
    for (size_t i = 0; i < L->scene.lights.alias_table.size(); i++) { prob[i] = L->scene.lights.alias_table[i].first; alias[i] = (int64_t)L->scene.lights.alias_table[i].second; pdf[i] = L->scene.lights.alias_pdf[i]; }
}

// ---- BDPT MIS property (mis_tests.rs:99-352): for complete paths built from a light subpath closed at the camera, the
// MIS weights of all admissible (s,t) splits must sum to one.  out[i] = that sum for path i.
static std::vector<Vertex> reverse_path(std::vector<Vertex> pth) {                                   // mis_tests.rs:318-352
    std::reverse(pth.begin(), pth.end());
    for (size_t i = pth.size(); i-- > 1;) pth[i].wo = -pth[i - 1].wo;
    pth[0].wo = Vec3();
    for (auto& v : pth) std::swap(v.pdf_fwd, v.pdf_bck);
    return pth;
}
uint64_t oracle_mis_sums_from_light(void* h, uint64_t seed, uint64_t n_paths, double* out) {
    auto* L = (Loaded*)h;
    const Scene& sc = L->scene; const Camera& cam = L->camera;
    Rng rng = Rng::xorshift(seed);
    uint64_t done = 0, attempts = 0;
    while (done < n_paths && attempts < 200 * n_paths + 1000) {
        Lambda lam = Lambda::sample(rng.gen_float());
        const Lambda l = lam;
        Ray r = cam.generate_ray(Vec2(0, 0), rng.gen_vec2());                                      // mis_tests.rs:262-316
        Vec3 xc = r.origin;
        std::vector<Vertex> pth;
        bool ok = false;
        for (int tries = 0; tries < 200 && !ok; tries++, attempts++) {
            pth = light_path(sc, rng, 0.0, lam);
            if (pth.size() <= 2) continue;
            const Vertex& ls = pth.back();
            if (ls.is_delta(lam)) continue;
            const Vertex& ls_m = pth[pth.size() - 2];
            Vec3 xo = ls_m.h.p, wo = ls.wo, ngo = ls_m.h.ng, ngi = ls.h.ng, xi = ls.h.p;
            Float t2 = xi.distance_squared(xc);
            Vec3 wi = (xi - xc).normalize();
            Ray ri = Ray::make(xc, wi);
            Hit hh;
            if (sc.hit(ri, hh) && hh.t * hh.t < t2 - EPSILON * EPSILON) continue;
            pth.push_back(Vertex::camera(xc, 0.0, WHITE));
            size_t len = pth.size();
            pth[len - 1].wo = wi;
            Float pdf_sa = cam.pdf_wi(ri);
            Vec3 ngi2 = !pth[len - 2].is_surface() ? wi : ngi;
            pth[len - 2].pdf_bck = sa_to_area(pdf_sa, xc, xi, wi, ngi2);
            if (!pth[len - 3].is_delta(lam)) {
                Float p2 = pth[len - 2].bsdf_pdf(-wi, lam, true);
                Vec3 ngo2 = !pth[len - 3].is_surface() ? wo : ngo;
                pth[len - 3].pdf_bck = sa_to_area(p2, xo, xi, wo, ngo2);
            }
            ok = true;
        }
        if (!ok) continue;
        std::vector<Vertex> lp = pth, cp = reverse_path(pth);
        Float sumw = 0.0;
        for (size_t s = 0; s < lp.size(); s++) {                                                    // mis_tests.rs:119-152
            size_t t = lp.size() - s;
            if (t == 1 && s < 2) continue;
            if (lp[s].is_delta(l) || (s > 0 && lp[s - 1].is_delta(l))) continue;
            sumw += mis_weight(sc, cam, l, lp.data(), s, cp.data(), t);
        }
        out[done++] = sumw;
    }
    return done;
}

// ---- chi^2 material for BxDF sampling vs pdf (bxdf/chi2_tests.rs:94-236): histogram of sampled directions and the pdf
// integrated over the same (theta, phi) bins, both in the local shading frame of a z-up surface.
void oracle_bsdf_chi2_tables(void* h, int mat, const double* wo3, double lambda_u, uint64_t seed, uint64_t n_samples,
                             int theta_bins, int phi_bins, int sub, double* observed, double* expected) {
    auto* L = (Loaded*)h;
    const Material* m = L->scene.materials[mat].get();
    Vec3 wo(wo3[0], wo3[1], wo3[2]);
    Hit hh; hh.ng = Vec3(0, 0, 1); hh.ns = Vec3(0, 0, 1); hh.backface = false; hh.material = m;
    Rng rng = Rng::xorshift(seed);
    Lambda lam = Lambda::sample(lambda_u);
    for (int i = 0; i < theta_bins * phi_bins; i++) { observed[i] = 0.0; expected[i] = 0.0; }
    for (uint64_t i = 0; i < n_samples; i++) {                                                    // chi2_tests.rs:172-200
        Vec3 wi;
        Float ru = rng.gen_float(); Vec2 rs = rng.gen_vec2();
        if (!m->bsdf_sample(wo, hh, lam, ru, rs, wi)) continue;
        Float theta = lm_acos(clampf(wi.z, -1.0, 1.0)), phi = lm_atan2(wi.y, wi.x); if (phi < 0.0) phi += 2.0 * PI;
        int tb = std::min((int)(theta / PI * theta_bins), theta_bins - 1), pb = std::min((int)(phi / (2.0 * PI) * phi_bins), phi_bins - 1);
        observed[pb + tb * phi_bins] += 1.0;
    }
    for (int tb = 0; tb < theta_bins; tb++) for (int pb = 0; pb < phi_bins; pb++) {               // chi2_tests.rs:203-236 (midpoint rule instead of Simpson)
        Float acc = 0.0;
        for (int a = 0; a < sub; a++) for (int b = 0; b < sub; b++) {
            Float theta = (tb + (a + 0.5) / sub) * PI / theta_bins, phi = (pb + (b + 0.5) / sub) * 2.0 * PI / phi_bins;
            Vec3 wi(lm_sin(theta) * lm_cos(phi), lm_sin(theta) * lm_sin(phi), lm_cos(theta));
            acc += m->bsdf_pdf(wo, wi, hh, lam, false) * lm_sin(theta);
        }
        expected[pb + tb * phi_bins] = acc * (PI / theta_bins / sub) * (2.0 * PI / phi_bins / sub) * (Float)n_samples;
    }
}

// ---- unit hooks used by the property tests ------------------------------------------------------
// ---- textures (texture.rs, perlin.rs, image.rs) evaluated directly, for the texture property tests
uint64_t oracle_texture_count(void* h) { return ((Loaded*)h)->textures.size(); }
int oracle_texture_kind(void* h, uint64_t tex) { return ((Loaded*)h)->textures.at(tex)->kind; }
void oracle_texture_eval(void* h, uint64_t tex, const double* uv, uint64_t n, double lambda_u, double* out4) {
    const Texture* t = ((Loaded*)h)->textures.at(tex).get();
    Lambda lam = Lambda::sample(lambda_u);
    for (uint64_t i = 0; i < n; i++) { Color c = t->albedo_at(lam, Vec2(uv[2 * i], uv[2 * i + 1])); for (int k = 0; k < 4; k++) out4[4 * i + k] = c.s[k]; }
}
void oracle_bump_eval(void* h, uint64_t tex, const double* uv, uint64_t n, double* out3) {
    const Texture* t = ((Loaded*)h)->textures.at(tex).get();
    for (uint64_t i = 0; i < n; i++) { Vec3 v = t->bump.value_at(Vec2(uv[2 * i], uv[2 * i + 1])); out3[3 * i] = v.x; out3[3 * i + 1] = v.y; out3[3 * i + 2] = v.z; }
}
void oracle_perlin_noise(uint64_t seed, const double* p3, uint64_t n, double* out) {
    Perlin pn(seed);
    for (uint64_t i = 0; i < n; i++) out[i] = pn.noise_at(Vec3(p3[3 * i], p3[3 * i + 1], p3[3 * i + 2]));
}
void oracle_perlin_tables(uint64_t seed, double* out1536) {   // lattice (768) then perm x, y, z: the layout of the device blob's f64 pool
    Perlin pn(seed);
    for (int i = 0; i < 256; i++) { out1536[3 * i] = pn.lattice[i].x; out1536[3 * i + 1] = pn.lattice[i].y; out1536[3 * i + 2] = pn.lattice[i].z; }
    for (int i = 0; i < 256; i++) { out1536[768 + i] = (double)pn.px[i]; out1536[1024 + i] = (double)pn.py[i]; out1536[1280 + i] = (double)pn.pz[i]; }
}
double oracle_lambda_sample_one(double v) { return Lambda::sample_one(v); }
// lumo_math.h as compiled for the oracle: fn 0 sin, 1 cos, 2 atan2(x, y), 3 acos, 4 atanh, 5 cosh, 6 exp, 7 log, 8 pow(x, y)
void oracle_math_eval(int fn, const double* x, const double* y, uint64_t n, double* out) {
    for (uint64_t i = 0; i < n; i++) {
        const double a = x[i], b = y ? y[i] : 0.0;
        switch (fn) {
        case 0: out[i] = lm_sin(a); break; case 1: out[i] = lm_cos(a); break; case 2: out[i] = lm_atan2(a, b); break; case 3: out[i] = lm_acos(a); break;
        case 4: out[i] = lm_atanh(a); break; case 5: out[i] = lm_cosh(a); break; case 6: out[i] = lm_exp(a); break; case 7: out[i] = lm_log(a); break;
        default: out[i] = lm_pow(a, b); break;
        }
    }
}
// samples of one pixel from the reference samplers (kind 0..3): out[2i], out[2i+1]; returns how many
uint64_t oracle_sampler_points(int kind, uint64_t batch, uint64_t samples, uint64_t seed, uint64_t cap, double* out) {
    Sampler sm(kind, batch, samples, seed); Vec2 v; uint64_t n = 0;
    while (n < cap && sm.next(v)) { out[2 * n] = v.x; out[2 * n + 1] = v.y; n++; }
    return n;
}
void oracle_xorshift(uint64_t seed, uint64_t n, uint64_t* out) { Rng r = Rng::xorshift(seed); for (uint64_t i = 0; i < n; i++) out[i] = r.gen_u64(); }
void oracle_philox(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stream, uint64_t n, uint64_t* out) { Rng r = Rng::counter(seed, pixel, sample, stream); for (uint64_t i = 0; i < n; i++) out[i] = r.gen_u64(); }
double oracle_spectrum_sample(const float* c, double lambda) { Spectrum s; s.c0 = c[0]; s.c1 = c[1]; s.c2 = c[2]; s.scale = c[3]; return s.sample_one(lambda); }
void oracle_wb_matrix(int cs_id, int illum, double* out9, double* xyz2rgb9) {
    ColorSpace cs = ColorSpace::get(cs_id);
    Mat3 wb = cs.wb_matrix(illuminant_table(illum));
    const Vec3* r[3] = {&wb.y0, &wb.y1, &wb.y2}; const Vec3* q[3] = {&cs.XYZ_to_RGB.y0, &cs.XYZ_to_RGB.y1, &cs.XYZ_to_RGB.y2};
    for (int i = 0; i < 3; i++) { out9[3 * i] = r[i]->x; out9[3 * i + 1] = r[i]->y; out9[3 * i + 2] = r[i]->z; xyz2rgb9[3 * i] = q[i]->x; xyz2rgb9[3 * i + 1] = q[i]->y; xyz2rgb9[3 * i + 2] = q[i]->z; }
}
double oracle_filter_eval(int kind, double r, double p, double x, double y) { PixelFilter f; f.kind = kind; f.r = r; f.p = p; return f.eval(Vec2(x, y)); }
double oracle_filter_integral(int kind, double r, double p) { PixelFilter f; f.kind = kind; f.r = r; f.p = p; return f.integral(); }
// BSDF hooks in the local shading frame for white-furnace / chi2 style tests (material index in program order)
int oracle_bsdf_sample(void* h, int mat, const double* wo, double lambda_u, double rand_u, double r0, double r1, int mode, double* wi, double* f4, double* pdf) {
    auto* L = (Loaded*)h;
    const Material* m = L->scene.materials[mat].get();
    Lambda lam = Lambda::sample(lambda_u);
    Vec3 w(wo[0], wo[1], wo[2]), out;
    Hit hh; hh.ng = Vec3(0, 0, 1); hh.ns = Vec3(0, 0, 1); hh.backface = false; hh.material = m;
    if (!m->bsdf_sample(w, hh, lam, rand_u, Vec2(r0, r1), out)) return 0;
    wi[0] = out.x; wi[1] = out.y; wi[2] = out.z;
    Color f = m->bsdf_f(w, out, lam, mode, hh);
    for (int i = 0; i < 4; i++) f4[i] = f.s[i];
    *pdf = m->bsdf_pdf(w, out, hh, lam, false);
    return 1;
}
}
