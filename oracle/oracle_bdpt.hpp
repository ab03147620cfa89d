// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).  PARITY UNPINNED.
// CPU restatement of lumo's bidirectional path tracer:
// src/tracer/integrator/bd_path_trace.rs and bd_path_trace/{path_gen,vertex,mis,measure}.rs.
#pragma once
#include "oracle_shading.hpp"

namespace oracle {

static const size_t BDPT_MAX_DEPTH = 1024;   // bd_path_trace.rs:7
static const Material& blank_material() { static Material m; return m; }

static inline Float sa_to_area(Float pdf, Vec3 xo, Vec3 xi, Vec3 wi, Vec3 ngi) {               // measure.rs:9-11
    return pdf * std::fabs(wi.dot(ngi)) / xo.distance_squared(xi);
}

struct Vertex {                                                                                 // vertex.rs:5-12
    Hit h; Color gathered; Float pdf_fwd = 0, pdf_bck = 0; Vec3 wo; int64_t light = -1;

    static Vertex camera(Vec3 xo, Float pdf_fwd, Color gathered) {                              // vertex.rs:16-36
        Vertex v;
        v.h = hit_new(0.0, &blank_material(), -Vec3(1, 0, 0), xo, Vec3(), Vec3(1, 0, 0), Vec3(1, 0, 0), Vec2(1, 0));
        v.gathered = gathered; v.pdf_fwd = pdf_fwd;
        return v;
    }
    static Vertex mk_light(const Hit& h, size_t light, Color gathered, Float pdf_fwd) {         // vertex.rs:39-48
        Vertex v; v.h = h; v.gathered = gathered; v.light = (int64_t)light; v.pdf_fwd = pdf_fwd; return v;
    }
    static Vertex surface(Vec3 wo, const Hit& h, Color gathered, Float pdf_sa, const Lambda& lam, const Vertex& prev) {  // vertex.rs:51-84
        Vertex v;
        v.pdf_fwd = h.material->is_delta(lam) ? 0.0 : sa_to_area(pdf_sa, prev.h.p, h.p, -wo, h.ng);
        v.h = h; v.gathered = gathered; v.wo = wo;
        return v;
    }
    const Material* material() const { return h.material; }
    bool is_surface() const { return material()->kind != M_BLANK; }
    bool is_light() const { return light >= 0; }
    bool is_delta(const Lambda& lam) const { return material()->is_delta(lam); }
    Color emittance(const Lambda& lam) const { return material()->emit(lam, h); }
    Float shading_cosine(Vec3 wi) const { return material()->shading_cosine(wi, h.ns); }
    Float shading_correction(Vec3 wi) const {                                                   // vertex.rs:118-126
        const Material* m = material();
        return m->shading_cosine(wi, h.ng) * m->shading_cosine(wo, h.ns) / (m->shading_cosine(wo, h.ng) * m->shading_cosine(wi, h.ns));
    }
    Color f(const Vertex& next, const Lambda& lam, int mode) const {                            // vertex.rs:129-132
        Vec3 wi = (next.h.p - h.p).normalize();
        return material()->bsdf_f(wo, wi, lam, mode, h);
    }
    Float bsdf_pdf(Vec3 wi, const Lambda& lam, bool swap_dir) const { return material()->bsdf_pdf(wo, wi, h, lam, swap_dir); }
    Float pdf_prev(const Vertex& prev, Vec3 wi, const Lambda& lam) const {                      // vertex.rs:146-160
        if (is_delta(lam) || prev.is_delta(lam)) return 0.0;
        Float pdf_sa = bsdf_pdf(wi, lam, true);
        Vec3 ngp = !prev.is_surface() ? -wo : prev.h.ng;
        return sa_to_area(pdf_sa, h.p, prev.h.p, -wo, ngp);
    }
};

static inline std::vector<Vertex> walk(const Scene& sc, Ray ro, Rng& rng, Lambda& lam, Float delta, const Vertex& root,
                                       Color gathered, Float pdf_dir, int mode) {               // path_gen.rs:53-157
    size_t depth = 0;
    std::vector<Vertex> vs; vs.reserve(8);
    vs.push_back(root);
    Float pdf_fwd = pdf_dir;
    Hit ho;
    while (sc.hit(ro, ho)) {
        const Material* m = ho.material;
        gathered = gathered * WHITE;
        size_t prev = depth;
        Vec3 wo = -ro.dir;
        vs.push_back(Vertex::surface(wo, ho, gathered, pdf_fwd, lam, vs[prev]));
        depth += 1;
        size_t curr = depth;
        Float ru = rng.gen_float(); Vec2 rs = rng.gen_vec2();
        Vec3 wi;
        if (!m->bsdf_sample(wo, vs[curr].h, lam, ru, rs, wi)) {
            if (mode == IMPORTANCE) vs.pop_back();
            else vs[curr].light = sc.get_light_at(vs[curr].h);
            break;
        }
        Ray ri = hit_generate_ray(vs[curr].h, wi);
        wi = ri.dir;
        pdf_fwd = m->bsdf_pdf(wo, wi, vs[curr].h, lam, false);
        if (pdf_fwd == 0.0) break;
        Float corr = mode == RADIANCE ? 1.0 : vs[curr].shading_correction(wi);
        Color bsdf = m->bsdf_f(wo, wi, lam, mode, vs[curr].h);
        gathered = gathered * (bsdf * vs[curr].shading_cosine(wi) * corr / pdf_fwd);
        vs[prev].pdf_bck = vs[curr].pdf_prev(vs[prev], wi, lam);
        if (depth >= RR_DEPTH) {
            Float lum = color_luminance(gathered, lam);
            Float rr = fmin_(lum / delta, 1.0);
            if (rng.gen_float() > rr) break;
            if (depth >= BDPT_MAX_DEPTH) break;
            gathered = gathered / rr;
        }
        if (m->is_delta(lam)) pdf_fwd = 0.0;
        ro = ri;
    }
    return vs;
}
static inline std::vector<Vertex> camera_path(const Scene& sc, const Camera& cam, Ray r, Rng& rng, Float delta, Lambda& lam) {  // path_gen.rs:4-19
    Color gathered = WHITE;
    Float pdf_wi = cam.pdf_wi(r), pdf_xo = cam.pdf_xo(r);
    Vertex root = Vertex::camera(r.origin, pdf_xo, gathered);
    return walk(sc, r, rng, lam, delta, root, gathered, pdf_wi, RADIANCE);
}
static inline std::vector<Vertex> light_path(const Scene& sc, Rng& rng, Float delta, Lambda& lam) {  // path_gen.rs:21-50
    size_t light_idx = sc.sample_light(rng.gen_float());
    Float pdf_light;
    const Object* light = sc.get_light(light_idx, pdf_light);
    Vec2 r0 = rng.gen_vec2(); Vec2 r1 = rng.gen_vec2();
    Ray ri; Hit ho;
    light->sample_leaving(r0, r1, ri, ho);
    Float pdf_origin, pdf_dir;
    light->sample_leaving_pdf(ri, ho.ng, pdf_origin, pdf_dir);
    Color emit = ho.material->emit(lam, ho);
    Vertex root = Vertex::mk_light(ho, light_idx, emit, pdf_origin * pdf_light);
    Color gathered = emit * std::fabs(ri.dir.dot(ho.ns)) / (pdf_light * pdf_origin * pdf_dir);
    return walk(sc, ri, rng, lam, delta, root, gathered, pdf_dir, IMPORTANCE);
}

// ---- MIS (mis.rs) -------------------------------------------------------------------------------
static inline Float pdf_light_leaving(const Vertex& curr, const Vertex& next, const Scene& sc, const Lambda& lam) {  // mis.rs:4-33
    if (next.is_delta(lam)) return 0.0;
    if (curr.light < 0) return 0.0;
    Vec3 xo = curr.h.p, xi = next.h.p;
    Ray ri = Ray::make(xo, xi - xo);
    Vec3 wi = ri.dir;
    Float pl; const Object* light = sc.get_light((size_t)curr.light, pl);
    Float po, pdf_dir; light->sample_leaving_pdf(ri, curr.h.ng, po, pdf_dir);
    Vec3 ngi = !next.is_surface() ? wi : next.h.ng;
    return sa_to_area(pdf_dir, xo, xi, wi, ngi);
}
static inline Float pdf_camera_leaving(const Vertex& curr, const Vertex& next, const Camera& cam, const Lambda& lam) {  // mis.rs:36-54
    if (next.is_delta(lam)) return 0.0;
    Vec3 xo = curr.h.p, xi = next.h.p;
    Vec3 wi = (xi - xo).normalize();
    Float pdf_wi = cam.pdf_wi(Ray::make(xo, wi));
    Vec3 ngi = !next.is_surface() ? wi : next.h.ng;
    return sa_to_area(pdf_wi, xo, xi, wi, ngi);
}
static inline Float pdf_light_origin(const Vertex& v, const Scene& sc) {                        // mis.rs:57-64
    if (v.light < 0) return 0.0;
    Float pl; const Object* light = sc.get_light((size_t)v.light, pl);
    return pl / light->area();
}
static inline Float pdf_connection(const Vertex& curr, const Vertex& next, const Lambda& lam, const Vertex* prev) {  // mis.rs:69-96
    if (next.is_delta(lam)) return 0.0;
    Vec3 xo = curr.h.p, xi = next.h.p;
    Float pdf_sa; Vec3 wi;
    if (prev) { Vec3 wo = (prev->h.p - xo).normalize(); pdf_sa = curr.bsdf_pdf(wo, lam, true); wi = curr.wo; }
    else { wi = (xi - xo).normalize(); pdf_sa = curr.bsdf_pdf(wi, lam, false); }
    Vec3 ngi = !next.is_surface() ? wi : next.h.ng;
    return sa_to_area(pdf_sa, xo, xi, wi, ngi);
}
static inline Float mis_weight(const Scene& sc, const Camera& cam, const Lambda& lam, const Vertex* lp, size_t s, const Vertex* cp, size_t t) {  // mis.rs:103-239
    if (s + t == 2) return 1.0;
    auto map0 = [](Float p) { return p == 0.0 ? 1.0 : p; };
    const Vertex& ct1 = cp[t - 1];
    const Vertex& ls1 = s == 0 ? cp[0] : lp[s - 1];
    std::vector<Float> pdf_rad, pdf_imp; std::vector<char> is_delta;
    for (size_t i = 0; i + 2 < std::max<size_t>(s, 2); i++) { pdf_rad.push_back(lp[i].pdf_bck); pdf_imp.push_back(lp[i].pdf_fwd); is_delta.push_back(lp[i].is_delta(lam)); }
    if (s > 1) {
        const Vertex& ls2 = lp[s - 2];
        pdf_rad.push_back(pdf_connection(ls1, ls2, lam, &ct1)); pdf_imp.push_back(ls2.pdf_fwd); is_delta.push_back(ls2.is_delta(lam));
    }
    if (s > 0) {
        Float pb = t == 1 ? pdf_camera_leaving(ct1, ls1, cam, lam) : pdf_connection(ct1, ls1, lam, nullptr);
        pdf_rad.push_back(pb); pdf_imp.push_back(ls1.pdf_fwd); is_delta.push_back(0);
    }
    if (t > 0) {
        Float pb = s == 0 ? pdf_light_origin(ct1, sc) : (s == 1 ? pdf_light_leaving(ls1, ct1, sc, lam) : pdf_connection(ls1, ct1, lam, nullptr));
        pdf_rad.push_back(ct1.pdf_fwd); pdf_imp.push_back(pb); is_delta.push_back(0);
    }
    if (t > 1) {
        const Vertex& ct2 = cp[t - 2];
        Float pb = s == 0 ? pdf_light_leaving(ct1, ct2, sc, lam) : pdf_connection(ct1, ct2, lam, &ls1);
        pdf_rad.push_back(ct2.pdf_fwd); pdf_imp.push_back(pb); is_delta.push_back(ct2.is_delta(lam));
    }
    for (size_t i = std::max<size_t>(t, 2) - 2; i-- > 0;) { pdf_rad.push_back(cp[i].pdf_fwd); pdf_imp.push_back(cp[i].pdf_bck); is_delta.push_back(cp[i].is_delta(lam)); }
    Float sum_ri = 0.0, ri = 1.0;
    for (size_t i = s; i-- > 0;) {
        ri *= map0(pdf_rad[i]) / map0(pdf_imp[i]);
        if (!is_delta[i] && !(i > 0 && is_delta[i - 1])) sum_ri += ri * ri;
    }
    ri = 1.0; sum_ri += ri;
    for (size_t i = s; i + 1 < s + t; i++) {
        ri *= map0(pdf_imp[i]) / map0(pdf_rad[i]);
        if (!is_delta[i] && !is_delta[i + 1]) sum_ri += ri * ri;
    }
    return 1.0 / sum_ri;
}

// ---- connections (bd_path_trace.rs:77-290) -----------------------------------------------------
static inline bool connect_light_path(const Scene& sc, const Camera& cam, Rng& rng, const Lambda& lam, const Vertex* lp, size_t s, FilmSample& out) {  // :77-135
    const Vertex& ll = lp[s - 1];
    if (ll.is_delta(lam)) return false;
    Vec3 xi = ll.h.p;
    Ray ri;
    if (!cam.sample_towards(xi, rng.gen_vec2(), ri)) return false;
    Vec3 xo = ri.origin, wi = ri.dir;
    Float p_sct = ll.bsdf_pdf(-wi, lam, false);
    Float p_imp = cam.pdf_importance(ri, xi);
    if (p_sct == 0.0 || p_imp == 0.0) return false;
    Hit hh;
    if (!sc.hit(ri, hh)) return false;
    if ((hh.p - xi).abs().max_element() > std::sqrt(EPSILON)) return false;
    Color color; Vec2 raster;
    if (!cam.sample_importance(ri, color, raster)) return false;
    if (color.is_black()) return false;
    color = color / p_imp;
    Float p_xo = cam.pdf_xo(ri);
    Vertex cl = Vertex::camera(xo, p_xo, color / p_imp);
    Float t2 = xo.distance_squared(xi);
    (void)t2;
    color = color * (ll.gathered * WHITE * ll.shading_cosine(-wi) * ll.shading_correction(-wi)
                     * ll.f(cl, lam, IMPORTANCE) * mis_weight(sc, cam, lam, lp, s, &cl, 1));
    out = FilmSample{raster, color, lam, true, 0};
    return true;
}
static inline Color add_camera_path(const Scene& sc, const Camera& cam, const Lambda& lam, const Vertex* cp, size_t t) {  // :137-156
    if (!cp[t - 1].is_light()) return BLACK;
    const Vertex& ct = cp[t - 1];
    Color rad = ct.gathered * ct.emittance(lam);
    if (rad.is_black()) return BLACK;
    return rad * mis_weight(sc, cam, lam, nullptr, 0, cp, t);
}
static inline Color connect_camera_path(const Scene& sc, const Camera& cam, Rng& rng, const Lambda& lam, const Vertex* cp, size_t t) {  // :158-213
    if (cp[t - 1].is_delta(lam) || cp[t - 1].is_light()) return BLACK;
    const Vertex& cl = cp[t - 1];
    size_t light_idx = sc.sample_light(rng.gen_float());
    Float pdf_light; const Object* light = sc.get_light(light_idx, pdf_light);
    const Hit& ho = cl.h;
    Vec3 xo = ho.p;
    Vec3 wi = light->sample_towards(xo, rng.gen_vec2());
    Float p_sct = cl.bsdf_pdf(wi, lam, false);
    if (p_sct == 0.0) return BLACK;
    Ray ri = hit_generate_ray(ho, wi);
    Hit hi;
    if (!sc.hit_light(ri, light, hi)) return BLACK;
    Vec3 xi = hi.p;
    Vec3 ngi = !cl.is_surface() ? wi : hi.ng;
    Float p_lig = light->sample_towards_pdf(ri, xi, ngi) * pdf_light;
    if (p_lig == 0.0) return BLACK;
    wi = ri.dir;
    Float pdf_origin = sa_to_area(p_lig, xo, xi, wi, ngi);
    Color emittance = hi.material->emit(lam, hi);
    Vertex ll = Vertex::mk_light(hi, light_idx, emittance, pdf_origin);
    Color bsdf = cl.f(ll, lam, RADIANCE);
    Float cos_wi = cl.shading_cosine(wi);
    Color radiance = cl.gathered * bsdf * emittance * WHITE * cos_wi / p_lig;
    return radiance * mis_weight(sc, cam, lam, &ll, 1, cp, t);
}
static inline bool visible(const Scene& sc, const Hit& h1, const Hit& h2) {                     // :279-290
    Vec3 xo = h1.p, xi = h2.p;
    Ray ri = hit_generate_ray(h1, xi - xo);
    Vec3 wi = ri.dir;
    if (wi.dot(h1.ng) < EPSILON) return false;
    return std::fabs(xo.distance(xi) - sc.hit_t(ri)) < EPSILON;
}
static inline Color connect_paths(const Scene& sc, const Camera& cam, const Lambda& lam, const Vertex* lp, size_t s, const Vertex* cp, size_t t) {  // :216-277
    const Vertex& ll = lp[s - 1]; const Vertex& cl = cp[t - 1];
    if (cl.is_delta(lam) || cl.is_light() || ll.is_delta(lam) || !visible(sc, ll.h, cl.h)) return BLACK;
    Vec3 xc = cl.h.p, xl = ll.h.p;
    Vec3 wi = (xl - xc).normalize();
    Float p_sct = cl.bsdf_pdf(wi, lam, false) * ll.bsdf_pdf(-wi, lam, false);
    if (p_sct == 0.0) return BLACK;
    Color lb = ll.f(cl, lam, IMPORTANCE), cb = cl.f(ll, lam, RADIANCE);
    Color radiance = ll.gathered * lb * ll.shading_cosine(-wi) * cl.gathered * cb * cl.shading_cosine(wi) * WHITE / xc.distance_squared(xl);
    if (radiance.is_black()) return BLACK;
    return radiance * mis_weight(sc, cam, lam, lp, s, cp, t);
}

static inline std::vector<FilmSample> bdpt_integrate(const Scene& sc, const Camera& cam, Ray r, Rng& rng, Lambda lam, Float delta, Vec2 raster_xy) {  // :23-75
    std::vector<Vertex> lp = light_path(sc, rng, delta, lam);
    std::vector<Vertex> cp = camera_path(sc, cam, r, rng, delta, lam);
    Color radiance = BLACK;
    std::vector<FilmSample> samples;
    size_t cost = lp.size() + cp.size();
    for (size_t s = 2; s <= lp.size(); s++) {
        if (!lp[s - 1].is_delta(lam)) cost += 1;
        FilmSample fs;
        if (connect_light_path(sc, cam, rng, lam, lp.data(), s, fs)) samples.push_back(fs);
    }
    radiance = radiance + add_camera_path(sc, cam, lam, cp.data(), cp.size());
    for (size_t t = 2; t <= cp.size(); t++) {
        if (!cp[t - 1].is_delta(lam) && cp[t - 1].is_light()) cost += 1;
        radiance = radiance + connect_camera_path(sc, cam, rng, lam, cp.data(), t);
    }
    for (size_t t = 2; t <= cp.size(); t++) for (size_t s = 2; s <= lp.size(); s++) {
        cost += 1;
        radiance = radiance + connect_paths(sc, cam, lam, lp.data(), s, cp.data(), t);
    }
    samples.push_back(FilmSample{raster_xy, radiance, lam, false, cost});
    return samples;
}

}  // namespace oracle
