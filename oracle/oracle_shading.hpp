// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).  PARITY UNPINNED.
// CPU restatement of lumo's colour, material/BSDF, camera, film, sampler, scene and the
// PathTrace / DirectLight integrators.  Citations are to /root/reference/src/... file:line.
#pragma once
#include "oracle_geom.hpp"
#include "oracle_spectra_data.h"

namespace oracle {

enum Transport { RADIANCE = 0, IMPORTANCE = 1 };

// ---- Color (src/tracer/color.rs) --------------------------------------------------------------
static const Float LAMBDA_MIN = 360.0, LAMBDA_MAX = 830.0;
struct Color {
    Float s[4];
    Color() { s[0] = s[1] = s[2] = s[3] = 0; }
    explicit Color(Float v) { s[0] = s[1] = s[2] = s[3] = v; }
    bool is_black() const { return s[0] == 0.0 && s[1] == 0.0 && s[2] == 0.0 && s[3] == 0.0; }
    Float mean() const { return (((0.0 + s[0]) + s[1]) + s[2] + s[3]) / 4.0; }
    Float maxv() const { Float m = -INF; for (int i = 0; i < 4; i++) m = fmax_(m, s[i]); return m; }
};
#define ORC_COLOR_OP(op) \
    static inline Color operator op(Color a, Color b) { Color r; for (int i = 0; i < 4; i++) r.s[i] = a.s[i] op b.s[i]; return r; } \
    static inline Color operator op(Color a, Float b) { Color r; for (int i = 0; i < 4; i++) r.s[i] = a.s[i] op b; return r; } \
    static inline Color operator op(Float a, Color b) { Color r; for (int i = 0; i < 4; i++) r.s[i] = a op b.s[i]; return r; }
ORC_COLOR_OP(+) ORC_COLOR_OP(-) ORC_COLOR_OP(*)
static inline Color operator/(Color a, Color b) { Color r; for (int i = 0; i < 4; i++) r.s[i] = b.s[i] == 0.0 ? 0.0 : a.s[i] / b.s[i]; return r; }  // color.rs:239-249
static inline Color operator/(Color a, Float b) { Color r; for (int i = 0; i < 4; i++) r.s[i] = b == 0.0 ? 0.0 : a.s[i] / b; return r; }           // color.rs:251-271
static const Color WHITE(1.0), BLACK(0.0);

// ---- ColorWavelength (src/tracer/color/wavelength.rs) -----------------------------------------
struct Lambda {
    Float l[4];
    static Float sample_one(Float v) { return 538.0 - 138.888889 * lm_atanh(0.85691062 - 253.819 * v * 0.0072); }  // :48-51
    static Float pdf_one(Float lam) {                                                                                // :60-66
        if (lam < LAMBDA_MIN || lam > LAMBDA_MAX) return 0.0;
        return 1.0 / (253.819 * powi(lm_cosh(0.0072 * (lam - 538.05)), 2));
    }
    static Lambda sample(Float u) {                                                                                  // :35-44
        Lambda r;
        for (int i = 0; i < 4; i++) {
            Float v = u + (Float)i / 4.0;
            v = v > 1.0 ? v - 1.0 : v;
            r.l[i] = sample_one(v);
        }
        return r;
    }
    bool is_terminated() const { return l[1] == 0.0 && l[2] == 0.0 && l[3] == 0.0; }
    Float terminate() { l[1] = l[2] = l[3] = 0.0; return l[0]; }
    Float leading() const { return l[0]; }
    Color pdf() const {                                                                                              // :24-32
        Color c; for (int i = 0; i < 4; i++) c.s[i] = pdf_one(l[i]);
        if (is_terminated()) c.s[0] /= 4.0;
        return c;
    }
};

// ---- DenseSpectrum (src/tracer/color/dense_spectrum.rs) ---------------------------------------
struct DenseSpectrum {
    Float v[95];
    static DenseSpectrum from(const double* p) { DenseSpectrum d; for (int i = 0; i < 95; i++) d.v[i] = p[i]; return d; }
    static DenseSpectrum constant(Float c) { DenseSpectrum d; for (int i = 0; i < 95; i++) d.v[i] = c; return d; }
    bool is_constant() const { for (int i = 0; i < 95; i++) if (v[i] != v[0]) return false; return true; }
    static Float sample_one_raw(const double* v, Float lambda) {                                                     // :77-97
        const Float STEP = (LAMBDA_MAX - LAMBDA_MIN) / (95.0 - 1.0);
        uint64_t b1 = sat_u64(std::ceil((lambda - LAMBDA_MIN) / STEP));
        Float l1 = LAMBDA_MIN + STEP * (Float)b1;
        if (lambda == 0.0) return 0.0;
        if (lambda == l1) return v[b1];
        assert(b1 >= 1 && b1 < 95);
        uint64_t b0 = b1 - 1;
        Float l0 = l1 - STEP;
        Float x1 = (lambda - l0) / STEP, x0 = 1.0 - x1;
        return v[b0] * x0 + v[b1] * x1;
    }
    static Color sample_raw(const double* v, const Lambda& lam) { Color c; for (int i = 0; i < 4; i++) c.s[i] = sample_one_raw(v, lam.l[i]); return c; }
    Float sample_one(Float lambda) const { return sample_one_raw(v, lambda); }
    Color sample(const Lambda& lam) const { return sample_raw(v, lam); }
    Float dot(const double* rhs) const { Float sum = 0.0; for (int i = 0; i < 95; i++) sum += v[i] * rhs[i]; return sum; }
    Vec3 to_xyz() const {                                                                                            // :99-105
        const Float YI = 106.856895;
        return Vec3(dot(spectra::X) / YI, dot(spectra::Y) / YI, dot(spectra::Z) / YI);
    }
};
static const Float Y_INTEGRAL = 106.856895;   // color/xyz.rs:32
static const DenseSpectrum& CIE_X() { static DenseSpectrum d = DenseSpectrum::from(spectra::X); return d; }
static const DenseSpectrum& CIE_Y() { static DenseSpectrum d = DenseSpectrum::from(spectra::Y); return d; }
static const DenseSpectrum& CIE_Z() { static DenseSpectrum d = DenseSpectrum::from(spectra::Z); return d; }
static inline const double* illuminant_table(int id) {
    switch (id) { case 0: return spectra::A; case 1: return spectra::D50; case 2: return spectra::D65;
                  case 3: return spectra::F2; case 4: return spectra::F7; default: return spectra::CORNELL; }
}

static inline Float color_luminance(const Color& c, const Lambda& lam) {                                              // color.rs:88-91
    Color pdf = lam.pdf();
    return (CIE_Y().sample(lam) * c / pdf).mean() / Y_INTEGRAL;
}
static inline Vec3 color_xyz(const Color& c, const Lambda& lam) {                                                     // color.rs:93-101
    Color pdf = lam.pdf();
    return Vec3((CIE_X().sample(lam) * c / pdf).mean(), (CIE_Y().sample(lam) * c / pdf).mean(),
                (CIE_Z().sample(lam) * c / pdf).mean()) / Y_INTEGRAL;
}

// ---- Spectrum (src/tracer/color/spectrum.rs): sigmoid polynomial, f32 ---------------------------
struct Spectrum {
    float c0 = 0, c1 = 0, c2 = 0, scale = 0;
    Float sample_one(Float lambda) const {                                                                           // :108-118
        float l = (float)lambda;
        float x = c0 * l * l + c1 * l + c2;
        float sg = 0.5f + x / (2.0f * std::sqrt(1.0f + x * x));
        return (Float)(scale * sg);
    }
    Color sample(const Lambda& lam) const { Color c; for (int i = 0; i < 4; i++) c.s[i] = sample_one(lam.l[i]); return c; }
};

// ---- Textures (src/tracer/texture.rs, src/perlin.rs, src/image.rs) ------------------------------
struct Perlin {                                                                                   // perlin.rs:14-46
    std::vector<Vec3> lattice; std::vector<size_t> px, py, pz;
    explicit Perlin(uint64_t seed) {
        Rng rng = Rng::xorshift(seed);
        for (int i = 0; i < 256; i++) lattice.push_back(square_to_sphere(rng.gen_vec2()));
        px = gen_perm(rng); py = gen_perm(rng); pz = gen_perm(rng);
    }
    static std::vector<size_t> gen_perm(Rng& rng) {                                               // rng.rs:104-116
        const size_t n = 256;
        std::vector<size_t> perm(n);
        for (size_t i = 0; i < n; i++) perm[i] = i;
        for (size_t i = 0; i < n - 1; i++) { size_t rnd = (size_t)rng.gen_u64(); size_t j = i + (rnd % (n - i)); std::swap(perm[i], perm[j]); }
        return perm;
    }
    size_t hash(size_t x, size_t y, size_t z) const { return px[x % 256] ^ py[y % 256] ^ pz[z % 256]; }   // perlin.rs:71-75
    Float noise_at(Vec3 p) const {                                                                // perlin.rs:50-68
        Vec3 weight(fract(p.x), fract(p.y), fract(p.z));
        Vec3 fl = p.floor();
        Vec3 normals[8];
        int n = 0;
        for (size_t i = 0; i < 2; i++) for (size_t j = 0; j < 2; j++) for (size_t k = 0; k < 2; k++)
            normals[n++] = lattice[hash((size_t)sat_u64(fl.x) + i, (size_t)sat_u64(fl.y) + j, (size_t)sat_u64(fl.z) + k)];
        auto ss = [](Float x) { return ((6.0 * x - 15.0) * x + 10.0) * x * x * x; };              // _smootherstep, perlin.rs:83-85
        const Vec3 w(ss(weight.x), ss(weight.y), ss(weight.z));
        Float acc = 0.0;                                                                          // interp, perlin.rs:92-109
        n = 0;
        for (int xi = 0; xi < 2; xi++) for (int yi = 0; yi < 2; yi++) for (int zi = 0; zi < 2; zi++) {
            const Vec3 idx((Float)xi, (Float)yi, (Float)zi);
            const Vec3 widx = 2.0 * w * idx + Vec3(1, 1, 1) - w - idx;
            acc = acc + widx.x * widx.y * widx.z * normals[n++].dot(w - idx);
        }
        return acc;
    }
};
struct ImageSpectrum {                                                                            // image.rs:8-16, 99-199
    std::vector<Spectrum> buffer; uint32_t width = 0, height = 0;
    static uint32_t sat_u32(Float v) { uint64_t u = sat_u64(v); return u > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)u; }
    // bilin_interp (image.rs:99-131): texel indices and the two weights
    void taps(Vec2 uv, uint32_t& xo, uint32_t& yo, uint32_t& xi, uint32_t& yi, Float& wx, Float& wy) const {
        const Float w = (Float)width, h = (Float)height;
        const Vec2 xy(uv.x * w, (1.0 - uv.y) * h);
        const Vec2 xoyo(std::floor(xy.x - 0.5), std::floor(xy.y - 0.5));
        const Vec2 x1y1(xy.x - xoyo.x - 0.5, xy.y - xoyo.y - 0.5);
        wx = 1.0 - x1y1.x; wy = 1.0 - x1y1.y;
        xo = sat_u32(xoyo.x + w) % width; yo = sat_u32(xoyo.y + h) % height;
        xi = (xo + 1) % width; yi = (yo + 1) % height;
    }
    Color value_at(Vec2 uv, const Lambda& lam) const {                                            // image.rs:184-193
        uint32_t xo, yo, xi, yi; Float wx, wy;
        taps(uv, xo, yo, xi, yi, wx, wy);
        auto lerp = [&](const Spectrum& s0, const Spectrum& s1, Float v) { return s0.sample(lam) * v + s1.sample(lam) * (1.0 - v); };
        const Color y0 = lerp(buffer[xo + yo * width], buffer[xi + yo * width], wx);
        const Color y1 = lerp(buffer[xo + yi * width], buffer[xi + yi * width], wx);
        return y0 * wy + y1 * (1.0 - wy);
    }
};
struct ImageNormal {                                                                              // image.rs:134-181
    std::vector<Vec3> buffer; uint32_t width = 0, height = 0;
    Vec3 value_at(Vec2 uv) const {
        ImageSpectrum g; g.width = width; g.height = height;
        uint32_t xo, yo, xi, yi; Float wx, wy;
        g.taps(uv, xo, yo, xi, yi, wx, wy);
        auto lerp = [](Vec3 n0, Vec3 n1, Float v) { return (n0 * v + n1 * (1.0 - v)).normalize(); };
        const Vec3 y0 = lerp(buffer[xo + yo * width], buffer[xi + yo * width], wx);
        const Vec3 y1 = lerp(buffer[xo + yi * width], buffer[xi + yi * width], wx);
        return lerp(y0, y1, wy);
    }
};
enum TexKind { TEX_SOLID = 0, TEX_CHECKER = 1, TEX_MARBLE = 2, TEX_IMAGE = 3, TEX_MANDELBROT = 4, TEX_BUMP = 5 };
struct Texture {                                                                                  // texture.rs:23-38
    int kind = TEX_SOLID;
    Spectrum spec;                       // Solid / Marble colour; Image: mean
    const Texture *t1 = nullptr, *t2 = nullptr; Float scale = 1.0;
    std::unique_ptr<Perlin> pn;
    ImageSpectrum img; ImageNormal bump;
    static Float turbulence(const Perlin& pn, Float acc, Vec3 p, int depth) {                     // texture.rs:105-112
        if (depth >= 6) return acc;
        const Float w = powi(0.5, depth);
        return turbulence(pn, acc + w * std::fabs(pn.noise_at(p)), 2.0 * p, depth + 1);
    }
    Color albedo_at(const Lambda& lam, Vec2 uv) const {                                           // texture.rs:53-92
        switch (kind) {
        case TEX_MARBLE: {
            const Vec3 uvw(uv.x, uv.y, 0.0);
            const Float turb = turbulence(*pn, 0.0, 4.0 * uvw.abs(), 0);
            const Float scaled = 1.0 - powi(0.5 + 0.5 * lm_sin(60.0 * uvw.x + 20.0 * turb), 6);
            return spec.sample(lam) * scaled;
        }
        case TEX_CHECKER: {
            const Vec2 uvs = uv * scale;
            return (sat_u64(std::floor(uvs.x) + std::floor(uvs.y)) % 2 == 0) ? t1->albedo_at(lam, uv) : t2->albedo_at(lam, uv);
        }
        case TEX_IMAGE: return img.value_at(uv, lam);
        case TEX_MANDELBROT: {
            int depth = 0;
            const Float cre = 2.0 * (uv.x - 0.75), cim = 2.0 * (uv.y - 0.5);
            Float zre = 0.0, zim = 0.0;
            while (depth < 256 && zre * zre + zim * zim < 64.0 * 64.0) {
                const Float nre = zre * zre - zim * zim, nim = zre * zim + zim * zre;            // complex.rs Mul
                zre = nre + cre; zim = nim + cim;
                depth++;
            }
            return depth == 256 ? WHITE : BLACK;
        }
        default: return spec.sample(lam);
        }
    }
    Color power(const Lambda& lam) const { return spec.sample(lam); }                             // texture.rs:95-101 (Solid; Image: mean)
};

// ---- Materials (src/tracer/material.rs, bsdf.rs, bxdf.rs, microfacet.rs, bxdf/*.rs) -----------
enum MatKind { M_BLANK = 0, M_LAMBERTIAN = 1, M_MFDIFFUSE = 2, M_MFCONDUCTOR = 3, M_MFDIELECTRIC = 4, M_LIGHT = 5 };
struct Material {
    int kind = M_BLANK;
    Spectrum spec;                 // Lambertian
    Vec2 roughness;                // MicrofacetConfig (microfacet.rs:6-37); only Ggx is ever constructed (:51-60)
    DenseSpectrum eta, k;
    bool eta_const = true;
    Spectrum kd, ks, tf;           // Texture::Solid spectra, unless the matching *_tex below is set
    Spectrum ke; const double* illum = nullptr; Float scale = 1.0; bool two_sided = false;   // Light
    const Texture *kd_tex = nullptr, *ks_tex = nullptr, *tf_tex = nullptr, *ke_tex = nullptr, *bump_tex = nullptr;
    Color kd_at(const Lambda& lam, Vec2 uv) const { return kd_tex ? kd_tex->albedo_at(lam, uv) : kd.sample(lam); }   // microfacet.rs:120-134
    Color ks_at(const Lambda& lam, Vec2 uv) const { return ks_tex ? ks_tex->albedo_at(lam, uv) : ks.sample(lam); }
    Color tf_at(const Lambda& lam, Vec2 uv) const { return tf_tex ? tf_tex->albedo_at(lam, uv) : tf.sample(lam); }
    Vec3 map_normal(Vec3 ns, Vec2 uv) const {                                                 // material.rs:324-331
        if (!bump_tex) return ns;
        Onb onb(ns);
        return onb.to_world(bump_tex->bump.value_at(uv)).normalize();
    }

    bool mf_is_specular() const { return (roughness.x + roughness.y) / 2.0 < 0.01; }          // microfacet.rs:73-76
    bool mf_is_delta() const { return (roughness.x + roughness.y) / 2.0 < 1e-3; }             // microfacet.rs:80-83
    Float eta_at(Float wl) const { return eta.sample_one(wl); }
    Float k_at(Float wl) const { return k.sample_one(wl); }

    bool is_light() const { return kind == M_LIGHT; }
    bool bx_is_specular() const {                                                             // bxdf.rs:35-41
        if (kind == M_MFDIELECTRIC) return true;
        if (kind == M_MFCONDUCTOR) return mf_is_specular();
        return false;
    }
    bool is_specular() const { return (kind >= M_LAMBERTIAN && kind <= M_MFDIELECTRIC) ? bx_is_specular() : false; }  // material.rs:203-209
    bool bx_is_transmission() const { return kind == M_MFDIELECTRIC; }                        // bxdf.rs:48-53
    bool bx_is_reflection() const { return !bx_is_transmission(); }
    bool is_delta(const Lambda& lam) const {                                                  // material.rs:212-217, bxdf.rs:59-67
        if (kind == M_MFCONDUCTOR) return mf_is_delta();
        if (kind == M_MFDIELECTRIC) return mf_is_delta() || eta_at(lam.leading()) == 1.0;
        return false;
    }
    Color emit(const Lambda& lam, const Hit& h) const {                                       // material.rs:220-231
        if (kind != M_LIGHT) return BLACK;
        if (!two_sided && h.backface) return BLACK;
        return scale * (ke_tex ? ke_tex->albedo_at(lam, h.uv) : ke.sample(lam)) * DenseSpectrum::sample_raw(illum, lam);
    }
    Color power(const Lambda& lam) const {                                                    // material.rs:234-242
        if (kind != M_LIGHT) return BLACK;
        Color phi = scale * (ke_tex ? ke_tex->power(lam) : ke.sample(lam)) * DenseSpectrum::sample_raw(illum, lam);
        return two_sided ? 2.0 * phi : phi;
    }
    Float shading_cosine(Vec3 wi, Vec3 ns) const {                                            // material.rs:315-321
        return (kind >= M_LAMBERTIAN && kind <= M_MFDIELECTRIC) ? std::fabs(ns.dot(wi)) : 1.0;
    }

    // ---- MfDistribution (microfacet.rs) ----
    Float f_schlick(Float f0, Float f90, Float cos_theta) const { return f0 + (f90 - f0) * powi(1.0 - cos_theta, 5); }  // :198-200
    Float disney_diffuse(Float cwo, Float cwi, Float cwh) const {                             // :129-145
        Float r2 = powi(roughness.x, 2);
        Float energy_bias = 0.5 * r2;
        Float fd90 = energy_bias + 2.0 * powi(cwh, 2) * r2;
        Float vs = f_schlick(1.0, fd90, cwo), ls = f_schlick(1.0, fd90, cwi);
        Float ef = 1.0 + r2 * (1.0 / 1.51 - 1.0);
        return vs * ls * ef;
    }
    Float d(Vec3 wh) const {                                                                  // :147-195 (Ggx)
        Float tan2 = sph::tan2_theta(wh);
        if (std::isinf(tan2)) return 0.0;
        Float cos4 = powi(sph::cos2_theta(wh), 2);
        if (cos4 < powi(EPSILON, 2)) return 0.0;
        Float cp = sph::cos_phi(wh), sp = sph::sin_phi(wh);
        Float alpha2 = roughness.x * roughness.y;
        Float e = tan2 * (powi(cp / roughness.x, 2) + powi(sp / roughness.y, 2));
        return 1.0 / (PI * alpha2 * cos4 * powi(1.0 + e, 2));
    }
    Float fr_complex(Vec3 wo, Vec3 wh, Float eta_, Float k_) const {                          // :226-241
        Complex eta(eta_, k_);
        Float cos_o = clampf(wo.dot(wh), 0.0, 1.0);
        Float sin2_o = 1.0 - cos_o * cos_o;
        Complex sin2_i = sin2_o / (eta * eta);
        Complex cos_i = (1.0 - sin2_i).sqrt();
        Complex r_par = (eta * cos_o - cos_i) / (eta * cos_o + cos_i);
        Complex r_per = (cos_o - eta * cos_i) / (cos_o + eta * cos_i);
        return (r_par.norm_sqr() + r_per.norm_sqr()) / 2.0;
    }
    Float fr_real(Vec3 wo, Vec3 wh, Float eta_) const {                                       // :244-265
        Float cos_o = wo.dot(wh);
        bool inside = cos_o < 0.0;
        Float eta = inside ? 1.0 / eta_ : eta_;
        cos_o = std::fabs(cos_o);
        Float sin2_o = 1.0 - cos_o * cos_o;
        Float sin2_i = sin2_o / (eta * eta);
        if (sin2_i >= 1.0) return 1.0;
        Float cos_i = std::sqrt(fmax_(1.0 - sin2_i, 0.0));
        Float r_par = (eta * cos_o - cos_i) / (eta * cos_o + cos_i);
        Float r_per = (cos_o - eta * cos_i) / (cos_o + eta * cos_i);
        return (r_par * r_par + r_per * r_per) / 2.0;
    }
    Float f_at(Vec3 wo, Vec3 wh, Float wl) const {                                            // :210-223
        Float e = eta_at(wl), kk = k_at(wl);
        if (kk == 0.0) return e == 0.0 ? 0.0 : fr_real(wo, wh, e);
        return fr_complex(wo, wh, e, kk);
    }
    Color f_col(Vec3 wo, Vec3 wh, const Lambda& lam) const { Color c; for (int i = 0; i < 4; i++) c.s[i] = f_at(wo, wh, lam.l[i]); return c; }
    bool chi_pass(Vec3 wo, Vec3 wh) const {                                                   // :268-274
        Float chi = signum(sph::cos_theta(wh)) * wo.dot(wh) * sph::cos_theta(wo);
        return chi > EPSILON;
    }
    Float lambda_(Vec3 w) const {                                                             // :296-327 (Ggx)
        Float tan2 = sph::tan2_theta(w);
        if (std::isinf(tan2)) return 0.0;
        Float cp = sph::cos_phi(w), sp = sph::sin_phi(w);
        Float alpha2 = powi(roughness.x * cp, 2) + powi(roughness.y * sp, 2);
        return (std::sqrt(fmax_(1.0 + alpha2 * tan2, 0.0)) - 1.0) / 2.0;
    }
    Float g(Vec3 wo, Vec3 wi, Vec3 wh) const { return !chi_pass(wo, wh) ? 0.0 : 1.0 / (1.0 + lambda_(wo) + lambda_(wi)); }  // :277-283
    Float g1(Vec3 wo, Vec3 wh) const { return !chi_pass(wo, wh) ? 0.0 : 1.0 / (1.0 + lambda_(wo)); }                        // :286-292
    Float sample_normal_pdf(Vec3 wh, Vec3 wo) const {                                         // :330-349 (Ggx)
        Float pdf = g1(wo, wh) * d(wh) * std::fabs(wh.dot(wo)) / std::fabs(sph::cos_theta(wo));
        return fmax_(pdf, 0.0);
    }
    Vec3 sample_normal(Vec3 wo, Vec2 rs) const {                                              // :352-446 (Ggx, Heitz 2018)
        Vec3 ws = Vec3(wo.x * roughness.x, wo.y * roughness.y, wo.z).normalize();
        if (ws.z < 0.0) ws = -ws;
        Vec3 u = (1.0 - ws.z < EPSILON) ? Vec3(1, 0, 0) : ws.cross(Vec3(0, 0, 1)).normalize();
        Vec3 v = u.cross(ws);
        Float r = std::sqrt(rs.x);
        Float theta = 2.0 * PI * rs.y;
        Float x = r * lm_cos(theta);
        Float h = std::sqrt(fmax_(1.0 - x * x, 0.0));
        Float lerp = (1.0 + ws.z) / 2.0;
        Float y = (1.0 - lerp) * h + lerp * r * lm_sin(theta);
        Vec3 wm(x, y, std::sqrt(fmax_(1.0 - x * x - y * y, 0.0)));
        wm = wm.x * u + wm.y * v + wm.z * ws;
        return Vec3(roughness.x * wm.x, roughness.y * wm.y, fmax_(wm.z, EPSILON)).normalize();
    }

    // ---- bxdf/microfacet.rs util ----
    static bool reflect(Vec3 wo, Vec3 no, Vec3& wi) {                                         // :7-15
        Vec3 proj = no * wo.dot(no) / no.length_squared();   // Vec3::project_onto (vec3.rs:142-144)
        wi = 2.0 * proj - wo;
        return sph::same_hemisphere(wi, wo);
    }
    static bool refract(Float eta, Vec3 wo, Vec3 no, Vec3& wi) {                              // :17-44
        Float cos_to, eta_ratio; Vec3 n;
        if (no.dot(wo) < 0.0) { cos_to = -no.dot(wo); eta_ratio = 1.0 / eta; n = -no; }
        else { cos_to = no.dot(wo); eta_ratio = eta; n = no; }
        Float sin2_to = 1.0 - cos_to * cos_to;
        Float sin2_ti = sin2_to / powi(eta_ratio, 2);
        if (sin2_ti >= 1.0) { assert(false && "unreachable (bxdf/microfacet.rs:32)"); return false; }
        Float cos_ti = std::sqrt(fmax_(1.0 - sin2_ti, 0.0));
        wi = -wo / eta_ratio + (cos_to / eta_ratio - cos_ti) * n;
        return !sph::same_hemisphere(wi, wo);
    }
    Color reflect_coeff(Vec3 wo, Vec3 wi, const Lambda& lam) const {                          // :46-62
        Float cwo = sph::cos_theta(wo), cwi = sph::cos_theta(wi);
        Vec3 wh = (wi + wo).normalize();
        Float dd = d(wh); Color f = f_col(wo, wh, lam); Float gg = g(wo, wi, wh);
        return dd * f * gg / (4.0 * std::fabs(cwo) * std::fabs(cwi));
    }
    static Float lambertian_pdf(Vec3 wo, Vec3 wi) {                                           // bxdf/scatter.rs:14-25
        if (!sph::same_hemisphere(wo, wi)) return 0.0;
        Float c = sph::cos_theta(wi);
        return c > 0.0 ? c / PI : 0.0;
    }

    // ---- BxDF::f / sample / pdf in the local frame (bxdf.rs:69-151) ----
    Color bx_f(Vec3 wo, Vec3 wi, const Lambda& lam, bool reflection, bool backface, int mode, Vec2 uv = Vec2()) const {
        if ((!reflection || backface) && bx_is_reflection()) return BLACK;
        switch (kind) {
        case M_LAMBERTIAN: return spec.sample(lam) / PI;                                      // scatter.rs:6-8
        case M_MFDIFFUSE: {                                                                   // microfacet.rs:131-156
            Vec3 wh = (wo + wi).normalize();
            Float cwo = sph::cos_theta(wo), cwi = sph::cos_theta(wi), cwh = sph::cos_theta(wh);
            Float dd = d(wh); Color f = f_col(wo, wh, lam); Float gg = g(wo, wi, wh);
            Color fr = dd * f * gg / (4.0 * std::fabs(cwo) * std::fabs(cwi));
            Float fd = disney_diffuse(cwo, cwi, cwh);
            Color ksc = ks_at(lam, uv), kdc = kd_at(lam, uv);
            return fr * ksc + kdc * (WHITE - f) * fd / PI;
        }
        case M_MFCONDUCTOR: {                                                                 // microfacet.rs:71-85
            Color ksc = ks_at(lam, uv);
            if (mf_is_delta()) { Color f = f_col(wo, Vec3(0, 0, 1), lam); return ksc * f / std::fabs(sph::cos_theta(wi)); }
            return ksc * reflect_coeff(wo, wi, lam);
        }
        case M_MFDIELECTRIC: {                                                                // microfacet.rs:222-283
            Float cwo = sph::cos_theta(wo), cwi = sph::cos_theta(wi);
            bool wo_inside = cwo < 0.0;
            Float wl = lam.leading();
            Float e = eta_at(wl);
            Float eta_ratio = reflection ? 1.0 : (wo_inside ? 1.0 / e : e);
            Vec3 wh = (e == 1.0 || mf_is_delta()) ? Vec3(0, 0, 1) : (wi * eta_ratio + wo).normalize();
            if (reflection) {
                Color ksc = ks_at(lam, uv);
                if (e == 1.0 || mf_is_delta()) { Color f = f_col(wo, wh, lam); return ksc * f / std::fabs(cwi); }
                return ksc * reflect_coeff(wo, wi, lam);
            }
            Color f = f_col(wo, wh, lam);
            if (sph::cos_theta(wh) < 0.0) wh = -wh;
            Float scale_ = mode == RADIANCE ? eta_ratio * eta_ratio : 1.0;
            Color tfc = tf_at(lam, uv);
            if (e == 1.0 || mf_is_delta()) return tfc * (WHITE - f) / (scale_ * std::fabs(cwi));
            Float dd = d(wh), gg = g(wo, wi, wh);
            Float wh_dot_wo = wh.dot(wo), wh_dot_wi = wh.dot(wi);
            return tfc * dd * (WHITE - f) * gg / scale_
                * std::fabs(wh_dot_wi * wh_dot_wo / (cwi * cwo))
                / powi(eta_ratio * wh_dot_wi + wh_dot_wo, 2);
        }
        default: return BLACK;
        }
    }
    bool bx_sample(Vec3 wo, bool backface, Lambda& lam, Float rand_u, Vec2 rs, Vec3& wi) const {
        if (backface && bx_is_reflection()) return false;
        switch (kind) {
        case M_LAMBERTIAN: wi = square_to_cos_hemisphere(rs); return true;
        case M_MFDIFFUSE: {                                                                   // microfacet.rs:159-177
            Float pr = f_schlick(0.04, 1.0, sph::cos_theta(wo)), ps = 1.0 - pr;
            if (rand_u < pr / (pr + ps)) {
                Vec3 wh = mf_is_delta() ? Vec3(0, 0, 1) : sample_normal(wo, rs);
                return reflect(wo, wh, wi);
            }
            wi = square_to_cos_hemisphere(rs); return true;
        }
        case M_MFCONDUCTOR: {                                                                 // microfacet.rs:87-99
            if (mf_is_delta()) { wi = Vec3(-wo.x, -wo.y, wo.z); return true; }
            Vec3 wh = sample_normal(wo, rs);
            return reflect(wo, wh, wi);
        }
        case M_MFDIELECTRIC: {                                                                // microfacet.rs:285-311
            Float wl = eta_const ? lam.leading() : lam.terminate();
            Float e = eta_at(wl);
            Vec3 wh = (e == 1.0 || mf_is_delta()) ? Vec3(0, 0, 1) : sample_normal(wo, rs);
            Float pr = f_at(wo, wh, wl), pt = 1.0 - pr;
            if (rand_u < pr / (pr + pt)) return reflect(wo, wh, wi);
            return refract(e, wo, wh, wi);
        }
        default: return false;
        }
    }
    Float refl_pdf_half(Vec3 wo, Vec3 wh) const {   // shared tail of conductor/diffuse pdf
        if (mf_is_delta()) return (1.0 - sph::cos_theta(wh) < EPSILON) ? 1.0 : 0.0;
        Float wh_dot_wo = wo.dot(wh);
        return sample_normal_pdf(wh, wo) / (4.0 * std::fabs(wh_dot_wo));
    }
    Float bx_pdf(Vec3 wo, Vec3 wi, bool reflection, const Lambda& lam) const {
        if (!reflection && bx_is_reflection()) return 0.0;
        switch (kind) {
        case M_LAMBERTIAN: return lambertian_pdf(wo, wi);
        case M_MFDIFFUSE: {                                                                   // microfacet.rs:180-205
            if (!sph::same_hemisphere(wi, wo)) return 0.0;
            Vec3 wh = (wo + wi).normalize();
            Float pr = f_schlick(0.04, 1.0, sph::cos_theta(wo)), ps = 1.0 - pr;
            Float p_ref = refl_pdf_half(wo, wh);
            Float p_sct = lambertian_pdf(wo, wi);
            return pr * p_ref + ps * p_sct;
        }
        case M_MFCONDUCTOR: {                                                                 // microfacet.rs:101-120
            if (!sph::same_hemisphere(wi, wo)) return 0.0;
            Vec3 wh = (wo + wi).normalize();
            if (sph::cos_theta(wh) < 0.0) wh = -wh;
            return refl_pdf_half(wo, wh);
        }
        case M_MFDIELECTRIC: {                                                                // microfacet.rs:313-373
            Float cwo = sph::cos_theta(wo), cwi = sph::cos_theta(wi);
            bool wo_inside = cwo < 0.0;
            Float wl = lam.leading();
            Float e = eta_at(wl);
            Float eta_ratio = reflection ? 1.0 : (wo_inside ? 1.0 / e : e);
            Vec3 wh = e == 1.0 ? Vec3(0, 0, 1) : (wo + wi * eta_ratio).normalize();
            if (sph::cos_theta(wh) < 0.0) wh = -wh;
            Float wh_dot_wo = wo.dot(wh), wh_dot_wi = wi.dot(wh);
            if (wh_dot_wo == 0.0 || wh_dot_wi == 0.0) return 0.0;
            if (wh_dot_wo * cwo < 0.0 || wh_dot_wi * cwi < 0.0) return 0.0;
            Float pr = f_at(wo, wh, wl), pt = 1.0 - pr;
            if (reflection && (e == 1.0 || mf_is_delta()))
                return (1.0 - sph::cos_theta(wh) < EPSILON) ? pr / (pr + pt) : 0.0;
            if (reflection) return sample_normal_pdf(wh, wo) / (4.0 * std::fabs(wh_dot_wo)) * pr / (pr + pt);
            if (e == 1.0 || mf_is_delta())
                return (1.0 - sph::cos_theta(wh) < EPSILON) ? pt / (pr + pt) : 0.0;
            return sample_normal_pdf(wh, wo) * std::fabs(wh_dot_wi) / powi(wh_dot_wi + wh_dot_wo / eta_ratio, 2) * pt / (pr + pt);
        }
        default: return 0.0;
        }
    }

    // ---- BSDF world<->local wrapper (bsdf.rs:28-90) + Material dispatch (material.rs:245-312) ----
    static bool is_reflection(Vec3 wo, Vec3 wi, Vec3 ng) { return ng.dot(wi) * ng.dot(wo) >= 0.0; }
    bool is_standard() const { return kind >= M_LAMBERTIAN && kind <= M_MFDIELECTRIC; }
    Color bsdf_f(Vec3 wo, Vec3 wi, const Lambda& lam, int mode, const Hit& h) const {
        if (!is_standard()) return BLACK;
        bool refl = is_reflection(wo, wi, h.ng);
        Onb uvw(map_normal(h.ns, h.uv));
        return bx_f(uvw.to_local(wo), uvw.to_local(wi), lam, refl, h.backface, mode, h.uv);
    }
    bool bsdf_sample(Vec3 wo, const Hit& h, Lambda& lam, Float rand_u, Vec2 rs, Vec3& wi) const {
        if (!is_standard()) return false;
        Onb uvw(map_normal(h.ns, h.uv));
        Vec3 wl;
        if (!bx_sample(uvw.to_local(wo), h.backface, lam, rand_u, rs, wl)) return false;
        wi = uvw.to_world(wl);
        return true;
    }
    Float bsdf_pdf(Vec3 wo, Vec3 wi, const Hit& h, const Lambda& lam, bool swap_dir) const {
        if (swap_dir) std::swap(wo, wi);
        if (!is_standard()) return 0.0;
        bool refl = is_reflection(wo, wi, h.ng);
        Onb uvw(map_normal(h.ns, h.uv));
        return bx_pdf(uvw.to_local(wo), uvw.to_local(wi), refl, lam);
    }
};

// ---- colour spaces / white balance (src/tracer/color/space.rs, xyz.rs) -------------------------
static inline Vec3 from_xyY(Vec2 xy, Float Y) {                                                // xyz.rs:8-18
    if (xy.y == 0.0) return Vec3();
    return Vec3(xy.x * Y / xy.y, Y, (1.0 - xy.x - xy.y) * Y / xy.y);
}
static inline Vec2 to_xyY(Vec3 xyz) { return Vec2(xyz.x / (xyz.x + xyz.y + xyz.z), xyz.y / (xyz.x + xyz.y + xyz.z)); }
struct ColorSpace {
    Mat3 XYZ_to_RGB; Vec3 W; int trc;   // trc 0 = sRGB, 1 = rec_2020
    static Mat3 xyz_to_rgb(Vec2 r, Vec2 g, Vec2 b, Vec3 W) {                                   // space.rs:162-177
        Vec3 R = from_xyY(r, 1.0), G = from_xyY(g, 1.0), B = from_xyY(b, 1.0);
        Mat3 RGB_c = Mat3(R, G, B).transpose();
        Vec3 C = RGB_c.inv().mul_vec3(W);
        Mat3 RGB_to_XYZ = RGB_c.mul_mat3(Mat3::diag(C));
        return RGB_to_XYZ.inv();
    }
    static ColorSpace get(int id) {                                                            // space.rs:50-115
        Vec3 W = from_xyY(to_xyY(DenseSpectrum::from(spectra::D65).to_xyz()), 1.0);
        ColorSpace cs; cs.W = W; cs.trc = 0;
        if (id == 0) cs.XYZ_to_RGB = xyz_to_rgb(Vec2(0.64, 0.33), Vec2(0.3, 0.6), Vec2(0.15, 0.06), W);
        else if (id == 1) cs.XYZ_to_RGB = xyz_to_rgb(Vec2(0.68, 0.32), Vec2(0.265, 0.69), Vec2(0.15, 0.06), W);
        else { cs.XYZ_to_RGB = xyz_to_rgb(Vec2(0.708, 0.292), Vec2(0.170, 0.797), Vec2(0.131, 0.046), W); cs.trc = 1; }
        return cs;
    }
    Mat3 wb_matrix(const double* illuminant) const {                                           // space.rs:144-151
        const Mat3 XYZ_to_LMS(Vec3(0.210576, 0.855098, -0.0396983), Vec3(-0.417076, 1.177260, 0.0786283), Vec3(0.0, 0.0, 0.5168350));
        const Mat3 LMS_to_XYZ = XYZ_to_LMS.inv();
        Vec2 illum_xy = to_xyY(DenseSpectrum::from(illuminant).to_xyz());
        Vec3 diagonal = XYZ_to_LMS.mul_vec3(W) / XYZ_to_LMS.mul_vec3(from_xyY(illum_xy, 1.0));
        return LMS_to_XYZ.mul_mat3(Mat3::diag(diagonal)).mul_mat3(XYZ_to_LMS);
    }
};

// ---- pixel filter (src/tracer/filter.rs) -------------------------------------------------------
struct PixelFilter {
    int kind = 2; Float r = 1.5, p = 0.375;   // Gaussian(1.5, 1.5/4) default (:20-24)
    uint64_t r_disc() const { return sat_u64(std::ceil(r - 0.5)); }                            // :70-79
    static Float gauss(Float x, Float sigma) {                                                 // :118-123
        return lm_exp(-powi(x, 2) / (2.0 * sigma * sigma)) / std::sqrt(fmax_(2.0 * PI * sigma * sigma, 0.0));
    }
    static Float mitch(Float x, Float b, Float c) {                                            // :132-150
        x = std::fabs(x);
        Float q = 0.0;
        if (x < 1.0) q = (12.0 - 9.0 * b - 6.0 * c) * powi(x, 3) + (-18.0 + 12.0 * b + 6.0 * c) * powi(x, 2) + (6.0 - 2.0 * b);
        else if (x < 2.0) q = (-b - 6.0 * c) * powi(x, 3) + (6.0 * b + 30.0 * c) * powi(x, 2) + (-12.0 * b - 48.0 * c) * x + (8.0 * b + 24.0 * c);
        return q / 6.0;
    }
    Float eval(Vec2 px) const {                                                                // :82-102
        switch (kind) {
        case 0: return (std::fabs(px.x) < r && std::fabs(px.y) < r) ? 1.0 : 0.0;
        case 1: { Float ox = fmax_(r - std::fabs(px.x), 0.0), oy = fmax_(r - std::fabs(px.y), 0.0); return ox * oy; }
        case 2: { Float gx = gauss(px.x, p), gy = gauss(px.y, p), gr = gauss(r, p); return fmax_(gx - gr, 0.0) * fmax_(gy - gr, 0.0); }
        default: { Float c = (1.0 - p) / 2.0; return mitch(2.0 * px.x / r, p, c) * mitch(2.0 * px.y / r, p, c); }
        }
    }
    Float integral() const {                                                                   // :105-116
        switch (kind) {
        case 0: return 2.0 * r * 2.0 * r;
        case 1: return r * r * r * r;
        case 3: return r * r * 0.25;
        default: {
            Float denom = p * std::sqrt(2.0);
            Float ig = 0.5 * (std::erf(-(-r) / denom) - std::erf(-r / denom));   // libm::erf in the reference (:126-129)
            Float gr = gauss(r, p);
            return powi(ig - 2.0 * r * gr, 2);
        }
        }
    }
};

// ---- camera (src/tracer/camera.rs, camera/builder.rs, camera/matrices.rs) ----------------------
struct Camera {
    bool ortho = false;
    uint64_t res_x = 1024, res_y = 768;
    Float focal_length = 0, lens_radius = 0, image_plane_area = 0;
    Transform camera_to_screen, screen_to_raster, world_to_camera;
    PixelFilter filter; int color_space = 1; int illuminant = 2;

    static Camera build(Vec3 origin, Vec3 towards, Vec3 up, Float zoom, Float lens_radius, Float focal_length, Float vfov,
                        uint64_t w, uint64_t h, bool ortho, PixelFilter filt, int cs, int illum) {  // builder.rs:124-150
        Camera c; c.ortho = ortho; c.res_x = w; c.res_y = h; c.focal_length = focal_length; c.lens_radius = lens_radius;
        c.filter = filt; c.color_space = cs; c.illuminant = illum;
        if (!ortho) {                                                                          // matrices.rs:4-14
            Float near = 1e-2, far = 1e3;
            Transform proj = Transform::perspective(near, far);
            Float tvi = 1.0 / std::tan((vfov * (PI / 180.0)) / 2.0);   // f64::to_radians = x * (PI/180)
            c.camera_to_screen = Transform::scale(tvi, tvi, 1.0).mul(proj);
        } else {
            c.camera_to_screen = Transform::scale(1.0, 1.0, 1.0 / (1.0 - 0.0)).mul(Transform::translation(0.0, 0.0, -0.0));
        }
        {                                                                                      // matrices.rs:24-37
            Vec3 forward = (towards - origin).normalize();
            Vec3 right = forward.cross(up).normalize();
            Vec3 up2 = right.cross(forward);
            c.world_to_camera = Transform::translation(-origin.dot(right), -origin.dot(up2), -origin.dot(forward))
                .mul(Transform::mat3(Mat3(right, up2, forward)));
        }
        {                                                                                      // matrices.rs:39-70
            Float ar = (Float)w / (Float)h;
            Vec2 smin, smax;
            if (ar > 1.0) { smin = Vec2(-ar, -1.0); smax = Vec2(ar, 1.0); }
            else { smin = Vec2(-1.0, -1.0 / ar); smax = Vec2(1.0, 1.0 / ar); }
            Vec2 sd = smax - smin;
            c.screen_to_raster = Transform::scale((Float)w, -((Float)h), 1.0)
                .mul(Transform::scale(1.0 / sd.x, 1.0 / sd.y, 1.0))
                .mul(Transform::translation(-smin.x, -smax.y, 0.0))
                .mul(Transform::scale(zoom, zoom, zoom));
        }
        {                                                                                      // camera.rs:51-67
            Vec3 p_min = c.screen_to_raster.transform_pt_inv(Vec3());
            Vec3 p_max = c.screen_to_raster.transform_pt_inv(Vec3((Float)w, (Float)h, 0.0));
            p_min = c.camera_to_screen.transform_pt_inv(p_min);
            p_max = c.camera_to_screen.transform_pt_inv(p_max);
            Float zmin = p_min.z == 0.0 ? 1.0 : p_min.z, zmax = p_max.z == 0.0 ? 1.0 : p_max.z;
            Vec2 a(p_min.x / zmin, p_min.y / zmin), b(p_max.x / zmax, p_max.y / zmax);
            Vec2 dlt = b - a;
            c.image_plane_area = std::fabs(dlt.x * dlt.y);
        }
        return c;
    }
    Vec3 raster_to_camera(Vec2 xy) const {                                                     // camera.rs:111-115
        Vec3 s = screen_to_raster.transform_pt_inv(Vec3(xy.x, xy.y, 0.0));
        return camera_to_screen.transform_pt_inv(s);
    }
    Vec2 camera_to_raster(Vec3 xl) const {                                                     // camera.rs:117-121
        Vec3 s = camera_to_screen.transform_pt(xl);
        Vec3 r = screen_to_raster.transform_pt(s);
        return Vec2(r.x, r.y);
    }
    Ray add_dof(Vec3 xo_local, Vec3 wi_local, Vec2 rs) const {                                 // camera.rs:221-243
        if (lens_radius != 0.0) {
            Vec2 l = lens_radius * square_to_disk(rs);
            Vec3 lens(l.x, l.y, 0.0);
            Float fd = focal_length / wi_local.z;
            Vec3 focus = fd * wi_local;
            xo_local = xo_local + lens; wi_local = focus - lens;
        }
        Vec3 xo = world_to_camera.transform_pt_inv(xo_local);
        Vec3 wi = world_to_camera.transform_dir_inv(wi_local);
        return Ray::make(xo, wi);
    }
    Ray generate_ray(Vec2 raster_xy, Vec2 rs) const {                                          // camera.rs:257-268
        if (!ortho) return add_dof(Vec3(), raster_to_camera(raster_xy).normalize(), rs);
        return add_dof(raster_to_camera(raster_xy), Vec3(0, 0, 1), rs);
    }
    bool bounds(Vec2 r) const { return r.x >= 0.0 && r.x < (Float)res_x && r.y >= 0.0 && r.y < (Float)res_y; }
    bool raster_xy(const Ray& ri, Vec2& out) const {                                           // camera.rs:167-214
        if (ortho) {
            Vec3 xl = world_to_camera.transform_pt(ri.origin);
            out = camera_to_raster(xl);
            return bounds(out);
        }
        Vec3 wl = world_to_camera.transform_dir(ri.dir);
        Float ct = wl.z;
        if (ct <= 0.0) return false;
        Float fl = lens_radius == 0.0 ? 1.0 / ct : focal_length / ct;
        Vec3 xl = world_to_camera.transform_pt(ri.origin);
        out = camera_to_raster(xl + wl * fl);
        return bounds(out);
    }
    Float lens_area() const { return lens_radius == 0.0 ? 1.0 : PI * powi(lens_radius, 2); }   // camera.rs:246-253
    bool sample_towards(Vec3 xi, Vec2 rs, Ray& ri) const {                                     // camera.rs:271-297
        if (ortho) {
            Vec2 dk = square_to_disk(rs);
            Vec3 lens = lens_radius * Vec3(dk.x, dk.y, 0.0);
            Vec3 xil = world_to_camera.transform_pt(xi);
            Vec3 xol = xil * Vec3(1.0, 1.0, 0.0);
            Vec3 xo = world_to_camera.transform_pt_inv(xol + lens);
            ri = Ray::make(xo, (xi - xo).normalize());
        } else {
            Vec2 dk = square_to_disk(rs);
            Vec3 xol = lens_radius * Vec3(dk.x, dk.y, 0.0);
            Vec3 xil = world_to_camera.transform_pt(xi);
            Vec3 wil = (xil - xol).normalize();
            ri = Ray::make(world_to_camera.transform_pt_inv(xol), world_to_camera.transform_dir_inv(wil));
        }
        Vec2 dummy;
        return raster_xy(ri, dummy);
    }
    Float pdf_xo(const Ray& ri) const {                                                        // camera.rs:300-321
        Vec2 d;
        if (ortho) return raster_xy(ri, d) ? 1.0 / image_plane_area : 0.0;
        Vec3 xl = world_to_camera.transform_pt(ri.origin);
        Float r2 = powi(lens_radius + EPSILON, 2);
        return xl.distance_squared(Vec3()) < r2 ? 1.0 / lens_area() : 0.0;
    }
    Float pdf_wi(const Ray& ri) const {                                                        // camera.rs:324-352
        Vec2 d;
        Vec3 wl = world_to_camera.transform_dir(ri.dir);
        if (ortho) return (1.0 - wl.z < EPSILON) ? 1.0 : 0.0;
        if (!raster_xy(ri, d)) return 0.0;
        return 1.0 / (image_plane_area * powi(wl.z, 3));
    }
    Float pdf_importance(const Ray& ri, Vec3 xi) const {                                       // camera.rs:355-370 (perspective only)
        Vec2 d;
        if (!raster_xy(ri, d)) return 0.0;
        Vec3 ng = world_to_camera.to_normal_inv().mul_vec3(Vec3(0, 0, 1));
        Float pdf = xi.distance_squared(ri.origin) / (std::fabs(ng.dot(ri.dir)) * lens_area());
        return fmax_(pdf, 0.0);
    }
    bool sample_importance(const Ray& ri, Color& imp, Vec2& raster) const {                    // camera.rs:373-388
        if (!raster_xy(ri, raster)) return false;
        if (ortho) { imp = (1.0 / image_plane_area) * WHITE; return true; }
        Vec3 wl = world_to_camera.transform_dir(ri.dir);
        Float denom = image_plane_area * powi(wl.z, 4) * lens_area();
        imp = (1.0 / denom) * WHITE;
        return true;
    }
};

// ---- film (src/tracer/film.rs, film/tile.rs, tone_mapping.rs) ----------------------------------
struct FilmSample { Vec2 raster_xy; Color color; Lambda lambda; bool splat; size_t cost; };
struct Pixel { Vec3 color; Float w = 0; };
struct TileSplat { Vec3 color; uint64_t x, y; };
static inline uint64_t ssub(uint64_t a, uint64_t b) { return a > b ? a - b : 0; }   // UVec2 Sub = saturating_sub (vec2.rs:162)

struct FilmTile {
    uint64_t px_min_x, px_min_y, px_max_x, px_max_y, width, height, res_x, res_y;
    std::vector<Pixel> pixels; std::vector<TileSplat> splats;
    const ColorSpace* cs; Mat3 wb; const PixelFilter* filter;
    FilmTile(uint64_t x0, uint64_t y0, uint64_t x1, uint64_t y1, uint64_t rx, uint64_t ry, const ColorSpace* cs_, const Mat3& wb_, const PixelFilter* f)
        : px_min_x(x0), px_min_y(y0), px_max_x(x1), px_max_y(y1), res_x(rx), res_y(ry), cs(cs_), wb(wb_), filter(f) {  // tile.rs:34-63
        uint64_t radius = f->r_disc();
        uint64_t fmin_x = ssub(x0, radius), fmin_y = ssub(y0, radius);
        uint64_t fmax_x = std::min(x1 + radius, rx), fmax_y = std::min(y1 + radius, ry);
        width = fmax_x - fmin_x; height = fmax_y - fmin_y;
        pixels.assign(width * height, Pixel());
    }
    void add_sample(const FilmSample& s) {                                                     // tile.rs:65-111
        Vec3 rgb = cs->XYZ_to_RGB.mul_vec3(wb.mul_vec3(color_xyz(s.color, s.lambda)));
        uint64_t px = sat_u64(std::floor(s.raster_xy.x)), py = sat_u64(std::floor(s.raster_xy.y));
        uint64_t r = filter->r_disc();
        uint64_t mi_x, mi_y, mx_x, mx_y;
        if (s.splat) { mi_x = ssub(px, r); mi_y = ssub(py, r); mx_x = std::min(px + r, res_x - 1); mx_y = std::min(py + r, res_y - 1); }
        else { mi_x = std::max(ssub(px, r), px_min_x); mi_y = std::max(ssub(py, r), px_min_y);
               mx_x = std::min(px + r, px_max_x - 1); mx_y = std::min(py + r, px_max_y - 1); }
        for (uint64_t fy = mi_y; fy <= mx_y; fy++) for (uint64_t fx = mi_x; fx <= mx_x; fx++) {
            Vec2 mid = 0.5 + Vec2((Float)fx, (Float)fy);
            Vec2 v = s.raster_xy - mid;
            Float w = filter->eval(v);
            if (w != 0.0) {
                if (s.splat) splats.push_back(TileSplat{rgb * w, fx, fy});
                else {
                    uint64_t idx = (fx - px_min_x) + width * (fy - px_min_y);
                    pixels[idx].color = pixels[idx].color + rgb * w;
                    pixels[idx].w += w;
                }
            }
        }
    }
};
static inline Color tone_map(int kind, Float arg, const FilmSample& s) {                        // tone_mapping.rs:38-63
    if (kind == 1) { Color c; for (int i = 0; i < 4; i++) c.s[i] = clampf(s.color.s[i], 0.0, arg); return c; }
    if (kind == 2) return s.color / (1.0 + color_luminance(s.color, s.lambda));
    return s.color;
}

// ---- Scene (src/tracer/scene.rs) ---------------------------------------------------------------
struct Scene {
    BVH objects, lights;
    const Material* env = nullptr;
    AABB bounds;
    std::vector<std::unique_ptr<Material>> materials;
    std::vector<std::unique_ptr<TriangleMesh>> meshes;

    void build() {                                                                             // scene.rs:33-52
        objects.build();
        bounds = objects.bounding_box().merge(lights.bounding_box());
        if (env) {
            Vec3 center = bounds.center();
            Float radius = center.distance(bounds.ax_min);
            auto inst = std::make_shared<Instance>(std::make_shared<Sphere>(radius, env));
            inst->apply(Transform::translation(center.x, center.y, center.z));
            bounds = bounds.merge(inst->bounding_box());
            lights.add(inst);
            env = nullptr;
        }
        lights.build();
        build_alias();
    }
    void build_alias() {                                                                       // bvh.rs:105-166
        Lambda lam = Lambda::sample(0.0);
        Float sum = 0.0; size_t n = lights.objects.size();
        auto& at = lights.alias_table; auto& ap = lights.alias_pdf;
        for (size_t i = 0; i < n; i++) {
            auto& o = lights.objects[i];
            Color power = o->area() * o->material()->power(lam);
            Float pw = (power / lam.pdf()).mean();
            sum += pw; ap.push_back(pw); at.push_back({1.0, i});
        }
        std::vector<size_t> large, small; std::vector<Float> pdf;
        Float pu = 1.0 / (Float)n;
        for (size_t i = 0; i < n; i++) {
            ap[i] /= sum; pdf.push_back(ap[i]);
            if (ap[i] > pu) large.push_back(i); else small.push_back(i);
        }
        size_t is = small.size(), il = large.size();
        while (is > 0 && il > 0) {
            is--; il--;
            size_t s = small[is], l = large[il];
            at[s] = {pdf[s] * (Float)n, l};
            pdf[l] += pdf[s] - pu;
            if (pdf[l] > pu) { large[il] = l; il++; } else { small[is] = l; is++; }
        }
        while (is > 0) { is--; at[small[is]].first = 1.0; }
        while (il > 0) { il--; at[large[il]].first = 1.0; }
    }
    size_t num_lights() const { return lights.objects.size(); }
    size_t num_shadow_rays() const {                                                           // scene.rs:90-92
        size_t n = num_lights(); size_t lg = 0; while ((n >> (lg + 1)) != 0) lg++;
        return std::max<size_t>(lg, 1);
    }
    size_t sample_light(Float u) const {                                                       // bvh.rs:67-77
        Float ru = u * (Float)lights.objects.size();
        size_t idx = (size_t)sat_u64(std::floor(ru));
        Float fr = fract(ru);
        return fr < lights.alias_table[idx].first ? idx : lights.alias_table[idx].second;
    }
    const Object* get_light(size_t idx, Float& pdf) const { pdf = lights.alias_pdf[idx]; return lights.objects[idx].get(); }
    int64_t get_light_at(const Hit& h) const {                                                 // bvh.rs:97-102
        Vec3 xo = hit_ray_origin(h, true);
        Ray ri = Ray::make(xo, -h.ng);
        return lights._hit<true>(ri, 0.0, INF);
    }
    bool hit(const Ray& r, Hit& h) const {                                                     // scene.rs:119-147
        g_cnt.closest++;
        Float t_max = INF;
        bool have = objects.hit(r, 0.0, t_max, h);
        if (have) t_max = h.t;
        Hit hl;
        if (lights.hit(r, 0.0, t_max, hl)) { h = hl; h.obj += (int32_t)objects.objects.size(); have = true; }
        return have;
    }
    Float hit_t(const Ray& r) const {                                                          // scene.rs:150-162
        g_cnt.occlusion++;
        Float t = INF;
        t = fmin_(t, objects.hit_t(r, 0.0, t));
        t = fmin_(t, lights.hit_t(r, 0.0, t));
        return t;
    }
    bool occluded(const Ray& r, Float t_max) const {   // the two tests of hit_light (scene.rs:180-186)
        g_cnt.occlusion++;
        if (objects.hit_t(r, 0.0, t_max) < t_max) return true;
        if (lights.hit_t(r, 0.0, t_max) < t_max) return true;
        return false;
    }
    bool hit_light(const Ray& r, const Object* light, Hit& lh) const {                         // scene.rs:165-189
        if (!light->hit(r, 0.0, INF, lh)) return false;
        Float t_max = lh.t - EPSILON;
        return !occluded(r, t_max);
    }
};

// ---- integrators (src/tracer/integrator.rs, integrator/path_trace.rs, direct_light.rs) --------
static thread_local bool g_dbg = false;
static inline Color mis_sample(const Scene&, Vec3 wo, Vec3 wi, const Hit& ho, const Hit& hi, const Lambda& lam, bool li, Float p_lig, Float p_sct) {  // integrator.rs:139-184
    if (p_lig == 0.0 || p_sct == 0.0) return BLACK;
    const Material* m = ho.material;
    Color bsdf = m->bsdf_f(wo, wi, lam, RADIANCE, ho);
    auto heur = [](Float p) { return p * p; };
    Float denom = heur(p_lig) + heur(p_sct);
    Float weight = li ? heur(p_lig) / denom : heur(p_sct) / denom;
    Float p_denom = li ? p_lig : p_sct;
    return bsdf * WHITE * hi.material->emit(lam, hi) * m->shading_cosine(wi, ho.ns) * weight / p_denom;
}
static inline Color single_shadow_ray(const Scene& sc, Vec3 wo, Lambda& lam, const Hit& ho, Rng& rng) {  // integrator.rs:89-137
    const Material* m = ho.material;
    Vec3 xo = ho.p;
    Float pdf_light;
    const Object* light = sc.get_light(sc.sample_light(rng.gen_float()), pdf_light);
    Color radiance = BLACK;
    {
        Vec3 wi = light->sample_towards(xo, rng.gen_vec2());
        Ray ri = hit_generate_ray(ho, wi);
        Hit hi;
        if (sc.hit_light(ri, light, hi)) {
            Float p_lig = light->sample_towards_pdf(ri, hi.p, hi.ng);
            Float p_sct = m->bsdf_pdf(wo, wi, ho, lam, false);
            Color c_ = mis_sample(sc, wo, wi, ho, hi, lam, true, p_lig, p_sct);
            if (g_dbg) std::printf("  [oracle] A vis t=%.17g p_lig=%.17g p_sct=%.17g c0=%.17g\n", hi.t, p_lig, p_sct, c_.s[0]);
            radiance = radiance + c_;
        }
    }
    Float rand_u = rng.gen_float();
    Vec2 rs = rng.gen_vec2();
    Vec3 wi;
    if (m->bsdf_sample(wo, ho, lam, rand_u, rs, wi)) {
        Ray ri = hit_generate_ray(ho, wi);
        Hit hi;
        if (sc.hit_light(ri, light, hi)) {
            Float p_lig = light->sample_towards_pdf(ri, hi.p, hi.ng);
            Float p_sct = m->bsdf_pdf(wo, wi, ho, lam, false);
            Color c_ = mis_sample(sc, wo, wi, ho, hi, lam, false, p_lig, p_sct);
            if (g_dbg) std::printf("  [oracle] B vis t=%.17g p_lig=%.17g p_sct=%.17g c0=%.17g\n", hi.t, p_lig, p_sct, c_.s[0]);
            radiance = radiance + c_;
        }
    }
    return radiance / pdf_light;
}
static inline Color shadow_rays(const Scene& sc, Vec3 wo, Color gathered, Lambda& lam, const Hit& ho, Rng& rng) {  // integrator.rs:74-87
    Color acc = BLACK;
    size_t n = sc.num_shadow_rays();
    for (size_t i = 0; i < n; i++) acc = acc + gathered * single_shadow_ray(sc, wo, lam, ho, rng);
    return acc / (Float)n;
}

static const size_t RR_DEPTH = 5;
static inline FilmSample path_trace(const Scene& sc, Ray ro, Rng& rng, Lambda lam, Float delta, Vec2 raster_xy) {  // path_trace.rs:5-82
    bool last_specular = true;
    Color radiance = BLACK, gathered = WHITE;
    size_t depth = 0;
    Hit ho;
    while (sc.hit(ro, ho)) {
        const Material* m = ho.material;
        gathered = gathered * WHITE;   // transmittance without a medium (scene.rs:111-116)
        if (g_dbg) std::printf("[oracle] depth=%zu obj=%d tri=%d t=%.17g g0=%.17g rad0=%.17g\n", depth, (int)ho.obj, (int)ho.tri, ho.t, gathered.s[0], radiance.s[0]);
        Vec3 wo = -ro.dir;
        Float ru = rng.gen_float(); Vec2 rs = rng.gen_vec2();
        Vec3 wi;
        if (!m->bsdf_sample(wo, ho, lam, ru, rs, wi)) {
            if (last_specular) radiance = radiance + gathered * m->emit(lam, ho);
            break;
        }
        if (!m->is_delta(lam)) radiance = radiance + shadow_rays(sc, -ro.dir, gathered, lam, ho, rng);
        Ray ri = hit_generate_ray(ho, wi);
        wi = ri.dir;
        Float p_scatter = m->bsdf_pdf(wo, wi, ho, lam, false);
        if (p_scatter <= 0.0) break;
        Color bsdf = m->bsdf_f(wo, wi, lam, RADIANCE, ho);
        gathered = gathered * (bsdf * m->shading_cosine(wi, ho.ns) / p_scatter);
        if (depth >= RR_DEPTH) {
            Float lum = color_luminance(gathered, lam);
            Float rr = fmin_(lum / delta, 1.0);
            if (rng.gen_float() > rr) break;
            gathered = gathered / rr;
        }
        last_specular = m->is_specular();
        depth += 1;
        ro = ri;
    }
    if (g_dbg) std::printf("[oracle] end depth=%zu rad0=%.17g\n", depth, radiance.s[0]);
    return FilmSample{raster_xy, radiance, lam, false, depth};
}
static inline FilmSample direct_light(const Scene& sc, Ray ro, Rng& rng, Lambda lam, Vec2 raster_xy) {  // direct_light.rs:5-73
    const size_t MAX_RECURSION = 50;
    size_t depth = 0;
    Color radiance = BLACK, gathered = WHITE;
    Hit ho;
    while (sc.hit(ro, ho)) {
        const Material* m = ho.material;
        gathered = gathered * WHITE;
        Vec3 wo = -ro.dir;
        Float ru = rng.gen_float(); Vec2 rs = rng.gen_vec2();
        Vec3 wi;
        if (!m->bsdf_sample(wo, ho, lam, ru, rs, wi)) { radiance = radiance + gathered * m->emit(lam, ho); break; }
        if (!m->is_specular()) { radiance = radiance + shadow_rays(sc, -ro.dir, gathered, lam, ho, rng); break; }
        if (depth >= MAX_RECURSION) break;
        Ray ri = hit_generate_ray(ho, wi);
        wi = ri.dir;
        Float p_scatter = m->bsdf_pdf(wo, wi, ho, lam, false);
        if (p_scatter <= 0.0) break;
        Color bsdf = m->bsdf_f(wo, wi, lam, RADIANCE, ho);
        gathered = gathered * (bsdf * m->shading_cosine(wi, ho.ns) / p_scatter);
        depth += 1;
        ro = ri;
    }
    return FilmSample{raster_xy, radiance, lam, false, depth + 1};
}

}  // namespace oracle
