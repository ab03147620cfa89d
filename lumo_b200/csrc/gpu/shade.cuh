// Device-side shading: spectra, hero-wavelength colour, Hit reconstruction, materials / BSDFs,
// light sampling, camera rays, film accumulation, counter-based RNG.  f64 like the reference
// (the Spectrum polynomial is f32, spectrum.rs:108-118); compiled with -fmad=false.
// Every function cites the reference code it restates for the device.
#pragma once
#include "trace.cuh"
#include "../common/lumo_math.h"

namespace lumo_dev {

#define LUMO_PI 3.14159265358979323846264338327950288

// ---- numeric idioms -----------------------------------------------------------------
__device__ __forceinline__ double signum(double x) { return isnan(x) ? x : copysign(1.0, x); }
__device__ __forceinline__ double clampd(double v, double lo, double hi) { double r = v; if (r < lo) r = lo; if (r > hi) r = hi; return r; }
__device__ __forceinline__ double fractd(double x) { return x - trunc(x); }
__device__ __forceinline__ unsigned long long sat_u64(double x) { if (!(x > 0.0)) return 0ull; if (x >= 18446744073709551616.0) return ~0ull; return (unsigned long long)x; }
__device__ __forceinline__ double powi(double a, int b) {   // compiler-rt __powidf2
    double r = 1.0;
    while (true) { if (b & 1) r *= a; b /= 2; if (b == 0) break; a *= a; }
    return r;
}

// ---- counter-based RNG (replaces the per-tile sequential xorshift, src/rng.rs; north_star) -------
// Philox4x32-10 keyed by the render seed; counter = (pixel, sample, draw/2, stream).  Draw k of a
// path is the low/high 64 bits of block k/2, so a path's random numbers depend only on
// (seed, pixel, sample index, k) — independent of scheduling, wave size and GPU count.
struct Rng { uint32_t k0, k1, pixel, sample, stream, draws; uint32_t buf[4]; };
__device__ __forceinline__ void philox_block(uint32_t* c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c[0];
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ Rng rng_make(unsigned long long seed, uint32_t pixel, uint32_t sample, uint32_t stream, uint32_t draws) {
    Rng r; r.k0 = (uint32_t)seed; r.k1 = (uint32_t)(seed >> 32); r.pixel = pixel; r.sample = sample; r.stream = stream; r.draws = draws;
    if (draws & 1u) { r.buf[0] = pixel; r.buf[1] = sample; r.buf[2] = draws >> 1; r.buf[3] = stream; philox_block(r.buf, r.k0, r.k1); }
    return r;
}
__device__ __forceinline__ unsigned long long rng_u64(Rng& r) {
    if ((r.draws & 1u) == 0) { r.buf[0] = r.pixel; r.buf[1] = r.sample; r.buf[2] = r.draws >> 1; r.buf[3] = r.stream; philox_block(r.buf, r.k0, r.k1); }
    const unsigned long long v = (r.draws & 1u) == 0 ? (((unsigned long long)r.buf[1] << 32) | r.buf[0]) : (((unsigned long long)r.buf[3] << 32) | r.buf[2]);
    r.draws++;
    return v;
}
// Xorshift::gen_float (rng.rs:71-75): u64 * 2^-64, clamped below 1
__device__ __forceinline__ double rng_float(Rng& r) { return fmin(__ull2double_rn(rng_u64(r)) * 5.421010862427522170037e-20, 1.0 - LUMO_EPS); }

// Correlated multi-jitter permutation (Kensler 2013): stateless stand-in for the Fisher-Yates
// tables of MultiJitteredSampler (samplers.rs:136-192)
__device__ __forceinline__ uint32_t cmj_permute(uint32_t i, uint32_t l, uint32_t p) {
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

// ---- Color / wavelengths (color.rs, color/wavelength.rs, dense_spectrum.rs, spectrum.rs) ----------
struct C4 { double s[4]; };
__device__ __forceinline__ C4 c4(double v) { C4 c; c.s[0] = c.s[1] = c.s[2] = c.s[3] = v; return c; }
#define LUMO_C4_OP(op) \
    __device__ __forceinline__ C4 operator op(C4 a, C4 b) { C4 r; _Pragma("unroll") for (int i = 0; i < 4; i++) r.s[i] = a.s[i] op b.s[i]; return r; } \
    __device__ __forceinline__ C4 operator op(C4 a, double b) { C4 r; _Pragma("unroll") for (int i = 0; i < 4; i++) r.s[i] = a.s[i] op b; return r; } \
    __device__ __forceinline__ C4 operator op(double a, C4 b) { C4 r; _Pragma("unroll") for (int i = 0; i < 4; i++) r.s[i] = a op b.s[i]; return r; }
LUMO_C4_OP(+) LUMO_C4_OP(-) LUMO_C4_OP(*)
__device__ __forceinline__ C4 operator/(C4 a, C4 b) { C4 r; _Pragma("unroll") for (int i = 0; i < 4; i++) r.s[i] = b.s[i] == 0.0 ? 0.0 : a.s[i] / b.s[i]; return r; }   // color.rs:239-249
__device__ __forceinline__ C4 operator/(C4 a, double b) { C4 r; _Pragma("unroll") for (int i = 0; i < 4; i++) r.s[i] = b == 0.0 ? 0.0 : a.s[i] / b; return r; }          // color.rs:251-271
__device__ __forceinline__ bool is_black(C4 c) { return c.s[0] == 0.0 && c.s[1] == 0.0 && c.s[2] == 0.0 && c.s[3] == 0.0; }
__device__ __forceinline__ double mean4(C4 c) { return (((0.0 + c.s[0]) + c.s[1]) + c.s[2] + c.s[3]) / 4.0; }

struct Lam { double l[4]; };
__device__ __forceinline__ double lam_sample_one(double v) { return 538.0 - 138.888889 * lm_atanh(0.85691062 - 253.819 * v * 0.0072); }   // wavelength.rs:48-51
__device__ __forceinline__ double lam_pdf_one(double l) {                                                                       // wavelength.rs:60-66
    if (l < 360.0 || l > 830.0) return 0.0;
    const double c = lm_cosh(0.0072 * (l - 538.05));
    return 1.0 / (253.819 * (c * c));
}
__device__ __noinline__ Lam lam_sample(double u) {                                                                           // wavelength.rs:35-44
    Lam r;
#pragma unroll
    for (int i = 0; i < 4; i++) { double v = u + (double)i / 4.0; v = v > 1.0 ? v - 1.0 : v; r.l[i] = lam_sample_one(v); }
    return r;
}
__device__ __forceinline__ bool lam_terminated(const Lam& l) { return l.l[1] == 0.0 && l.l[2] == 0.0 && l.l[3] == 0.0; }
__device__ __noinline__ C4 lam_pdf(const Lam& l) {                                                                           // wavelength.rs:24-32
    C4 c;
#pragma unroll
    for (int i = 0; i < 4; i++) c.s[i] = lam_pdf_one(l.l[i]);
    if (lam_terminated(l)) c.s[0] /= 4.0;
    return c;
}
__device__ __forceinline__ double dense_one(const double* v, double lambda) {                                                   // dense_spectrum.rs:77-97
    const double STEP = (830.0 - 360.0) / (95.0 - 1.0);
    const unsigned long long b1 = sat_u64(ceil((lambda - 360.0) / STEP));
    const double l1 = 360.0 + STEP * (double)b1;
    if (lambda == 0.0) return 0.0;
    if (lambda == l1) return __ldg(v + b1);
    const unsigned long long b0 = b1 - 1;
    const double l0 = l1 - STEP;
    const double x1 = (lambda - l0) / STEP, x0 = 1.0 - x1;
    return __ldg(v + b0) * x0 + __ldg(v + b1) * x1;
}
__device__ __forceinline__ C4 dense4(const double* v, const Lam& l) { C4 c; _Pragma("unroll") for (int i = 0; i < 4; i++) c.s[i] = dense_one(v, l.l[i]); return c; }
__device__ __forceinline__ double spec_one(const float* c, double lambda) {                                                     // spectrum.rs:108-118
    const float l = (float)lambda;
    const float x = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(c[0], l), l), __fmul_rn(c[1], l)), c[2]);
    const float sg = __fadd_rn(0.5f, __fdiv_rn(x, __fmul_rn(2.0f, __fsqrt_rn(__fadd_rn(1.0f, __fmul_rn(x, x))))));
    return (double)__fmul_rn(c[3], sg);
}
__device__ __forceinline__ C4 spec4(const float* c, const Lam& l) { C4 r; _Pragma("unroll") for (int i = 0; i < 4; i++) r.s[i] = spec_one(c, l.l[i]); return r; }

// ---- Textures (texture.rs:23-113, perlin.rs:48-110, image.rs:99-193) -------------------------------
// Materials whose kd / ks / tf / ke is a Texture::Solid keep the spectrum inline (m.kd ...) and never come here.
__device__ __forceinline__ double perlin_noise(const double* tab, D3 p) {                                                       // perlin.rs:50-68
    const D3 weight = d3(fractd(p.x), fractd(p.y), fractd(p.z));
    const unsigned long long fx = sat_u64(floor(p.x)), fy = sat_u64(floor(p.y)), fz = sat_u64(floor(p.z));
    const double* px = tab + 768; const double* py = px + 256; const double* pz = py + 256;
    const D3 w = d3(((6.0 * weight.x - 15.0) * weight.x + 10.0) * weight.x * weight.x * weight.x,                                // perlin.rs:83-85
                    ((6.0 * weight.y - 15.0) * weight.y + 10.0) * weight.y * weight.y * weight.y,
                    ((6.0 * weight.z - 15.0) * weight.z + 10.0) * weight.z * weight.z * weight.z);
    double acc = 0.0;
#pragma unroll 1
    for (int c = 0; c < 8; c++) {                                                                                                // (i, j, k), k fastest
        const unsigned long long i = (unsigned long long)(c >> 2), j = (unsigned long long)((c >> 1) & 1), k = (unsigned long long)(c & 1);
        const unsigned long long h = (unsigned long long)__ldg(px + ((fx + i) & 255ull)) ^ (unsigned long long)__ldg(py + ((fy + j) & 255ull)) ^ (unsigned long long)__ldg(pz + ((fz + k) & 255ull));
        const D3 n = d3(__ldg(tab + 3 * h), __ldg(tab + 3 * h + 1), __ldg(tab + 3 * h + 2));
        const D3 idx = d3((double)i, (double)j, (double)k);
        const D3 widx = d3(2.0 * w.x * idx.x + 1.0 - w.x - idx.x, 2.0 * w.y * idx.y + 1.0 - w.y - idx.y, 2.0 * w.z * idx.z + 1.0 - w.z - idx.z);   // perlin.rs:92-109
        acc = acc + widx.x * widx.y * widx.z * dot(n, w - idx);
    }
    return acc;
}
// Image::bilin_interp (image.rs:99-131): the four texel indices and the two weights
struct Taps { uint32_t i00, i10, i01, i11; double wx, wy; };
__device__ __forceinline__ uint32_t sat_u32(double v) { const unsigned long long u = sat_u64(v); return u > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)u; }
__device__ __forceinline__ Taps image_taps(uint32_t width, uint32_t height, double u, double v) {
    const double w = (double)width, h = (double)height;
    const double x = u * w, y = (1.0 - v) * h;
    const double xo_f = floor(x - 0.5), yo_f = floor(y - 0.5);
    Taps t;
    t.wx = 1.0 - (x - xo_f - 0.5); t.wy = 1.0 - (y - yo_f - 0.5);
    const uint32_t xo = sat_u32(xo_f + w) % width, yo = sat_u32(yo_f + h) % height;
    const uint32_t xi = (xo + 1u) % width, yi = (yo + 1u) % height;
    t.i00 = xo + yo * width; t.i10 = xi + yo * width; t.i01 = xo + yi * width; t.i11 = xi + yi * width;
    return t;
}
__device__ __noinline__ C4 texture_albedo(const DevScene& S, uint32_t tex, double u, double v, const Lam& l) {                  // texture.rs:53-92
    const LumoTexture* T = S.textures + tex;
    while (T->kind == LTEX_CHECKER) {                                                                                            // children see the same uv
        const double su = u * T->scale, sv = v * T->scale;
        T = S.textures + ((sat_u64(floor(su) + floor(sv)) % 2ull == 0ull) ? T->a : T->b);
    }
    switch (T->kind) {
    case LTEX_MARBLE: {
        const double* tab = S.tex_f64 + T->data;
        D3 p = d3(4.0 * fabs(u), 4.0 * fabs(v), 4.0 * fabs(0.0));
        double turb = 0.0;
#pragma unroll 1
        for (int depth = 0; depth < 6; depth++) {                                                                                // turbulence, texture.rs:105-112
            turb = turb + powi(0.5, depth) * fabs(perlin_noise(tab, p));
            p = 2.0 * p;
        }
        const double scaled = 1.0 - powi(0.5 + 0.5 * lm_sin(60.0 * u + 20.0 * turb), 6);
        return spec4(T->spec, l) * scaled;
    }
    case LTEX_IMAGE: {                                                                                                           // image.rs:184-193
        const Taps t = image_taps(T->width, T->height, u, v);
        const float* px = S.tex_pixels + 4 * T->data;
        const C4 y0 = spec4(px + 4 * (size_t)t.i00, l) * t.wx + spec4(px + 4 * (size_t)t.i10, l) * (1.0 - t.wx);
        const C4 y1 = spec4(px + 4 * (size_t)t.i01, l) * t.wx + spec4(px + 4 * (size_t)t.i11, l) * (1.0 - t.wx);
        return y0 * t.wy + y1 * (1.0 - t.wy);
    }
    case LTEX_MANDELBROT: {
        int depth = 0;
        const double cre = 2.0 * (u - 0.75), cim = 2.0 * (v - 0.5);
        double zre = 0.0, zim = 0.0;
        while (depth < 256 && zre * zre + zim * zim < 64.0 * 64.0) {
            const double nre = zre * zre - zim * zim, nim = zre * zim + zim * zre;
            zre = nre + cre; zim = nim + cim;
            depth++;
        }
        return depth == 256 ? c4(1.0) : c4(0.0);
    }
    default: return spec4(T->spec, l);
    }
}
// The per-kind shade kernels also come in a "no textures anywhere in this scene" flavour (K | LUMO_K_SOLID): the texture and
// bump-map branches (and the calls they keep live registers for) are compiled out — untextured scenes shade as fast as before
// textures existed (bunny shade 29.5 -> 28.2 ms).  K < 0: runtime kind, textures on.
#define LUMO_K_SOLID 8
#define LUMO_KIND(K, m) ((K) < 0 ? (m).kind : (uint32_t)((K) & 7))
#define LUMO_TEX(K) ((K) < 0 || !((K) & LUMO_K_SOLID))
template <bool TEX = true>
__device__ __forceinline__ C4 tex4(const DevScene& S, const float* solid, uint32_t tex, double u, double v, const Lam& l) {
    if (!TEX) return spec4(solid, l);
    return tex == LUMO_NONE ? spec4(solid, l) : texture_albedo(S, tex, u, v, l);
}
#define LUMO_Y_INTEGRAL 106.856895
__device__ __forceinline__ const double* table(const DevScene& S, uint32_t id) { return S.tables + 96ull * id; }
__device__ __noinline__ double luminance(const DevScene& S, C4 c, const Lam& l) {                                            // color.rs:88-91
    const C4 pdf = lam_pdf(l);
    return mean4(dense4(table(S, LTAB_Y), l) * c / pdf) / LUMO_Y_INTEGRAL;
}
__device__ __noinline__ D3 color_xyz(const DevScene& S, C4 c, const Lam& l) {                                                // color.rs:93-101
    const C4 pdf = lam_pdf(l);
    return d3(mean4(dense4(table(S, LTAB_X), l) * c / pdf), mean4(dense4(table(S, LTAB_Y), l) * c / pdf), mean4(dense4(table(S, LTAB_Z), l) * c / pdf)) / LUMO_Y_INTEGRAL;
}

// ---- Hit (hit.rs) -----------------------------------------------------------------------------------
struct DevHit { double t; D3 p, fp_error, ns, ng; double u, v; bool backface; int material; };
__device__ __forceinline__ void wrap_uv(double& u, double& v) {                                                                 // hit.rs:62-68
    const double fu = fractd(u), fv = fractd(v);
    u = fu < 0.0 ? fu + 1.0 : fu; v = fv < 0.0 ? fv + 1.0 : fv;
}
__device__ __forceinline__ D3 hit_ray_origin(const DevHit& h, bool outside) {                                                   // hit.rs:85-111
    const D3 ne = h.ng;
    const double scaled = dot(h.fp_error, vabs(ne));
    const D3 off = outside ? ne * scaled : (-ne) * scaled;
    const D3 xi = h.p + off;
    return d3(off.x > 0.0 ? next_float(xi.x) : (off.x < 0.0 ? previous_float(xi.x) : xi.x),
              off.y > 0.0 ? next_float(xi.y) : (off.y < 0.0 ? previous_float(xi.y) : xi.y),
              off.z > 0.0 ? next_float(xi.z) : (off.z < 0.0 ? previous_float(xi.z) : xi.z));
}
__device__ __noinline__ Ray hit_generate_ray(const DevHit& h, D3 wi) {                                                       // hit.rs:115-122
    Ray r; r.o = hit_ray_origin(h, dot(wi, h.ng) >= 0.0); r.d = normalize(wi); return r;
}
__device__ __forceinline__ D3 mul33(const double* m, D3 v) { return d3(dot(d3(m[0], m[1], m[2]), v), dot(d3(m[3], m[4], m[5]), v), dot(d3(m[6], m[7], m[8]), v)); }
// Instance::propagate_fp_err (instance.rs:40-50) with |m| rows
__device__ __forceinline__ D3 propagate_fp_err(const LumoInstance* I, D3 xo, D3 fp) {
    const D3 e3 = vabs(fp), p3 = vabs(xo);
    const double* m = I->m;
    const D3 tp = d3(fabs(m[0]) * p3.x + fabs(m[1]) * p3.y + fabs(m[2]) * p3.z + fabs(m[3]) * 1.0,
                     fabs(m[4]) * p3.x + fabs(m[5]) * p3.y + fabs(m[6]) * p3.z + fabs(m[7]) * 1.0,
                     fabs(m[8]) * p3.x + fabs(m[9]) * p3.y + fabs(m[10]) * p3.z + fabs(m[11]) * 1.0) / 1.0;
    if (e3.x == 0.0 && e3.y == 0.0 && e3.z == 0.0) return gamma_n(3) * tp;
    const D3 td = d3(fabs(m[0]) * e3.x + fabs(m[1]) * e3.y + fabs(m[2]) * e3.z + fabs(m[3]) * 0.0,
                     fabs(m[4]) * e3.x + fabs(m[5]) * e3.y + fabs(m[6]) * e3.z + fabs(m[7]) * 0.0,
                     fabs(m[8]) * e3.x + fabs(m[9]) * e3.y + fabs(m[10]) * e3.z + fabs(m[11]) * 0.0);
    return gamma_n(3) * tp + (gamma_n(3) + 1.0) * td;
}
// Rebuilds the reference's `Hit` from (object, triangle, t, barycentrics): the GEO tail of
// Triangle::_hit (triangle.rs:155-186), Sphere::hit (sphere.rs:62-74), Rectangle::hit's uv override
// (rectangle.rs:73-85) and Instance::hit's world transform (instance.rs:84-99).
// UV = false (scenes without a single texture record): a sphere's (u, v) — an atan2 and an acos — is not computed, nothing reads it.
template <bool UV = true>
__device__ __noinline__ DevHit reconstruct_hit(const DevScene& S, const Ray& r, const HitRec& rec) {
    const LumoObject o = S.objects[rec.obj];
    const Ray l = to_local<false>(S, o, r, nullptr);
    DevHit h; h.t = rec.t; h.material = o.material;
    if (o.kind == LOBJ_SPHERE) {
        const double radius = S.spheres[o.geom].radius;
        D3 xi = l.o + rec.t * l.d;
        xi = xi * radius / length(xi);
        h.fp_error = gamma_n(5) * vabs(xi);
        const D3 ni = xi / radius;
        if (UV) {
            h.u = (lm_atan2(-ni.z, ni.x) + LUMO_PI) / (2.0 * LUMO_PI);
            h.v = lm_acos(-ni.y) / LUMO_PI;
            wrap_uv(h.u, h.v);
        } else { h.u = 0.0; h.v = 0.0; }
        h.p = xi; h.ns = ni; h.ng = ni;
        h.backface = dot(l.d, ni) > 0.0;
    } else {
        const uint32_t ti = (o.kind == LOBJ_TRI) ? o.geom : S.kd_trees[o.geom].tri_base + rec.tri;
        D3 A, B, C; load_tri(S.tri_verts + ti, A, B, C);
        const double al = rec.bary.x, be = rec.bary.y, ga = rec.bary.z;
        const D3 ng = normalize(cross(B - A, C - A));
        const LumoTriShade sh = S.tri_shade[ti];
        D3 ns = ng;
        if (sh.flags & 1u) {                                                                                                    // triangle.rs:49-60
            const double* n = S.normals;
            const D3 na = d3(n[3 * sh.n[0]], n[3 * sh.n[0] + 1], n[3 * sh.n[0] + 2]), nb = d3(n[3 * sh.n[1]], n[3 * sh.n[1] + 1], n[3 * sh.n[1] + 2]),
                     nc = d3(n[3 * sh.n[2]], n[3 * sh.n[2] + 1], n[3 * sh.n[2] + 2]);
            ns = normalize(al * na + be * nb + ga * nc);
        }
        h.p = al * A + be * B + ga * C;
        double tau = 0.0, tav = 0.0, tbu = 1.0, tbv = 0.0, tcu = 1.0, tcv = 1.0;
        if (sh.flags & 2u) { const double* t = S.uvs; tau = t[2 * sh.t[0]]; tav = t[2 * sh.t[0] + 1]; tbu = t[2 * sh.t[1]]; tbv = t[2 * sh.t[1] + 1]; tcu = t[2 * sh.t[2]]; tcv = t[2 * sh.t[2] + 1]; }
        h.u = al * tau + be * tbu + ga * tcu; h.v = al * tav + be * tbv + ga * tcv;
        h.fp_error = gamma_n(7) * d3(fabs(al * A.x) * 1.0 + fabs(be * B.x) * 1.0 + fabs(ga * C.x) * 1.0,
                                     fabs(al * A.y) * 1.0 + fabs(be * B.y) * 1.0 + fabs(ga * C.y) * 1.0,
                                     fabs(al * A.z) * 1.0 + fabs(be * B.z) * 1.0 + fabs(ga * C.z) * 1.0);
        wrap_uv(h.u, h.v);
        h.ns = ns; h.ng = ng;
        h.backface = dot(l.d, ng) > 0.0;
        if (o.kind == LOBJ_RECT) {
            const LumoRect* R = S.rects + o.rect;
            h.u = dot(d3(R->b0[0], R->b0[1], R->b0[2]), h.p); h.v = dot(d3(R->b1[0], R->b1[1], R->b1[2]), h.p);
            wrap_uv(h.u, h.v);
        }
    }
    if (o.inst >= 0) {
        const LumoInstance* I = S.instances + o.inst;
        h.ns = normalize(mul33(I->nrm, h.ns));
        h.ng = normalize(mul33(I->nrm, h.ng));
        h.fp_error = propagate_fp_err(I, h.p, h.fp_error);
        h.p = xf_point(I->m, h.p);
    }
    return h;
}

// ---- Onb (onb.rs:19-62, Duff et al.) ---------------------------------------------------------------
struct Onb { D3 u, v, w; };
__device__ __noinline__ Onb onb_new(D3 w) {
    const double sgn = signum(w.z);
    const double a = -1.0 / (sgn + w.z);
    const double b = w.x * w.y * a;
    Onb o; o.w = w;
    o.u = d3(1.0 + sgn * w.x * w.x * a, sgn * b, -sgn * w.x);
    o.v = d3(b, sgn + w.y * w.y * a, -w.y);
    return o;
}
__device__ __forceinline__ D3 to_world(const Onb& o, D3 p) { return p.x * o.u + p.y * o.v + p.z * o.w; }
__device__ __forceinline__ D3 to_local(const Onb& o, D3 p) { return d3(dot(p, o.u), dot(p, o.v), dot(p, o.w)); }
// Image<Normal>::value_at (image.rs:134-151) + Material::map_normal (material.rs:324-331)
__device__ __noinline__ D3 bump_normal(const DevScene& S, uint32_t tex, D3 ns, double u, double v) {
    const LumoTexture* T = S.textures + tex;
    const Taps t = image_taps(T->width, T->height, u, v);
    const double* px = S.tex_f64 + T->data;
    const D3 n00 = d3(px[3 * (size_t)t.i00], px[3 * (size_t)t.i00 + 1], px[3 * (size_t)t.i00 + 2]), n10 = d3(px[3 * (size_t)t.i10], px[3 * (size_t)t.i10 + 1], px[3 * (size_t)t.i10 + 2]);
    const D3 n01 = d3(px[3 * (size_t)t.i01], px[3 * (size_t)t.i01 + 1], px[3 * (size_t)t.i01 + 2]), n11 = d3(px[3 * (size_t)t.i11], px[3 * (size_t)t.i11 + 1], px[3 * (size_t)t.i11 + 2]);
    const D3 y0 = normalize(n00 * t.wx + n10 * (1.0 - t.wx)), y1 = normalize(n01 * t.wx + n11 * (1.0 - t.wx));
    const D3 n = normalize(y0 * t.wy + y1 * (1.0 - t.wy));
    return normalize(to_world(onb_new(ns), n));
}

// ---- rng/maps.rs -----------------------------------------------------------------------------------
__device__ __forceinline__ void square_to_disk(double r0, double r1, double& dx, double& dy) {
    const double ox = 2.0 * r0 - 1.0, oy = 2.0 * r1 - 1.0;
    if (ox == 0.0 && oy == 0.0) { dx = 0.0; dy = 0.0; return; }
    double rr, th;
    if (fabs(ox) > fabs(oy)) { rr = ox; th = LUMO_PI * (oy / ox) / 4.0; } else { rr = oy; th = LUMO_PI * (0.5 - (ox / oy) / 4.0); }
    double sn, cs; lm_sincos(th, &sn, &cs);
    dx = rr * cs; dy = rr * sn;
}
__device__ __noinline__ D3 square_to_cos_hemisphere(double r0, double r1) {
    double dx, dy; square_to_disk(r0, r1, dx, dy);
    return d3(dx, dy, sqrt(fmax(1.0 - dx * dx - dy * dy, 0.0)));
}
__device__ __forceinline__ D3 square_to_sphere(double r0, double r1) {
    const double z = 1.0 - 2.0 * r1;
    const double rr = sqrt(fmax(1.0 - z * z, 0.0));
    const double phi = 2.0 * LUMO_PI * r0;
    double sn, cs; lm_sincos(phi, &sn, &cs);
    return d3(rr * cs, rr * sn, z);
}

// ---- spherical utils (math/spherical_utils.rs) -----------------------------------------------------
__device__ __forceinline__ double cos2_theta(D3 w) { return w.z * w.z; }
__device__ __forceinline__ double sin2_theta(D3 w) { return fmax(1.0 - cos2_theta(w), 0.0); }
__device__ __forceinline__ double sin_theta(D3 w) { return sqrt(sin2_theta(w)); }
__device__ __forceinline__ double tan2_theta(D3 w) { return sin2_theta(w) / cos2_theta(w); }
__device__ __forceinline__ double cos_phi(D3 w) { const double s = sin_theta(w); return s == 0.0 ? 1.0 : clampd(w.x / s, -1.0, 1.0); }
__device__ __forceinline__ double sin_phi(D3 w) { const double s = sin_theta(w); return s == 0.0 ? 0.0 : clampd(w.y / s, -1.0, 1.0); }
__device__ __forceinline__ bool same_hemisphere(D3 a, D3 b) { return a.z * b.z > 0.0; }

// ---- materials (material.rs, bsdf.rs, bxdf.rs, microfacet.rs, bxdf/{microfacet,scatter}.rs) -------
typedef LumoMaterial Mat;
__device__ __forceinline__ bool mf_is_specular(const Mat& m) { return (m.roughness + m.roughness) / 2.0 < 0.01; }             // microfacet.rs:73-76
__device__ __forceinline__ bool mf_is_delta(const Mat& m) { return (m.roughness + m.roughness) / 2.0 < 1e-3; }                // microfacet.rs:80-83
__device__ __forceinline__ bool mat_is_standard(const Mat& m) { return m.kind >= LMAT_LAMBERTIAN && m.kind <= LMAT_MFDIELECTRIC; }
__device__ __forceinline__ bool mat_is_specular(const Mat& m) {                                                                // material.rs:203-209, bxdf.rs:35-41
    if (m.kind == LMAT_MFDIELECTRIC) return true;
    if (m.kind == LMAT_MFCONDUCTOR) return mf_is_specular(m);
    return false;
}
__device__ __forceinline__ bool bx_is_reflection(const Mat& m) { return m.kind != LMAT_MFDIELECTRIC; }                          // bxdf.rs:46-56
__device__ __forceinline__ double eta_at(const DevScene& S, const Mat& m, double wl) { return dense_one(table(S, m.eta_table), wl); }
__device__ __forceinline__ double k_at(const DevScene& S, const Mat& m, double wl) { return dense_one(table(S, m.k_table), wl); }
__device__ __forceinline__ bool mat_is_delta(const DevScene& S, const Mat& m, const Lam& l) {                                  // material.rs:212-217, bxdf.rs:59-67
    if (m.kind == LMAT_MFCONDUCTOR) return mf_is_delta(m);
    if (m.kind == LMAT_MFDIELECTRIC) return mf_is_delta(m) || eta_at(S, m, l.l[0]) == 1.0;
    return false;
}
template <int K = -1>
__device__ __forceinline__ C4 mat_emit(const DevScene& S, const Mat& m, const Lam& l, const DevHit& h) {                         // material.rs:220-231
    if (m.kind != LMAT_LIGHT) return c4(0.0);
    if (!(m.flags & LMF_TWO_SIDED) && h.backface) return c4(0.0);
    return m.scale * tex4<LUMO_TEX(K)>(S, m.ke, m.ke_tex, h.u, h.v, l) * dense4(table(S, m.illum_table), l);
}
// the shading frame of a hit: Onb::new(ns), with the normal map applied first for materials that have one (material.rs:258-310)
template <int K = -1>
__device__ __forceinline__ Onb shading_onb(const DevScene& S, const Mat& m, const DevHit& h) {
    if (!LUMO_TEX(K)) return onb_new(h.ns);
    return onb_new(m.bump_tex == LUMO_NONE ? h.ns : bump_normal(S, m.bump_tex, h.ns, h.u, h.v));
}
__device__ __forceinline__ double shading_cosine(const Mat& m, D3 wi, D3 ns) { return mat_is_standard(m) ? fabs(dot(ns, wi)) : 1.0; }   // material.rs:315-321

__device__ __forceinline__ double f_schlick(double f0, double f90, double c) { return f0 + (f90 - f0) * powi(1.0 - c, 5); }     // microfacet.rs:198-200
__device__ __forceinline__ double disney_diffuse(const Mat& m, double cwo, double cwi, double cwh) {                            // microfacet.rs:129-145
    const double r2 = powi(m.roughness, 2);
    const double fd90 = 0.5 * r2 + 2.0 * powi(cwh, 2) * r2;
    return f_schlick(1.0, fd90, cwo) * f_schlick(1.0, fd90, cwi) * (1.0 + r2 * (1.0 / 1.51 - 1.0));
}
__device__ __noinline__ double ggx_d(const Mat& m, D3 wh) {                                                                  // microfacet.rs:147-176
    const double tan2 = tan2_theta(wh);
    if (isinf(tan2)) return 0.0;
    const double cos4 = powi(cos2_theta(wh), 2);
    if (cos4 < powi(LUMO_EPS, 2)) return 0.0;
    const double cp = cos_phi(wh), sp = sin_phi(wh);
    const double alpha2 = m.roughness * m.roughness;
    const double e = tan2 * (powi(cp / m.roughness, 2) + powi(sp / m.roughness, 2));
    return 1.0 / (LUMO_PI * alpha2 * cos4 * powi(1.0 + e, 2));
}
struct Cx { double re, im; };
__device__ __forceinline__ Cx cx(double r, double i) { Cx c; c.re = r; c.im = i; return c; }
__device__ __forceinline__ Cx cmul(Cx a, Cx b) { return cx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ Cx cdivf(Cx a, double b) { if (b == 0.0) return cx(nan(""), nan("")); return cx(a.re / b, a.im / b); }
__device__ __forceinline__ Cx cdiv(Cx a, Cx b) { if (b.re == 0.0 && b.im == 0.0) return cx(nan(""), nan("")); return cdivf(cmul(a, cx(b.re, -b.im)), b.re * b.re + b.im * b.im); }
__device__ __forceinline__ Cx fdivc(double a, Cx b) { if (b.re == 0.0 && b.im == 0.0) return cx(nan(""), nan("")); return cdivf(cx(a * b.re, a * -b.im), b.re * b.re + b.im * b.im); }
__device__ __forceinline__ Cx csqrt(Cx a) {                                                                                      // complex.rs:38-53
    const double nr = sqrt(sqrt(a.re * a.re + a.im * a.im)), ar = lm_atan2(a.im, a.re) / 2.0;
    double sn, cs; lm_sincos(ar, &sn, &cs);
    return cx(nr * cs, nr * sn);
}
__device__ __forceinline__ double fr_complex(D3 wo, D3 wh, double eta_, double k_) {                                             // microfacet.rs:226-241
    const Cx eta = cx(eta_, k_);
    const double cos_o = clampd(dot(wo, wh), 0.0, 1.0);
    const double sin2_o = 1.0 - cos_o * cos_o;
    const Cx sin2_i = fdivc(sin2_o, cmul(eta, eta));
    const Cx cos_i = csqrt(cx(1.0 - sin2_i.re, -sin2_i.im));
    const Cx ec = cx(eta.re * cos_o, eta.im * cos_o);
    const Cx r_par = cdiv(cx(ec.re - cos_i.re, ec.im - cos_i.im), cx(ec.re + cos_i.re, ec.im + cos_i.im));
    const Cx eci = cmul(eta, cos_i);
    const Cx r_per = cdiv(cx(cos_o - eci.re, -eci.im), cx(cos_o + eci.re, eci.im));
    return ((r_par.re * r_par.re + r_par.im * r_par.im) + (r_per.re * r_per.re + r_per.im * r_per.im)) / 2.0;
}
__device__ __forceinline__ double fr_real(D3 wo, D3 wh, double eta_) {                                                           // microfacet.rs:244-265
    double cos_o = dot(wo, wh);
    const double eta = cos_o < 0.0 ? 1.0 / eta_ : eta_;
    cos_o = fabs(cos_o);
    const double sin2_o = 1.0 - cos_o * cos_o;
    const double sin2_i = sin2_o / (eta * eta);
    if (sin2_i >= 1.0) return 1.0;
    const double cos_i = sqrt(fmax(1.0 - sin2_i, 0.0));
    const double r_par = (eta * cos_o - cos_i) / (eta * cos_o + cos_i);
    const double r_per = (cos_o - eta * cos_i) / (cos_o + eta * cos_i);
    return (r_par * r_par + r_per * r_per) / 2.0;
}
__device__ __forceinline__ double fresnel_at(const DevScene& S, const Mat& m, D3 wo, D3 wh, double wl) {                        // microfacet.rs:210-223
    const double e = eta_at(S, m, wl), k = k_at(S, m, wl);
    if (k == 0.0) return e == 0.0 ? 0.0 : fr_real(wo, wh, e);
    return fr_complex(wo, wh, e, k);
}
// the four hero wavelengths are independent: unrolled so that their (long, dependent) f64 chains overlap
__device__ __noinline__ C4 fresnel4(const DevScene& S, const Mat& m, D3 wo, D3 wh, const Lam& l) { C4 c; _Pragma("unroll") for (int i = 0; i < 4; i++) c.s[i] = fresnel_at(S, m, wo, wh, l.l[i]); return c; }
__device__ __forceinline__ bool chi_pass(D3 wo, D3 wh) { return signum(wh.z) * dot(wo, wh) * wo.z > LUMO_EPS; }                 // microfacet.rs:268-274
__device__ __noinline__ double ggx_lambda(const Mat& m, D3 w) {                                                              // microfacet.rs:296-311
    const double tan2 = tan2_theta(w);
    if (isinf(tan2)) return 0.0;
    const double cp = cos_phi(w), sp = sin_phi(w);
    const double alpha2 = powi(m.roughness * cp, 2) + powi(m.roughness * sp, 2);
    return (sqrt(fmax(1.0 + alpha2 * tan2, 0.0)) - 1.0) / 2.0;
}
__device__ __forceinline__ double ggx_g(const Mat& m, D3 wo, D3 wi, D3 wh) { return !chi_pass(wo, wh) ? 0.0 : 1.0 / (1.0 + ggx_lambda(m, wo) + ggx_lambda(m, wi)); }
__device__ __forceinline__ double ggx_g1(const Mat& m, D3 wo, D3 wh) { return !chi_pass(wo, wh) ? 0.0 : 1.0 / (1.0 + ggx_lambda(m, wo)); }
__device__ __forceinline__ double sample_normal_pdf(const Mat& m, D3 wh, D3 wo) {                                               // microfacet.rs:330-349
    return fmax(ggx_g1(m, wo, wh) * ggx_d(m, wh) * fabs(dot(wh, wo)) / fabs(wo.z), 0.0);
}
__device__ __noinline__ D3 sample_normal(const Mat& m, D3 wo, double r0, double r1) {                                        // microfacet.rs:352-430 (Heitz 2018)
    D3 ws = normalize(d3(wo.x * m.roughness, wo.y * m.roughness, wo.z));
    if (ws.z < 0.0) ws = -ws;
    const D3 u = (1.0 - ws.z < LUMO_EPS) ? d3(1, 0, 0) : normalize(cross(ws, d3(0, 0, 1)));
    const D3 v = cross(u, ws);
    const double r = sqrt(r0);
    const double th = 2.0 * LUMO_PI * r1;
    double sn, cs; lm_sincos(th, &sn, &cs);
    const double x = r * cs;
    const double h = sqrt(fmax(1.0 - x * x, 0.0));
    const double lerp = (1.0 + ws.z) / 2.0;
    const double y = (1.0 - lerp) * h + lerp * r * sn;
    D3 wm = d3(x, y, sqrt(fmax(1.0 - x * x - y * y, 0.0)));
    wm = wm.x * u + wm.y * v + wm.z * ws;
    return normalize(d3(m.roughness * wm.x, m.roughness * wm.y, fmax(wm.z, LUMO_EPS)));
}
__device__ __forceinline__ bool reflect(D3 wo, D3 no, D3& wi) {                                                                 // bxdf/microfacet.rs:7-15
    const D3 proj = no * dot(wo, no) / dot(no, no);
    wi = 2.0 * proj - wo;
    return same_hemisphere(wi, wo);
}
__device__ __forceinline__ bool refract(double eta, D3 wo, D3 no, D3& wi) {                                                     // bxdf/microfacet.rs:17-44
    double cos_to, ratio; D3 n;
    if (dot(no, wo) < 0.0) { cos_to = -dot(no, wo); ratio = 1.0 / eta; n = -no; } else { cos_to = dot(no, wo); ratio = eta; n = no; }
    const double sin2_to = 1.0 - cos_to * cos_to;
    const double sin2_ti = sin2_to / powi(ratio, 2);
    if (sin2_ti >= 1.0) return false;   // unreachable!() in the reference (TIR is sampled as reflection)
    const double cos_ti = sqrt(fmax(1.0 - sin2_ti, 0.0));
    wi = (-wo) / ratio + (cos_to / ratio - cos_ti) * n;
    return !same_hemisphere(wi, wo);
}
__device__ __forceinline__ C4 reflect_coeff(const DevScene& S, const Mat& m, D3 wo, D3 wi, const Lam& l) {                      // bxdf/microfacet.rs:46-62
    const D3 wh = normalize(wi + wo);
    const double d = ggx_d(m, wh); const C4 f = fresnel4(S, m, wo, wh, l); const double g = ggx_g(m, wo, wi, wh);
    return d * f * g / (4.0 * fabs(wo.z) * fabs(wi.z));
}
__device__ __forceinline__ double lambertian_pdf(D3 wo, D3 wi) {                                                                // bxdf/scatter.rs:14-25
    if (!same_hemisphere(wo, wi)) return 0.0;
    return wi.z > 0.0 ? wi.z / LUMO_PI : 0.0;
}
__device__ __forceinline__ double refl_pdf_half(const Mat& m, D3 wo, D3 wh) {
    if (mf_is_delta(m)) return (1.0 - wh.z < LUMO_EPS) ? 1.0 : 0.0;
    return sample_normal_pdf(m, wh, wo) / (4.0 * fabs(dot(wo, wh)));
}

// BxDF::f (bxdf.rs:69-106), local frame
// K = material kind known at compile time (the per-kind shade kernels), or -1 for a runtime switch.
template <int K>
__device__ __noinline__ C4 bx_f(const DevScene& S, const Mat& m, D3 wo, D3 wi, const Lam& l, bool reflection, bool backface, int mode, double u, double v) {
    const uint32_t kind = LUMO_KIND(K, m);
    if ((!reflection || backface) && kind != LMAT_MFDIELECTRIC) return c4(0.0);
    switch (kind) {
    case LMAT_LAMBERTIAN: return spec4(m.kd, l) / LUMO_PI;
    case LMAT_MFDIFFUSE: {                                                                                                      // bxdf/microfacet.rs:131-156
        const D3 wh = normalize(wo + wi);
        const double d = ggx_d(m, wh); const C4 f = fresnel4(S, m, wo, wh, l); const double g = ggx_g(m, wo, wi, wh);
        const C4 fr = d * f * g / (4.0 * fabs(wo.z) * fabs(wi.z));
        const double fd = disney_diffuse(m, wo.z, wi.z, wh.z);
        return fr * tex4<LUMO_TEX(K)>(S, m.ks, m.ks_tex, u, v, l) + tex4<LUMO_TEX(K)>(S, m.kd, m.kd_tex, u, v, l) * (c4(1.0) - f) * fd / LUMO_PI;
    }
    case LMAT_MFCONDUCTOR: {                                                                                                    // bxdf/microfacet.rs:71-85
        const C4 ks = tex4<LUMO_TEX(K)>(S, m.ks, m.ks_tex, u, v, l);
        if (mf_is_delta(m)) return ks * fresnel4(S, m, wo, d3(0, 0, 1), l) / fabs(wi.z);
        return ks * reflect_coeff(S, m, wo, wi, l);
    }
    case LMAT_MFDIELECTRIC: {                                                                                                   // bxdf/microfacet.rs:222-283
        const double e = eta_at(S, m, l.l[0]);
        const double ratio = reflection ? 1.0 : (wo.z < 0.0 ? 1.0 / e : e);
        const bool flat = e == 1.0 || mf_is_delta(m);
        D3 wh = flat ? d3(0, 0, 1) : normalize(wi * ratio + wo);
        if (reflection) {
            const C4 ks = tex4<LUMO_TEX(K)>(S, m.ks, m.ks_tex, u, v, l);
            if (flat) return ks * fresnel4(S, m, wo, wh, l) / fabs(wi.z);
            return ks * reflect_coeff(S, m, wo, wi, l);
        }
        const C4 f = fresnel4(S, m, wo, wh, l);
        if (wh.z < 0.0) wh = -wh;
        const double scale = mode == 0 ? ratio * ratio : 1.0;
        const C4 tf = tex4<LUMO_TEX(K)>(S, m.tf, m.tf_tex, u, v, l);
        if (flat) return tf * (c4(1.0) - f) / (scale * fabs(wi.z));
        const double d = ggx_d(m, wh), g = ggx_g(m, wo, wi, wh);
        const double hwo = dot(wh, wo), hwi = dot(wh, wi);
        return tf * d * (c4(1.0) - f) * g / scale * fabs(hwi * hwo / (wi.z * wo.z)) / powi(ratio * hwi + hwo, 2);
    }
    default: return c4(0.0);
    }
}
// BxDF::sample (bxdf.rs:108-131); may terminate the secondary wavelengths (dispersion)
template <int K>
__device__ __noinline__ bool bx_sample(const DevScene& S, const Mat& m, D3 wo, bool backface, Lam& l, double ru, double r0, double r1, D3& wi) {
    const uint32_t kind = LUMO_KIND(K, m);
    if (backface && kind != LMAT_MFDIELECTRIC) return false;
    switch (kind) {
    case LMAT_LAMBERTIAN: wi = square_to_cos_hemisphere(r0, r1); return true;
    case LMAT_MFDIFFUSE: {                                                                                                      // bxdf/microfacet.rs:159-177
        const double pr = f_schlick(0.04, 1.0, wo.z), ps = 1.0 - pr;
        if (ru < pr / (pr + ps)) { const D3 wh = mf_is_delta(m) ? d3(0, 0, 1) : sample_normal(m, wo, r0, r1); return reflect(wo, wh, wi); }
        wi = square_to_cos_hemisphere(r0, r1); return true;
    }
    case LMAT_MFCONDUCTOR: {                                                                                                    // bxdf/microfacet.rs:87-99
        if (mf_is_delta(m)) { wi = d3(-wo.x, -wo.y, wo.z); return true; }
        return reflect(wo, sample_normal(m, wo, r0, r1), wi);
    }
    case LMAT_MFDIELECTRIC: {                                                                                                   // bxdf/microfacet.rs:285-311
        if (!(m.flags & LMF_ETA_CONST)) { l.l[1] = 0.0; l.l[2] = 0.0; l.l[3] = 0.0; }
        const double wl = l.l[0];
        const double e = eta_at(S, m, wl);
        const D3 wh = (e == 1.0 || mf_is_delta(m)) ? d3(0, 0, 1) : sample_normal(m, wo, r0, r1);
        const double pr = fresnel_at(S, m, wo, wh, wl), pt = 1.0 - pr;
        if (ru < pr / (pr + pt)) return reflect(wo, wh, wi);
        return refract(e, wo, wh, wi);
    }
    default: return false;
    }
}
// BxDF::pdf (bxdf.rs:133-151)
template <int K>
__device__ __noinline__ double bx_pdf(const DevScene& S, const Mat& m, D3 wo, D3 wi, bool reflection, const Lam& l) {
    const uint32_t kind = LUMO_KIND(K, m);
    if (!reflection && kind != LMAT_MFDIELECTRIC) return 0.0;
    switch (kind) {
    case LMAT_LAMBERTIAN: return lambertian_pdf(wo, wi);
    case LMAT_MFDIFFUSE: {                                                                                                      // bxdf/microfacet.rs:180-205
        if (!same_hemisphere(wi, wo)) return 0.0;
        const D3 wh = normalize(wo + wi);
        const double pr = f_schlick(0.04, 1.0, wo.z), ps = 1.0 - pr;
        return pr * refl_pdf_half(m, wo, wh) + ps * lambertian_pdf(wo, wi);
    }
    case LMAT_MFCONDUCTOR: {                                                                                                    // bxdf/microfacet.rs:101-120
        if (!same_hemisphere(wi, wo)) return 0.0;
        D3 wh = normalize(wo + wi);
        if (wh.z < 0.0) wh = -wh;
        return refl_pdf_half(m, wo, wh);
    }
    case LMAT_MFDIELECTRIC: {                                                                                                   // bxdf/microfacet.rs:313-373
        const double wl = l.l[0];
        const double e = eta_at(S, m, wl);
        const double ratio = reflection ? 1.0 : (wo.z < 0.0 ? 1.0 / e : e);
        D3 wh = e == 1.0 ? d3(0, 0, 1) : normalize(wo + wi * ratio);
        if (wh.z < 0.0) wh = -wh;
        const double hwo = dot(wo, wh), hwi = dot(wi, wh);
        if (hwo == 0.0 || hwi == 0.0) return 0.0;
        if (hwo * wo.z < 0.0 || hwi * wi.z < 0.0) return 0.0;
        const double pr = fresnel_at(S, m, wo, wh, wl), pt = 1.0 - pr;
        const bool flat = e == 1.0 || mf_is_delta(m);
        if (reflection && flat) return (1.0 - wh.z < LUMO_EPS) ? pr / (pr + pt) : 0.0;
        if (reflection) return sample_normal_pdf(m, wh, wo) / (4.0 * fabs(hwo)) * pr / (pr + pt);
        if (flat) return (1.0 - wh.z < LUMO_EPS) ? pt / (pr + pt) : 0.0;
        return sample_normal_pdf(m, wh, wo) * fabs(hwi) / powi(hwi + hwo / ratio, 2) * pt / (pr + pt);
    }
    default: return 0.0;
    }
}
// BSDF wrappers (bsdf.rs:28-90) + Material dispatch (material.rs:245-312).  The shading frame is a
// pure function of the shading normal (bsdf.rs:41-46 rebuilds it per call); callers that evaluate the
// BSDF several times at one hit pass the frame in.
__device__ __forceinline__ bool is_reflection(D3 wo, D3 wi, D3 ng) { return dot(ng, wi) * dot(ng, wo) >= 0.0; }
template <int K>
__device__ __forceinline__ C4 bsdf_f(const DevScene& S, const Mat& m, const Onb& uvw, D3 wo, D3 wi, const Lam& l, int mode, const DevHit& h) {
    if (K < 0 && !mat_is_standard(m)) return c4(0.0);
    return bx_f<K>(S, m, to_local(uvw, wo), to_local(uvw, wi), l, is_reflection(wo, wi, h.ng), h.backface, mode, h.u, h.v);
}
template <int K>
__device__ __forceinline__ bool bsdf_sample(const DevScene& S, const Mat& m, const Onb& uvw, D3 wo, const DevHit& h, Lam& l, double ru, double r0, double r1, D3& wi) {
    if (K < 0 && !mat_is_standard(m)) return false;
    D3 wl;
    if (!bx_sample<K>(S, m, to_local(uvw, wo), h.backface, l, ru, r0, r1, wl)) return false;
    wi = to_world(uvw, wl);
    return true;
}
template <int K>
__device__ __forceinline__ double bsdf_pdf(const DevScene& S, const Mat& m, const Onb& uvw, D3 wo, D3 wi, const DevHit& h, const Lam& l, bool swap_dir) {
    if (swap_dir) { const D3 t = wo; wo = wi; wi = t; }
    if (K < 0 && !mat_is_standard(m)) return 0.0;
    return bx_pdf<K>(S, m, to_local(uvw, wo), to_local(uvw, wi), is_reflection(wo, wi, h.ng), l);
}

// ---- lights: Sampleable (object.rs:98-157, rectangle.rs:107-133, triangle.rs:207-241, sphere.rs:104-207, instance.rs:132-199)
__device__ __forceinline__ D3 rect_vec(const double* p) { return d3(p[0], p[1], p[2]); }
__device__ __forceinline__ double base_area(const DevScene& S, const LumoObject& o) {
    if (o.kind == LOBJ_RECT) { const LumoRect* R = S.rects + o.rect; return fabs(length(cross(rect_vec(R->b0), rect_vec(R->b1)))); }
    if (o.kind == LOBJ_SPHERE) { const double r = S.spheres[o.geom].radius; return 4.0 * LUMO_PI * r * r; }
    D3 A, B, C; load_tri(S.tri_verts + o.geom, A, B, C);
    return length(cross(B - A, C - A)) / 2.0;
}
// sample_on of the un-instanced object: point, normals, fp error (returns a DevHit with t = 0)
__device__ __forceinline__ DevHit base_sample_on(const DevScene& S, const LumoObject& o, double r0, double r1) {
    DevHit h; h.t = 0.0; h.u = 0.0; h.v = 0.0; h.material = o.material; h.backface = false;
    if (o.kind == LOBJ_RECT) {
        const LumoRect* R = S.rects + o.rect;
        const D3 org = rect_vec(R->origin), b0 = rect_vec(R->b0), b1 = rect_vec(R->b1);
        h.p = org + r0 * b0 + r1 * b1;
        h.ng = normalize(cross(b0, b1)); h.ns = h.ng;
        h.fp_error = gamma_n(4) * (vabs(org) + vabs(r0 * b0) + vabs(r1 * b1));
    } else if (o.kind == LOBJ_SPHERE) {
        const double radius = S.spheres[o.geom].radius;
        D3 xo = radius * square_to_sphere(r0, r1);
        xo = xo * radius / length(xo);
        h.fp_error = vabs(xo) * gamma_n(5);
        h.ng = xo / radius; h.ns = h.ng; h.p = xo;
    } else {
        D3 A, B, C; load_tri(S.tri_verts + o.geom, A, B, C);
        const double ga = 1.0 - sqrt(1.0 - r0), be = r1 * (1.0 - ga), al = 1.0 - ga - be;
        const D3 bma = B - A, cma = C - A;
        h.ng = normalize(cross(bma, cma));
        h.ns = h.ng;
        const LumoTriShade sh = S.tri_shade[o.geom];
        if (sh.flags & 1u) {
            const double* n = S.normals;
            h.ns = normalize(al * d3(n[3 * sh.n[0]], n[3 * sh.n[0] + 1], n[3 * sh.n[0] + 2]) + be * d3(n[3 * sh.n[1]], n[3 * sh.n[1] + 1], n[3 * sh.n[1] + 2])
                             + ga * d3(n[3 * sh.n[2]], n[3 * sh.n[2] + 1], n[3 * sh.n[2] + 2]));
        }
        h.p = A + be * bma + ga * cma;
        h.fp_error = gamma_n(6) * (vabs(A) + vabs(be * bma) + vabs(ga * cma));
    }
    // Hit::new(.., wo = -ng, ..): backface = (-ng).ng > 0 = false
    return h;
}
__device__ __forceinline__ D3 base_sample_towards(const DevScene& S, const LumoObject& o, D3 xo, double r0, double r1) {
    if (o.kind == LOBJ_SPHERE) {                                                                                                // sphere.rs:136-186
        const double radius = S.spheres[o.geom].radius;
        const double d2 = dot(xo, xo), rr2 = radius * radius;
        D3 xi;
        if (d2 < rr2) xi = base_sample_on(S, o, r0, r1).p;
        else {
            const Onb uvw = onb_new(-normalize(xo));
            const double d = sqrt(d2);
            const double sin2_max = rr2 / d2;
            const double cos_max = sqrt(fmax(1.0 - sin2_max, 0.0));
            const double cos_t = (1.0 - r0) + r0 * cos_max;
            const double sin_t = sqrt(fmax(1.0 - cos_t * cos_t, 0.0));
            const double phi = 2.0 * LUMO_PI * r1;
            const double ds = d * cos_t - sqrt(fmax(rr2 - d2 * sin_t * sin_t, 0.0));
            const double cos_a = (d2 + rr2 - ds * ds) / (2.0 * d * radius);
            const double sin_a = sqrt(fmax(1.0 - cos_a * cos_a, 0.0));
            double sn, cs; lm_sincos(phi, &sn, &cs);
            const D3 ngl = d3(cs * sin_a, sn * sin_a, cos_a);
            xi = normalize(to_world(uvw, -ngl)) * radius;
        }
        return normalize(xi - xo);
    }
    return normalize(base_sample_on(S, o, r0, r1).p - xo);                                                                     // object.rs:136-139
}
__device__ __forceinline__ double base_sample_towards_pdf(const DevScene& S, const LumoObject& o, const Ray& ri, D3 xi, D3 ng) {
    if (o.kind == LOBJ_SPHERE) {                                                                                                // sphere.rs:190-207
        const double radius = S.spheres[o.geom].radius;
        const double rr2 = radius * radius, d2 = dot(ri.o, ri.o);
        if (d2 < rr2) return (1.0 / base_area(S, o)) * dist2(ri.o, xi) / fabs(dot(ng, ri.d));
        const double cos_max = sqrt(fmax(1.0 - rr2 / d2, 0.0));
        return 1.0 / (2.0 * LUMO_PI * (1.0 - cos_max));
    }
    return (1.0 / base_area(S, o)) * dist2(ri.o, xi) / fabs(dot(ng, ri.d));                                                    // object.rs:148-156
}
__device__ __noinline__ D3 light_sample_towards(const DevScene& S, const LumoObject& o, D3 xo, double r0, double r1) {
    if (o.inst < 0) return base_sample_towards(S, o, xo, r0, r1);
    const LumoInstance* I = S.instances + o.inst;                                                                               // instance.rs:163-168
    const D3 dl = base_sample_towards(S, o, xf_point(I->inv, xo), r0, r1);
    return normalize(xf_dir(I->m, dl));
}
__device__ __forceinline__ double det33(const double* m) {   // Mat3::det on the 3x3 of a 3x4 (mat3.rs:41-52)
    const double pos = m[0] * m[5] * m[10] + m[1] * m[6] * m[8] + m[2] * m[4] * m[9];
    const double ng = m[2] * m[5] * m[8] + m[1] * m[4] * m[10] + m[0] * m[6] * m[9];
    return pos - ng;
}
__device__ __noinline__ double light_sample_towards_pdf(const DevScene& S, const LumoObject& o, const Ray& ri, D3 xi, D3 ng) {
    if (o.inst < 0) return base_sample_towards_pdf(S, o, ri, xi, ng);
    const LumoInstance* I = S.instances + o.inst;                                                                               // instance.rs:170-199
    // normal_transform.inv().transpose() (Mat3::inv, mat3.rs:64-72)
    const double* n = I->nrm;
    const D3 y0 = d3(n[0], n[1], n[2]), y1 = d3(n[3], n[4], n[5]), y2 = d3(n[6], n[7], n[8]);
    const double pos = y0.x * y1.y * y2.z + y0.y * y1.z * y2.x + y0.z * y1.x * y2.y;
    const double neg = y0.z * y1.y * y2.x + y0.y * y1.x * y2.z + y0.x * y1.z * y2.y;
    const double inv_det = 1.0 / (pos - neg);
    // inv() = transpose(rows r0,r1,r2); inv().transpose() = rows r0,r1,r2
    const D3 r0 = cross(y1, y2) * inv_det, r1 = cross(y2, y0) * inv_det, r2 = cross(y0, y1) * inv_det;
    const D3 ng_local = normalize(d3(dot(r0, ng), dot(r1, ng), dot(r2, ng)));
    const D3 xi_local = xf_point(I->inv, xi);
    Ray rl; rl.o = xf_point(I->inv, ri.o); rl.d = normalize(xf_dir(I->inv, ri.d));
    const double pdf_local = base_sample_towards_pdf(S, o, rl, xi_local, ng_local);
    const double height = fabs(dot(ng, xf_dir(I->m, ng_local)));
    const double volume = fabs(det33(I->m));
    const double jacobian = volume / height;
    const double sa_conv = dist2(ri.o, xi) * fabs(dot(rl.d, ng_local)) / (dist2(rl.o, xi_local) * fabs(dot(ri.d, ng)));
    return pdf_local * sa_conv / jacobian;
}
// Sampleable::sample_on through an optional Instance (instance.rs:148-161)
__device__ __noinline__ DevHit light_sample_on(const DevScene& S, const LumoObject& o, double r0, double r1) {
    DevHit h = base_sample_on(S, o, r0, r1);
    if (o.inst >= 0) {
        const LumoInstance* I = S.instances + o.inst;
        h.ng = normalize(mul33(I->nrm, h.ng));
        h.ns = normalize(mul33(I->nrm, h.ns));
        h.p = xf_point(I->m, h.p);
        h.fp_error = propagate_fp_err(I, h.p, h.fp_error);
    }
    return h;
}
// light.hit(r, 0, INF) for one light object: Object::hit + Hit reconstruction
template <bool UV = true>
__device__ __noinline__ bool light_hit(const DevScene& S, uint32_t obj_index, const Ray& r, DevHit& out) {
    HitRec rec;
    RayCtx w; make_ctx(r, w);
    if (!object_hit<false, LUMO_LIGHT_KD_STACK>(S, S.objects[obj_index], w, 0.0, LUMO_INF, rec, nullptr)) return false;
    rec.obj = obj_index;
    out = reconstruct_hit<UV>(S, r, rec);
    return true;
}
// BVH::sample_light (bvh.rs:67-77): alias table
__device__ __forceinline__ uint32_t sample_light(const DevScene& S, double u) {
    const double ru = u * (double)S.P.n_lights;
    const unsigned long long idx = sat_u64(floor(ru));
    const double fr = fractd(ru);
    const LumoLight L = S.lights[idx];
    return fr < L.alias_prob ? (uint32_t)idx : L.alias;
}

// ---- camera (camera.rs) -----------------------------------------------------------------------------
__device__ __forceinline__ D3 xf4_point(const double* m, D3 p) {   // Mat4 * (p,1) then project (mat4.rs:46-52,172-179)
    const double x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3] * 1.0, y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7] * 1.0;
    const double z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11] * 1.0, w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15] * 1.0;
    if (w == 0.0) return d3(x, y, z);
    return d3(x / w, y / w, z / w);
}
__device__ __forceinline__ D3 xf4_dir(const double* m, D3 p) {
    const double x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3] * 0.0, y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7] * 0.0;
    const double z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11] * 0.0, w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15] * 0.0;
    if (w == 0.0) return d3(x, y, z);
    return d3(x / w, y / w, z / w);
}
__device__ __noinline__ Ray camera_generate_ray(const LumoCamera& C, double rx, double ry, double l0, double l1) {           // camera.rs:221-268
    const D3 cam = xf4_point(C.camera_to_screen_inv, xf4_point(C.screen_to_raster_inv, d3(rx, ry, 0.0)));
    D3 xo_local, wi_local;
    if (!C.ortho) { xo_local = d3(0, 0, 0); wi_local = normalize(cam); } else { xo_local = cam; wi_local = d3(0, 0, 1); }
    if (C.lens_radius != 0.0) {
        double dx, dy; square_to_disk(l0, l1, dx, dy);
        const D3 lens = d3(C.lens_radius * dx, C.lens_radius * dy, 0.0);
        const double fd = C.focal_length / wi_local.z;
        const D3 focus = fd * wi_local;
        xo_local = xo_local + lens; wi_local = focus - lens;
    }
    Ray r; r.o = xf4_point(C.world_to_camera_inv, xo_local); r.d = normalize(xf4_dir(C.world_to_camera_inv, wi_local));
    return r;
}

// ---- film (film/tile.rs:65-111, filter.rs:82-102, tone_mapping.rs:38-63) -----------------------------
__device__ __forceinline__ double gauss(double x, double sigma) { return lm_exp(-powi(x, 2) / (2.0 * sigma * sigma)) / sqrt(fmax(2.0 * LUMO_PI * sigma * sigma, 0.0)); }
__device__ __forceinline__ double mitch(double x, double b, double c) {
    x = fabs(x);
    double p = 0.0;
    if (x < 1.0) p = (12.0 - 9.0 * b - 6.0 * c) * powi(x, 3) + (-18.0 + 12.0 * b + 6.0 * c) * powi(x, 2) + (6.0 - 2.0 * b);
    else if (x < 2.0) p = (-b - 6.0 * c) * powi(x, 3) + (6.0 * b + 30.0 * c) * powi(x, 2) + (-12.0 * b - 48.0 * c) * x + (8.0 * b + 24.0 * c);
    return p / 6.0;
}
// PixelFilter::eval (filter.rs:82-102) is a product of the same 1-D profile in x and y for every filter kind; the
// per-axis factors of a sample's footprint are evaluated once per row / column instead of once per pixel.
__device__ __forceinline__ double filter_1d(const LumoFilm& F, double v, double gr) {
    switch (F.filter_kind) {
    case 0: return fabs(v) < F.filter_r ? 1.0 : 0.0;
    case 1: return fmax(F.filter_r - fabs(v), 0.0);
    case 2: return fmax(gauss(v, F.filter_p) - gr, 0.0);
    default: return mitch(2.0 * v / F.filter_r, F.filter_p, (1.0 - F.filter_p) / 2.0);
    }
}
__device__ __forceinline__ double filter_eval(const LumoFilm& F, double x, double y) {
    const double gr = F.filter_kind == 2 ? gauss(F.filter_r, F.filter_p) : 0.0;
    return filter_1d(F, x, gr) * filter_1d(F, y, gr);
}
__device__ __forceinline__ C4 tone_map(const DevScene& S, int kind, double arg, C4 c, const Lam& l) {
    if (kind == 1) { C4 r; for (int i = 0; i < 4; i++) r.s[i] = clampd(c.s[i], 0.0, arg); return r; }
    if (kind == 2) return c / (1.0 + luminance(S, c, l));
    return c;
}
// FilmTile::add_sample for a main (non-splat) or splat sample, accumulating straight into the
// full-frame buffers with f64 atomics.  Non-splat footprints are clamped to the sample's 16x16 tile
// exactly like the reference's per-tile FilmTile (tile.rs:74-84), so the result equals
// Film::add_tile of all tiles.
__device__ __noinline__ void film_add_sample(const DevScene& S, double* pixels, double* splats, C4 color, const Lam& l, double rx, double ry, bool splat) {
    const LumoFilm& F = S.P.film;
    const D3 xyz = color_xyz(S, color, l);
    const D3 rgb = mul33(F.xyz_to_rgb, mul33(F.wb, xyz));
    const unsigned long long W = S.P.camera.res_x, H = S.P.camera.res_y;
    const unsigned long long px = sat_u64(floor(rx)), py = sat_u64(floor(ry)), r = F.r_disc;
    unsigned long long mi_x = px > r ? px - r : 0, mi_y = py > r ? py - r : 0, mx_x, mx_y;
    if (splat) { mx_x = min(px + r, W - 1); mx_y = min(py + r, H - 1); }
    else {
        const unsigned long long tx0 = (px / 16) * 16, ty0 = (py / 16) * 16;
        const unsigned long long tx1 = min(tx0 + 16, W), ty1 = min(ty0 + 16, H);
        mi_x = max(mi_x, tx0); mi_y = max(mi_y, ty0);
        mx_x = min(px + r, tx1 - 1); mx_y = min(py + r, ty1 - 1);
    }
    const double gr = F.filter_kind == 2 ? gauss(F.filter_r, F.filter_p) : 0.0;
    double wx[8];                                                     // column factors of the footprint (r_disc <= 3 in practice)
    const bool cached = mx_x - mi_x < 8ull;
    if (cached) for (unsigned long long fx = mi_x; fx <= mx_x; fx++) wx[fx - mi_x] = filter_1d(F, rx - (0.5 + (double)fx), gr);
    for (unsigned long long fy = mi_y; fy <= mx_y; fy++) {
        const double wy = filter_1d(F, ry - (0.5 + (double)fy), gr);
        for (unsigned long long fx = mi_x; fx <= mx_x; fx++) {
            const double w = (cached ? wx[fx - mi_x] : filter_1d(F, rx - (0.5 + (double)fx), gr)) * wy;
            if (w != 0.0) {
                const unsigned long long idx = fx + fy * W;
                if (splat) { atomicAdd(splats + 3 * idx, rgb.x * w); atomicAdd(splats + 3 * idx + 1, rgb.y * w); atomicAdd(splats + 3 * idx + 2, rgb.z * w); }
                else { atomicAdd(pixels + 4 * idx, rgb.x * w); atomicAdd(pixels + 4 * idx + 1, rgb.y * w); atomicAdd(pixels + 4 * idx + 2, rgb.z * w); atomicAdd(pixels + 4 * idx + 3, w); }
            }
        }
    }
}

}  // namespace lumo_dev
