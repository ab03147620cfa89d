// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of ekarpp/lumo (v0.6.1) math primitives.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, link or execute anything under oracle/.  The product (lumo_b200/) never includes this.
//
// PARITY UNPINNED: the reference ships no golden vectors for traversal / intersection /
// integrators and cannot be compiled here (no Rust toolchain), so
// this restatement is pinned only by following the reference op-for-op (citations inline) and by
// re-asserting the reference's own property tests (tests/test_oracle_*.py).
//
// Build flags matter: -ffp-contract=off -fno-fast-math (Rust never contracts a*b+c).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>
// sin / cos / atan2 / acos / atanh / cosh / exp: one FMA-free source shared with the device kernels and the host builder, so
// that identical inputs give identical bits on every side (see the header of that file; the reference's own libm is not
// pinned to the bit by anything it ships)
#include "../lumo_b200/csrc/common/lumo_math.h"

namespace oracle {

typedef double Float;
static const Float INF = std::numeric_limits<double>::infinity();
static const Float EPSILON = 1e-10;          // src/lib.rs:67
static const Float PI = 3.14159265358979323846264338327950288;  // std::f64::consts::PI

// Rust f64::min/max ignore a NaN operand == fmin/fmax
static inline Float fmin_(Float a, Float b) { return std::fmin(a, b); }
static inline Float fmax_(Float a, Float b) { return std::fmax(a, b); }
// f64::signum: +1 for +0, -1 for -0, NaN for NaN
static inline Float signum(Float x) { return std::isnan(x) ? x : std::copysign(1.0, x); }
// f64::clamp
static inline Float clampf(Float v, Float lo, Float hi) { Float r = v; if (r < lo) r = lo; if (r > hi) r = hi; return r; }
// `x as u64` saturating cast (NaN -> 0, negative -> 0, huge -> MAX)
static inline uint64_t sat_u64(Float x) {
    if (!(x > 0.0)) return 0;
    if (x >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)x;
}
// powi(n): compiler-rt square-and-multiply (__powidf2)
static inline Float powi(Float a, int b) {
    const bool recip = b < 0;
    Float r = 1.0;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}
static inline Float fract(Float x) { return x - std::trunc(x); }

struct Vec2 {
    Float x, y;
    Vec2() : x(0), y(0) {}
    Vec2(Float x_, Float y_) : x(x_), y(y_) {}
};
static inline Vec2 operator+(Vec2 a, Vec2 b) { return Vec2(a.x + b.x, a.y + b.y); }
static inline Vec2 operator-(Vec2 a, Vec2 b) { return Vec2(a.x - b.x, a.y - b.y); }
static inline Vec2 operator*(Vec2 a, Vec2 b) { return Vec2(a.x * b.x, a.y * b.y); }
static inline Vec2 operator*(Float s, Vec2 a) { return Vec2(s * a.x, s * a.y); }
static inline Vec2 operator*(Vec2 a, Float s) { return Vec2(a.x * s, a.y * s); }
static inline Vec2 operator/(Vec2 a, Float s) { return Vec2(a.x / s, a.y / s); }
static inline Vec2 operator+(Float s, Vec2 a) { return Vec2(s + a.x, s + a.y); }

// src/math/vec3.rs
struct Vec3 {
    Float x, y, z;
    Vec3() : x(0), y(0), z(0) {}
    Vec3(Float x_, Float y_, Float z_) : x(x_), y(y_), z(z_) {}
    static Vec3 splat(Float v) { return Vec3(v, v, v); }
    Float axis(int a) const { return a == 0 ? x : (a == 1 ? y : z); }
    Float dot(Vec3 r) const { return x * r.x + y * r.y + z * r.z; }           // vec3.rs:137-139
    Float length_squared() const { return dot(*this); }
    Float length() const { return std::sqrt(fmax_(length_squared(), 0.0)); }    // vec3.rs:91-93
    Vec3 cross(Vec3 r) const {                                                  // vec3.rs:128-134
        return Vec3(y * r.z - z * r.y, z * r.x - x * r.z, x * r.y - y * r.x);
    }
    Vec3 abs() const { return Vec3(std::fabs(x), std::fabs(y), std::fabs(z)); }
    Vec3 min(Vec3 o) const { return Vec3(fmin_(x, o.x), fmin_(y, o.y), fmin_(z, o.z)); }
    Vec3 max(Vec3 o) const { return Vec3(fmax_(x, o.x), fmax_(y, o.y), fmax_(z, o.z)); }
    Float min_element() const { return fmin_(x, fmin_(y, z)); }                 // vec3.rs:152-154
    Float max_element() const { return fmax_(x, fmax_(y, z)); }                 // vec3.rs:157-159
    Float distance_squared(Vec3 o) const;
    Float distance(Vec3 o) const;
    Vec3 normalize() const;
    Vec3 scale(Float s) const { return Vec3(x * s, y * s, z * s); }
    Vec3 floor() const { return Vec3(std::floor(x), std::floor(y), std::floor(z)); }
};
static inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline Vec3 operator*(Vec3 a, Vec3 b) { return Vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline Vec3 operator/(Vec3 a, Vec3 b) { return Vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline Vec3 operator*(Vec3 a, Float s) { return Vec3(a.x * s, a.y * s, a.z * s); }
static inline Vec3 operator*(Float s, Vec3 a) { return Vec3(s * a.x, s * a.y, s * a.z); }
static inline Vec3 operator/(Vec3 a, Float s) { return Vec3(a.x / s, a.y / s, a.z / s); }
static inline Vec3 operator/(Float s, Vec3 a) { return Vec3(s / a.x, s / a.y, s / a.z); }
static inline Vec3 operator+(Float s, Vec3 a) { return Vec3(s + a.x, s + a.y, s + a.z); }
static inline Vec3 operator-(Vec3 a) { return Vec3(-a.x, -a.y, -a.z); }
inline Float Vec3::distance_squared(Vec3 o) const { return (*this - o).length_squared(); }
inline Float Vec3::distance(Vec3 o) const { return std::sqrt(fmax_(distance_squared(o), 0.0)); }
inline Vec3 Vec3::normalize() const { return *this / length(); }               // vec3.rs:147-149

struct Vec4 {
    Float x, y, z, w;
    Vec4() : x(0), y(0), z(0), w(0) {}
    Vec4(Float a, Float b, Float c, Float d) : x(a), y(b), z(c), w(d) {}
    Float dot(const Vec4& o) const { return x * o.x + y * o.y + z * o.z + w * o.w; }  // mat4.rs:62-64
    Vec3 truncate() const { return Vec3(x, y, z); }
    Vec3 project() const { return w == 0.0 ? truncate() : truncate() / w; }           // mat4.rs:46-52
    Vec4 abs() const { return Vec4(std::fabs(x), std::fabs(y), std::fabs(z), std::fabs(w)); }
};
static inline Vec4 operator+(Vec4 a, Vec4 b) { return Vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline Vec4 operator*(Vec4 a, Float s) { return Vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
static inline Vec4 extend(Vec3 v, Float w) { return Vec4(v.x, v.y, v.z, w); }
static const Vec4 V4X(1, 0, 0, 0), V4Y(0, 1, 0, 0), V4Z(0, 0, 1, 0), V4W(0, 0, 0, 1);

// src/math/mat3.rs (row major)
struct Mat3 {
    Vec3 y0, y1, y2;
    Mat3() : y0(1, 0, 0), y1(0, 1, 0), y2(0, 0, 1) {}
    Mat3(Vec3 a, Vec3 b, Vec3 c) : y0(a), y1(b), y2(c) {}
    static Mat3 diag(Vec3 d) { return Mat3(Vec3(d.x, 0, 0), Vec3(0, d.y, 0), Vec3(0, 0, d.z)); }
    Float det() const {                                                               // mat3.rs:41-52
        Float pos = y0.x * y1.y * y2.z + y0.y * y1.z * y2.x + y0.z * y1.x * y2.y;
        Float neg = y0.z * y1.y * y2.x + y0.y * y1.x * y2.z + y0.x * y1.z * y2.y;
        return pos - neg;
    }
    Mat3 transpose() const {
        return Mat3(Vec3(y0.x, y1.x, y2.x), Vec3(y0.y, y1.y, y2.y), Vec3(y0.z, y1.z, y2.z));
    }
    Mat3 inv() const {                                                                // mat3.rs:64-72
        Float inv_det = 1.0 / det();
        return Mat3(y1.cross(y2).scale(inv_det), y2.cross(y0).scale(inv_det), y0.cross(y1).scale(inv_det)).transpose();
    }
    Vec3 mul_vec3(Vec3 r) const { return Vec3(y0.dot(r), y1.dot(r), y2.dot(r)); }
    Mat3 mul_mat3(const Mat3& rhs) const {
        Mat3 t = rhs.transpose();
        return Mat3(Vec3(y0.dot(t.y0), y0.dot(t.y1), y0.dot(t.y2)),
                    Vec3(y1.dot(t.y0), y1.dot(t.y1), y1.dot(t.y2)),
                    Vec3(y2.dot(t.y0), y2.dot(t.y1), y2.dot(t.y2)));
    }
};

// src/math/mat4.rs
struct Mat4 {
    Vec4 y0, y1, y2, y3;
    Mat4() : y0(V4X), y1(V4Y), y2(V4Z), y3(V4W) {}
    Mat4(Vec4 a, Vec4 b, Vec4 c, Vec4 d) : y0(a), y1(b), y2(c), y3(d) {}
    static Mat4 mat3(const Mat3& m) { return Mat4(extend(m.y0, 0), extend(m.y1, 0), extend(m.y2, 0), V4W); }
    Mat3 to_mat3() const { return Mat3(y0.truncate(), y1.truncate(), y2.truncate()); }
    Mat4 abs() const { return Mat4(y0.abs(), y1.abs(), y2.abs(), y3.abs()); }
    Mat4 transpose() const {
        return Mat4(Vec4(y0.x, y1.x, y2.x, y3.x), Vec4(y0.y, y1.y, y2.y, y3.y),
                    Vec4(y0.z, y1.z, y2.z, y3.z), Vec4(y0.w, y1.w, y2.w, y3.w));
    }
    Vec4 mul_vec4(const Vec4& r) const { return Vec4(y0.dot(r), y1.dot(r), y2.dot(r), y3.dot(r)); }
    Mat4 mul(const Mat4& rhs) const {                                                 // mat4.rs:182-213
        Mat4 t = rhs.transpose();
        return Mat4(Vec4(y0.dot(t.y0), y0.dot(t.y1), y0.dot(t.y2), y0.dot(t.y3)),
                    Vec4(y1.dot(t.y0), y1.dot(t.y1), y1.dot(t.y2), y1.dot(t.y3)),
                    Vec4(y2.dot(t.y0), y2.dot(t.y1), y2.dot(t.y2), y2.dot(t.y3)),
                    Vec4(y3.dot(t.y0), y3.dot(t.y1), y3.dot(t.y2), y3.dot(t.y3)));
    }
};

// src/math/transform.rs
struct Transform {
    Mat4 m, inv;
    Transform() {}
    Transform(const Mat4& m_, const Mat4& i_) : m(m_), inv(i_) {}
    static Transform mat3(const Mat3& m3) { return Transform(Mat4::mat3(m3), Mat4::mat3(m3.inv())); }
    static Transform translation(Float x, Float y, Float z) {                         // transform.rs:133-149
        return Transform(Mat4(V4X + V4W * x, V4Y + V4W * y, V4Z + V4W * z, V4W),
                         Mat4(V4X + V4W * (-x), V4Y + V4W * (-y), V4Z + V4W * (-z), V4W));
    }
    static Transform scale(Float x, Float y, Float z) { return mat3(Mat3::diag(Vec3(x, y, z))); }
    static Transform rotate_x(Float th) {
        Float c = lm_cos(th), s = lm_sin(th);
        return mat3(Mat3(Vec3(1, 0, 0), Vec3(0, c, -s), Vec3(0, s, c)));
    }
    static Transform rotate_y(Float th) {
        Float c = lm_cos(th), s = lm_sin(th);
        return mat3(Mat3(Vec3(c, 0, s), Vec3(0, 1, 0), Vec3(-s, 0, c)));
    }
    static Transform rotate_z(Float th) {
        Float c = lm_cos(th), s = lm_sin(th);
        return mat3(Mat3(Vec3(c, -s, 0), Vec3(s, c, 0), Vec3(0, 0, 1)));
    }
    static Transform perspective(Float near, Float far) {                             // transform.rs:113-131
        Float a = far / (far - near);
        Float b = -far * near / (far - near);
        return Transform(Mat4(V4X, V4Y, V4Z * a + V4W * b, V4Z),
                         Mat4(V4X, V4Y, V4W, V4Z * (1.0 / b) + V4W * (1.0 / near)));
    }
    Transform mul(const Transform& rhs) const { return Transform(m.mul(rhs.m), rhs.inv.mul(inv)); }  // :190-198
    Vec3 transform_pt(Vec3 p) const { return m.mul_vec4(extend(p, 1.0)).project(); }
    Vec3 transform_pt_inv(Vec3 p) const { return inv.mul_vec4(extend(p, 1.0)).project(); }
    Vec3 transform_dir(Vec3 d) const { return m.mul_vec4(extend(d, 0.0)).project(); }
    Vec3 transform_dir_inv(Vec3 d) const { return inv.mul_vec4(extend(d, 0.0)).project(); }
    Mat3 to_normal() const { return inv.to_mat3().transpose(); }
    Mat3 to_normal_inv() const { return m.to_mat3().transpose(); }
    Mat3 to_mat3() const { return m.to_mat3(); }
    Vec3 to_translation() const { return Vec3(m.y0.w, m.y1.w, m.y2.w); }
    Vec3 to_scale() const {
        Mat3 mt = m.to_mat3().transpose();
        return Vec3(mt.y0.length(), mt.y1.length(), mt.y2.length());
    }
    Transform abs() const { Transform t; t.m = m.abs(); Float n = std::nan(""); t.inv = Mat4(Vec4(n,n,n,n),Vec4(n,n,n,n),Vec4(n,n,n,n),Vec4(n,n,n,n)); return t; }
    const Vec4& row(int y) const { return y == 0 ? m.y0 : (y == 1 ? m.y1 : (y == 2 ? m.y2 : m.y3)); }
};

// src/efloat.rs
static inline Float gamma_(uint64_t n) {                                               // efloat.rs:5-8
    Float nf = (Float)n;
    const Float eps = std::numeric_limits<double>::epsilon();
    return (nf * eps) / (1.0 - nf * eps);
}
static inline uint64_t to_bits(Float v) { uint64_t b; std::memcpy(&b, &v, 8); return b; }
static inline Float from_bits(uint64_t b) { Float v; std::memcpy(&v, &b, 8); return v; }
static inline Float next_float(Float v) {                                              // efloat.rs:11-23
    if (std::isinf(v) && v > 0.0) return v;
    if (v == -0.0) v = 0.0;
    uint64_t bits = (v >= 0.0) ? to_bits(v) + 1 : to_bits(v) - 1;
    return from_bits(bits);
}
static inline Float previous_float(Float v) {                                          // efloat.rs:26-38
    if (std::isinf(v) && v < 0.0) return v;
    if (v == 0.0) v = -0.0;
    uint64_t bits = (v > 0.0) ? to_bits(v) - 1 : to_bits(v) + 1;
    return from_bits(bits);
}
struct EFloat {                                                                        // efloat.rs:41-204
    Float value, low, high;
    EFloat(Float v) : value(v), low(v), high(v) {}
    EFloat(Float v, Float l, Float h) : value(v), low(l), high(h) {}
    EFloat sqrt() const { return EFloat(std::sqrt(value), previous_float(std::sqrt(low)), next_float(std::sqrt(high))); }
};
static inline EFloat operator-(EFloat a) { return EFloat(-a.value, -a.low, -a.high); }
static inline EFloat operator+(EFloat a, EFloat b) {
    return EFloat(a.value + b.value, previous_float(a.low + b.low), next_float(a.high + b.high));
}
static inline EFloat operator-(EFloat a, EFloat b) {
    return EFloat(a.value - b.value, previous_float(a.low - b.high), next_float(a.high - b.low));
}
static inline EFloat operator*(EFloat a, EFloat b) {
    Float p[4] = {a.low * b.low, a.low * b.high, a.high * b.low, a.high * b.high};
    Float mn = fmin_(fmin_(fmin_(p[0], p[1]), p[2]), p[3]);
    Float mx = fmax_(fmax_(fmax_(p[0], p[1]), p[2]), p[3]);
    return EFloat(a.value * b.value, previous_float(mn), next_float(mx));
}
static inline EFloat operator/(EFloat a, EFloat b) {
    if (b.low < 0.0 && b.high > 0.0) return EFloat(a.value / b.value, -INF, INF);
    Float p[4] = {a.low / b.low, a.low / b.high, a.high / b.low, a.high / b.high};
    Float mn = fmin_(fmin_(fmin_(p[0], p[1]), p[2]), p[3]);
    Float mx = fmax_(fmax_(fmax_(p[0], p[1]), p[2]), p[3]);
    return EFloat(a.value / b.value, previous_float(mn), next_float(mx));
}
static inline bool efloat_quadratic(EFloat a, EFloat b, EFloat c, EFloat& t0, EFloat& t1) {  // efloat.rs:67-83
    Float disc = b.value * b.value - 4.0 * a.value * c.value;
    if (disc < 0.0) return false;
    EFloat disc_root = EFloat(disc).sqrt();
    t0 = (-b - disc_root) / (EFloat(2.0) * a);
    t1 = (-b + disc_root) / (EFloat(2.0) * a);
    if (t0.value > t1.value) std::swap(t0, t1);
    return true;
}

// src/math/complex.rs
struct Complex {
    Float Re, Im;
    Complex(Float r, Float i) : Re(r), Im(i) {}
    Float norm_sqr() const { return Re * Re + Im * Im; }
    Complex co() const { return Complex(Re, -Im); }
    Float norm() const { return std::sqrt(norm_sqr()); }
    Float arg() const { return lm_atan2(Im, Re); }   // libm::atan2 in the reference: ulp-level parity unpinned
    Complex sqrt() const {
        return Complex(std::sqrt(norm()) * lm_cos(arg() / 2.0), std::sqrt(norm()) * lm_sin(arg() / 2.0));
    }
};
static inline Complex operator+(Complex a, Complex b) { return Complex(a.Re + b.Re, a.Im + b.Im); }
static inline Complex operator+(Float a, Complex b) { return Complex(a + b.Re, b.Im); }
static inline Complex operator-(Complex a, Complex b) { return Complex(a.Re - b.Re, a.Im - b.Im); }
static inline Complex operator-(Float a, Complex b) { return Complex(a - b.Re, -b.Im); }
static inline Complex operator*(Complex a, Complex b) { return Complex(a.Re * b.Re - a.Im * b.Im, a.Re * b.Im + a.Im * b.Re); }
static inline Complex operator*(Complex a, Float b) { return Complex(a.Re * b, a.Im * b); }
static inline Complex operator*(Float a, Complex b) { return Complex(a * b.Re, a * b.Im); }
static inline Complex operator/(Complex a, Float b) {
    if (b == 0.0) return Complex(std::nan(""), std::nan(""));
    return Complex(a.Re / b, a.Im / b);
}
static inline Complex operator/(Complex a, Complex b) {
    if (b.Re == 0.0 && b.Im == 0.0) return Complex(std::nan(""), std::nan(""));
    return a * b.co() / b.norm_sqr();
}
static inline Complex operator/(Float a, Complex b) {
    if (b.Re == 0.0 && b.Im == 0.0) return Complex(std::nan(""), std::nan(""));
    return a * b.co() / b.norm_sqr();
}

// src/math/spherical_utils.rs
namespace sph {
static inline Float cos_theta(Vec3 w) { return w.z; }
static inline Float cos2_theta(Vec3 w) { return w.z * w.z; }
static inline Float sin2_theta(Vec3 w) { return fmax_(1.0 - cos2_theta(w), 0.0); }
static inline Float sin_theta(Vec3 w) { return std::sqrt(sin2_theta(w)); }
static inline Float tan2_theta(Vec3 w) { return sin2_theta(w) / cos2_theta(w); }
static inline Float cos_phi(Vec3 w) { Float s = sin_theta(w); return s == 0.0 ? 1.0 : clampf(w.x / s, -1.0, 1.0); }
static inline Float sin_phi(Vec3 w) { Float s = sin_theta(w); return s == 0.0 ? 0.0 : clampf(w.y / s, -1.0, 1.0); }
static inline bool same_hemisphere(Vec3 v, Vec3 u) { return cos_theta(v) * cos_theta(u) > 0.0; }
}

// src/tracer/onb.rs:19-62 (Duff et al. 2017)
struct Onb {
    Vec3 u, v, w;
    explicit Onb(Vec3 w_) : w(w_) {
        Float sgn = signum(w.z);
        Float a = -1.0 / (sgn + w.z);
        Float b = w.x * w.y * a;
        u = Vec3(1.0 + sgn * w.x * w.x * a, sgn * b, -sgn * w.x);
        v = Vec3(b, sgn + w.y * w.y * a, -w.y);
    }
    Vec3 to_world(Vec3 p) const { return p.x * u + p.y * v + p.z * w; }
    Vec3 to_local(Vec3 p) const { return Vec3(p.dot(u), p.dot(v), p.dot(w)); }
};

// src/rng/maps.rs
static inline Vec2 square_to_disk(Vec2 r) {
    Vec2 offset = 2.0 * r - Vec2(1, 1);
    if (offset.x == 0.0 && offset.y == 0.0) return Vec2(0, 0);
    Float rr, theta;
    if (std::fabs(offset.x) > std::fabs(offset.y)) { rr = offset.x; theta = PI * (offset.y / offset.x) / 4.0; }
    else { rr = offset.y; theta = PI * (0.5 - (offset.x / offset.y) / 4.0); }
    return rr * Vec2(lm_cos(theta), lm_sin(theta));
}
static inline Vec3 square_to_cos_hemisphere(Vec2 r) {
    Vec2 d = square_to_disk(r);
    Float z = std::sqrt(fmax_(1.0 - d.x * d.x - d.y * d.y, 0.0));
    return Vec3(d.x, d.y, z);
}
static inline Vec3 square_to_sphere(Vec2 r) {
    Float z = 1.0 - 2.0 * r.y;
    Float rr = std::sqrt(fmax_(1.0 - z * z, 0.0));
    Float phi = 2.0 * PI * r.x;
    return Vec3(rr * lm_cos(phi), rr * lm_sin(phi), z);
}

// ---- random streams -------------------------------------------------------------------------
// Rng::Xorshift restates src/rng.rs:39-116.  Rng::Philox is NOT in the reference: it is the
// counter-based per-(pixel,sample) stream the GPU path uses (BASELINE north_star "counter-based
// per-pixel streams"); the oracle carries it so GPU and CPU can be compared path-for-path.
struct Rng {
    // xorshift state
    uint64_t hi = 1, lo = 1;
    // philox state
    bool philox = false;
    uint32_t key0 = 0, key1 = 0;
    uint32_t c_pixel = 0, c_sample = 0, c_stream = 0;
    uint32_t draws = 0;
    uint32_t buf[4];

    static Rng xorshift(uint64_t seed) {                                      // rng.rs:40-48
        Rng r; r.lo = std::max<uint64_t>(seed, 1); r.hi = r.lo; r.step(); r.step(); r.step(); return r;
    }
    static Rng counter(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stream) {
        Rng r; r.philox = true; r.key0 = (uint32_t)seed; r.key1 = (uint32_t)(seed >> 32);
        r.c_pixel = pixel; r.c_sample = sample; r.c_stream = stream; r.draws = 0; return r;
    }
    uint64_t step() {                                                         // rng.rs:50-59
        uint64_t l = lo, h = hi;
        hi = l;
        h ^= h << 23; h ^= h >> 17; h ^= l;
        lo = h + l;
        return h;
    }
    static inline void philox_round(uint32_t* c, uint32_t k0, uint32_t k1) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    static inline void philox4x32_10(uint32_t* c, uint32_t k0, uint32_t k1) {
        for (int i = 0; i < 10; i++) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    }
    uint64_t gen_u64() {
        if (!philox) return step();
        // draw k uses block k/2, half k%2
        if ((draws & 1u) == 0) {
            buf[0] = c_pixel; buf[1] = c_sample; buf[2] = draws >> 1; buf[3] = c_stream;
            philox4x32_10(buf, key0, key1);
        }
        uint64_t v = (draws & 1u) == 0 ? (((uint64_t)buf[1] << 32) | buf[0]) : (((uint64_t)buf[3] << 32) | buf[2]);
        draws++;
        return v;
    }
    Float gen_float() {                                                       // rng.rs:71-75
        Float v = (Float)gen_u64();
        return fmin_(v * std::ldexp(1.0, -64), 1.0 - EPSILON);
    }
    Vec2 gen_vec2() { Float a = gen_float(); Float b = gen_float(); return Vec2(a, b); }
};

}  // namespace oracle
