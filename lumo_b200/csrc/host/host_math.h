// Host-side f64 math for the scene builder.  Operation order follows the reference's
// src/math/{vec3,mat3,mat4,transform}.rs so that bounding boxes, instance matrices and camera
// matrices come out bit-identical to lumo's (compile with -ffp-contract=off).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include "../common/lumo_math.h"

namespace lumo_host {

static const double kInf = std::numeric_limits<double>::infinity();
static const double kPi = 3.14159265358979323846264338327950288;

struct V3 { double x, y, z; };
static inline V3 v3(double x, double y, double z) { V3 r = {x, y, z}; return r; }
static inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 mul(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 divs(V3 a, double s) { return v3(a.x / s, a.y / s, a.z / s); }
static inline V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline double length(V3 a) { return std::sqrt(std::fmax(dot(a, a), 0.0)); }
static inline V3 normalize(V3 a) { return divs(a, length(a)); }
static inline V3 vmin(V3 a, V3 b) { return v3(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)); }
static inline V3 vmax(V3 a, V3 b) { return v3(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)); }
static inline double axis(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }
static inline double max_element(V3 a) { return std::fmax(a.x, std::fmax(a.y, a.z)); }

struct Box {
    V3 lo, hi;
    static Box empty() { Box b = {v3(kInf, kInf, kInf), v3(-kInf, -kInf, -kInf)}; return b; }
};
static inline Box merge(const Box& a, const Box& b) { Box r = {vmin(a.lo, b.lo), vmax(a.hi, b.hi)}; return r; }
static inline V3 center(const Box& b) { return add(b.lo, divs(sub(b.hi, b.lo), 2.0)); }       // aabb.rs:46-48
static inline double area(const Box& b) {                                                      // aabb.rs:56-60
    V3 d = sub(b.hi, b.lo);
    return 2.0 * (d.x * d.y + d.x * d.z + d.y * d.z);
}
static inline bool cuts(const Box& b, int ax, double p) { return axis(b.lo, ax) < p && p < axis(b.hi, ax); }
static inline void split_box(const Box& b, int ax, double v, Box& l, Box& r) {                  // aabb.rs:98-120
    l = b; r = b;
    if (ax == 0) { l.hi.x = v; r.lo.x = v; } else if (ax == 1) { l.hi.y = v; r.lo.y = v; } else { l.hi.z = v; r.lo.z = v; }
}

// row-major 4x4
struct M4 {
    double a[16];
    static M4 identity() { M4 m; std::memset(m.a, 0, sizeof m.a); m.a[0] = m.a[5] = m.a[10] = m.a[15] = 1.0; return m; }
};
static inline double dot4(const double* r, const double* c0, int stride) {   // mat4.rs:62-64 order
    return r[0] * c0[0] + r[1] * c0[stride] + r[2] * c0[2 * stride] + r[3] * c0[3 * stride];
}
static inline M4 matmul(const M4& l, const M4& r) {                                             // mat4.rs:182-213
    M4 o;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) o.a[4 * i + j] = dot4(&l.a[4 * i], &r.a[j], 4);
    return o;
}
struct M3 { double a[9]; };
static inline double det3(const M3& m) {                                                        // mat3.rs:41-52
    const double* a = m.a;
    double pos = a[0] * a[4] * a[8] + a[1] * a[5] * a[6] + a[2] * a[3] * a[7];
    double ng = a[2] * a[4] * a[6] + a[1] * a[3] * a[8] + a[0] * a[5] * a[7];
    return pos - ng;
}
static inline M3 transpose3(const M3& m) { M3 t = {{m.a[0], m.a[3], m.a[6], m.a[1], m.a[4], m.a[7], m.a[2], m.a[5], m.a[8]}}; return t; }
static inline M3 inv3(const M3& m) {                                                            // mat3.rs:64-72
    double id = 1.0 / det3(m);
    V3 y0 = v3(m.a[0], m.a[1], m.a[2]), y1 = v3(m.a[3], m.a[4], m.a[5]), y2 = v3(m.a[6], m.a[7], m.a[8]);
    V3 r0 = mul(cross(y1, y2), id), r1 = mul(cross(y2, y0), id), r2 = mul(cross(y0, y1), id);
    M3 rows = {{r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z}};
    return transpose3(rows);
}
static inline V3 mulv3(const M3& m, V3 v) { return v3(dot(v3(m.a[0], m.a[1], m.a[2]), v), dot(v3(m.a[3], m.a[4], m.a[5]), v), dot(v3(m.a[6], m.a[7], m.a[8]), v)); }
static inline M3 matmul3(const M3& l, const M3& r) {                                            // mat3.rs:83-102
    M3 o;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        o.a[3 * i + j] = l.a[3 * i] * r.a[j] + l.a[3 * i + 1] * r.a[3 + j] + l.a[3 * i + 2] * r.a[6 + j];
    return o;
}
static inline M4 from3(const M3& m) { M4 o = M4::identity(); for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o.a[4 * i + j] = m.a[3 * i + j]; return o; }
static inline M3 to3(const M4& m) { M3 o; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o.a[3 * i + j] = m.a[4 * i + j]; return o; }

// src/math/transform.rs
struct Xform {
    M4 m, inv;
    static Xform identity() { Xform t = {M4::identity(), M4::identity()}; return t; }
    static Xform from_m3(const M3& m3) { Xform t = {from3(m3), from3(inv3(m3))}; return t; }
    static Xform translation(double x, double y, double z) {                                    // :133-149
        Xform t = identity();
        // X + W*x etc: 1 + 0*x on the diagonal stays exactly 1, off-diagonals 0 + 0*x = 0
        t.m.a[3] = 0.0 + 1.0 * x; t.m.a[7] = 0.0 + 1.0 * y; t.m.a[11] = 0.0 + 1.0 * z;
        t.inv.a[3] = 0.0 + 1.0 * (-x); t.inv.a[7] = 0.0 + 1.0 * (-y); t.inv.a[11] = 0.0 + 1.0 * (-z);
        return t;
    }
    static Xform scale(double x, double y, double z) { M3 d = {{x, 0, 0, 0, y, 0, 0, 0, z}}; return from_m3(d); }
    static Xform rot_x(double th) { double c = lm_cos(th), s = lm_sin(th); M3 r = {{1, 0, 0, 0, c, -s, 0, s, c}}; return from_m3(r); }
    static Xform rot_y(double th) { double c = lm_cos(th), s = lm_sin(th); M3 r = {{c, 0, s, 0, 1, 0, -s, 0, c}}; return from_m3(r); }
    static Xform rot_z(double th) { double c = lm_cos(th), s = lm_sin(th); M3 r = {{c, -s, 0, s, c, 0, 0, 0, 1}}; return from_m3(r); }
    static Xform perspective(double near, double far) {                                         // :113-131
        double a = far / (far - near), b = -far * near / (far - near);
        Xform t = identity();
        // m: rows X, Y, Z*a + W*b, Z
        t.m.a[8] = 0.0 * a + 0.0 * b; t.m.a[9] = 0.0 * a + 0.0 * b; t.m.a[10] = 1.0 * a + 0.0 * b; t.m.a[11] = 0.0 * a + 1.0 * b;
        t.m.a[12] = 0; t.m.a[13] = 0; t.m.a[14] = 1; t.m.a[15] = 0;
        // inv: rows X, Y, W, Z*(1/b) + W*(1/near)
        t.inv.a[8] = 0; t.inv.a[9] = 0; t.inv.a[10] = 0; t.inv.a[11] = 1;
        t.inv.a[12] = 0.0 * (1.0 / b) + 0.0 * (1.0 / near); t.inv.a[13] = t.inv.a[12];
        t.inv.a[14] = 1.0 * (1.0 / b) + 0.0 * (1.0 / near); t.inv.a[15] = 0.0 * (1.0 / b) + 1.0 * (1.0 / near);
        return t;
    }
};
static inline Xform compose(const Xform& l, const Xform& r) { Xform t = {matmul(l.m, r.m), matmul(r.inv, l.inv)}; return t; }  // :190-198
static inline V3 apply(const M4& m, V3 p, double w) {                                           // mat4.rs:172-179 + project :46-52
    double v[4] = {p.x, p.y, p.z, w};
    double o[4];
    for (int i = 0; i < 4; i++) o[i] = m.a[4 * i] * v[0] + m.a[4 * i + 1] * v[1] + m.a[4 * i + 2] * v[2] + m.a[4 * i + 3] * v[3];
    if (o[3] == 0.0) return v3(o[0], o[1], o[2]);
    return v3(o[0] / o[3], o[1] / o[3], o[2] / o[3]);
}

static inline double gamma_n(uint64_t n) { double nf = (double)n; const double e = std::numeric_limits<double>::epsilon(); return (nf * e) / (1.0 - nf * e); }
static inline uint64_t sat_u64(double x) { if (!(x > 0.0)) return 0; if (x >= 18446744073709551616.0) return UINT64_MAX; return (uint64_t)x; }

}  // namespace lumo_host
