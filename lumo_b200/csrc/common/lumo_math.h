// lumo_math.h — the transcendental functions of the hot path as ONE source for every side.
//
// Rust's f64::sin / cos / atanh / cosh / exp / acos and libm::atan2 (src/math/complex.rs:52) are
// not pinned to the bit by anything the reference ships ("parity unpinned at ulp
// level"), and glibc's libm and CUDA's device libm differ in the last bit now and then.  A last-bit
// difference in a sampled direction can flip a geometric decision further down the path, so films
// rendered from identical Philox streams on the CPU and on the GPU were only equal on ~94-99 % of
// the pixels.  This header removes that: every function below is written with IEEE-754 +, -, *, /,
// sqrt and integer bit operations only — no fused multiply-add, no library call — so the same
// source compiled by g++ (-ffp-contract=off) and by nvcc (-fmad=false) returns the same bits.
// The device kernels (shade.cuh), the host builder and the oracle all call these.
//
// The algorithms are the classical argument-reduction + minimax-polynomial ones published with
// FreeBSD msun / fdlibm (Sun Microsystems, 1993; the polynomial coefficients are the published
// ones); accuracy is ~1 ulp on the ranges the renderer uses (tests/test_lumo_math.py measures it
// against the host libm).
#pragma once
#include <stdint.h>
#include <string.h>

// LM_FN: small helpers, always inlined.  LM_API: the entry points; out of line on the device so that the shade kernels
// (instruction-cache bound, profiles/README.md) hold one copy of each instead of one per call site.
#if defined(__CUDACC__)
#define LM_FN __host__ __device__ __forceinline__
#define LM_API __host__ __device__ __noinline__
#define LM_SQRT(x) sqrt(x)
#else
#include <math.h>
#define LM_FN static inline
#define LM_API static inline
#define LM_SQRT(x) sqrt(x)
#endif

LM_FN uint64_t lm_bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
LM_FN double lm_from_bits(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
LM_FN double lm_abs(double x) { return lm_from_bits(lm_bits(x) & 0x7FFFFFFFFFFFFFFFull); }
LM_FN int lm_isnan(double x) { return (lm_bits(x) & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull; }
LM_FN int lm_isinf(double x) { return (lm_bits(x) & 0x7FFFFFFFFFFFFFFFull) == 0x7FF0000000000000ull; }
LM_FN double lm_nan(void) { return lm_from_bits(0x7FF8000000000000ull); }
LM_FN double lm_inf(void) { return lm_from_bits(0x7FF0000000000000ull); }
LM_FN double lm_copysign(double m, double s) { return lm_from_bits((lm_bits(m) & 0x7FFFFFFFFFFFFFFFull) | (lm_bits(s) & 0x8000000000000000ull)); }
// 2^k for k in [-1022, 1023]
LM_FN double lm_pow2(int k) { return lm_from_bits((uint64_t)(k + 1023) << 52); }
// round to nearest integer (ties to even) for |x| < 2^51, without a library call
LM_FN double lm_rint(double x) {
    const double big = 6755399441055744.0;   // 1.5 * 2^52
    if (lm_abs(x) >= 4503599627370496.0) return x;
    const double t = x + big;                // not folded: neither compiler reassociates without fast-math
    return t - big;
}

// ---- exp -------------------------------------------------------------------------------------------
LM_API double lm_exp(double x) {
    const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    if (lm_isnan(x)) return x;
    if (x > 709.782712893383973096) return lm_inf();
    if (x < -745.13321910194110842) return 0.0;
    double hi = x, lo = 0.0; int k = 0;
    if (lm_abs(x) > 0.34657359027997264 /* 0.5 ln2 */) {
        k = (int)(invln2 * x + (x < 0.0 ? -0.5 : 0.5));
        hi = x - (double)k * ln2hi; lo = (double)k * ln2lo;
    } else if (lm_abs(x) < 3.725290298461914e-09 /* 2^-28 */) return 1.0 + x;
    const double r = hi - lo;
    const double t = r * r;
    const double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    double y;
    if (k == 0) return 1.0 - ((r * c) / (c - 2.0) - r);
    y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    if (k >= -1021 && k <= 1023) return y * lm_pow2(k);
    if (k > 1023) return y * lm_pow2(1023) * lm_pow2(k - 1023);
    return y * lm_pow2(k + 1000) * lm_pow2(-1000);
}

// ---- log -------------------------------------------------------------------------------------------
LM_API double lm_log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
    if (lm_isnan(x)) return x;
    if (x < 0.0) return lm_nan();
    if (x == 0.0) return -lm_inf();
    if (lm_isinf(x)) return x;
    int k = 0;
    uint64_t u = lm_bits(x);
    if ((u >> 52) == 0) { x = x * 18014398509481984.0 /* 2^54 */; u = lm_bits(x); k = -54; }
    // x = 2^k * m with sqrt(2)/2 <= m < sqrt(2)
    uint32_t hx = (uint32_t)(u >> 32);
    hx += 0x3ff00000u - 0x3fe6a09eu;
    k += (int)(hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
    const double m = lm_from_bits(((uint64_t)hx << 32) | (u & 0xffffffffull));
    const double f = m - 1.0;
    const double hfsq = 0.5 * f * f;
    const double s = f / (2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    const double R = t2 + t1;
    const double dk = (double)k;
    return s * (hfsq + R) + dk * ln2_lo - hfsq + f + dk * ln2_hi;
}
// log(1 + t), accurate for small t (Kahan's correction of log(u) by t / (u - 1))
LM_FN double lm_log1p(double t) {
    const double u = 1.0 + t;
    if (u == 1.0) return t;
    if (lm_isinf(u)) return u;
    return lm_log(u) * (t / (u - 1.0));
}
LM_API double lm_atanh(double x) {
    const double a = lm_abs(x);
    if (lm_isnan(x) || a > 1.0) return lm_nan();
    if (a == 1.0) return lm_copysign(lm_inf(), x);
    const double r = 0.5 * lm_log1p((a + a) / (1.0 - a));
    return lm_copysign(r, x);
}
LM_API double lm_cosh(double x) {
    const double a = lm_abs(x);
    if (a > 709.0) return lm_isnan(x) ? x : lm_inf();
    const double t = lm_exp(a);
    return 0.5 * t + 0.5 / t;
}
// x^y for x >= 0 (the transfer curves of the film): exp(y log x)
LM_API double lm_pow(double x, double y) {
    if (lm_isnan(x) || lm_isnan(y)) return lm_nan();
    if (y == 0.0) return 1.0;
    if (x < 0.0) return lm_nan();
    if (x == 0.0) return y > 0.0 ? 0.0 : lm_inf();
    return lm_exp(y * lm_log(x));
}

// ---- sin / cos -------------------------------------------------------------------------------------
// x = n * pi/2 + (y0 + y1), |y0 + y1| <= pi/4 (two-constant Cody-Waite; |x| up to ~1e6 keeps n * pio2_1 exact)
LM_FN int lm_rem_pio2(double x, double* y0, double* y1) {
    const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11,
                 pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21;
    const double fn = lm_rint(x * invpio2);
    double r = x - fn * pio2_1;
    double w = fn * pio2_1t;
    double y = r - w;
    // second step when the first one cancelled more than ~16 bits
    const int ex = (int)((lm_bits(x) >> 52) & 0x7ff), ey = (int)((lm_bits(y) >> 52) & 0x7ff);
    if (ex - ey > 16) {
        const double t = r;
        w = fn * pio2_2; r = t - w;
        w = fn * pio2_2t - ((t - r) - w);
        y = r - w;
    }
    *y0 = y; *y1 = (r - y) - w;
    return (int)((long long)fn & 3ll);
}
LM_FN double lm_ksin(double x, double y, int iy) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = x * x, w = z * z;
    const double r = S2 + z * (S3 + z * S4) + z * w * (S5 + z * S6);
    const double v = z * x;
    if (iy == 0) return x + v * (S1 + z * r);
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
LM_FN double lm_kcos(double x, double y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double z = x * x, w = z * z;
    const double r = z * (C1 + z * (C2 + z * C3)) + w * w * (C4 + z * (C5 + z * C6));
    const double hz = 0.5 * z;
    const double ww = 1.0 - hz;
    return ww + (((1.0 - ww) - hz) + (z * r - x * y));
}
LM_API double lm_sin(double x) {
    if (lm_isnan(x) || lm_isinf(x)) return lm_nan();
    if (lm_abs(x) <= 0.78539816339744827900) { if (lm_abs(x) < 7.450580596923828e-09) return x; return lm_ksin(x, 0.0, 0); }
    double y0, y1;
    const int n = lm_rem_pio2(x, &y0, &y1);
    switch (n) {
    case 0: return lm_ksin(y0, y1, 1);
    case 1: return lm_kcos(y0, y1);
    case 2: return -lm_ksin(y0, y1, 1);
    default: return -lm_kcos(y0, y1);
    }
}
LM_API double lm_cos(double x) {
    if (lm_isnan(x) || lm_isinf(x)) return lm_nan();
    if (lm_abs(x) <= 0.78539816339744827900) { if (lm_abs(x) < 7.450580596923828e-09) return 1.0; return lm_kcos(x, 0.0); }
    double y0, y1;
    const int n = lm_rem_pio2(x, &y0, &y1);
    switch (n) {
    case 0: return lm_kcos(y0, y1);
    case 1: return -lm_ksin(y0, y1, 1);
    case 2: return -lm_kcos(y0, y1);
    default: return lm_ksin(y0, y1, 1);
    }
}

// sin and cos of the same argument with one reduction; the same bits as lm_sin / lm_cos
LM_API void lm_sincos(double x, double* s, double* c) {
    if (lm_isnan(x) || lm_isinf(x)) { *s = lm_nan(); *c = lm_nan(); return; }
    if (lm_abs(x) <= 0.78539816339744827900) {
        const int tiny = lm_abs(x) < 7.450580596923828e-09;
        *s = tiny ? x : lm_ksin(x, 0.0, 0); *c = tiny ? 1.0 : lm_kcos(x, 0.0);
        return;
    }
    double y0, y1;
    const int n = lm_rem_pio2(x, &y0, &y1);
    const double ks = lm_ksin(y0, y1, 1), kc = lm_kcos(y0, y1);
    switch (n) {
    case 0: *s = ks; *c = kc; break;
    case 1: *s = kc; *c = -ks; break;
    case 2: *s = -ks; *c = -kc; break;
    default: *s = -kc; *c = ks; break;
    }
}

// ---- atan / atan2 / acos ---------------------------------------------------------------------------
LM_API double lm_atan(double x) {
    const double hi0 = 4.63647609000806093515e-01, hi1 = 7.85398163397448278999e-01, hi2 = 9.82793723247329054082e-01, hi3 = 1.57079632679489655800e+00;
    const double lo0 = 2.26987774529616870924e-17, lo1 = 3.06161699786838301793e-17, lo2 = 1.39033110312309984516e-17, lo3 = 6.12323399573676603587e-17;
    const double a0 = 3.33333333333329318027e-01, a1 = -1.99999999998764832476e-01, a2 = 1.42857142725034663711e-01, a3 = -1.11111104054623557880e-01,
                 a4 = 9.09088713343650656196e-02, a5 = -7.69187620504482999495e-02, a6 = 6.66107313738753120669e-02, a7 = -5.83357013379057348645e-02,
                 a8 = 4.97687799461593236017e-02, a9 = -3.65315727442169155270e-02, a10 = 1.62858201153657823623e-02;
    if (lm_isnan(x)) return x;
    const double ax = lm_abs(x);
    if (ax >= 7.378697629483821e19 /* 2^66 */) return lm_copysign(hi3 + lo3, x);
    int id; double t;
    if (ax < 0.4375) { if (ax < 3.725290298461914e-09) return x; id = -1; t = ax; }
    else if (ax < 0.6875) { id = 0; t = (2.0 * ax - 1.0) / (2.0 + ax); }
    else if (ax < 1.1875) { id = 1; t = (ax - 1.0) / (ax + 1.0); }
    else if (ax < 2.4375) { id = 2; t = (ax - 1.5) / (1.0 + 1.5 * ax); }
    else { id = 3; t = -1.0 / ax; }
    const double z = t * t, w = z * z;
    const double s1 = z * (a0 + w * (a2 + w * (a4 + w * (a6 + w * (a8 + w * a10)))));
    const double s2 = w * (a1 + w * (a3 + w * (a5 + w * (a7 + w * a9))));
    double r;
    if (id < 0) r = t - t * (s1 + s2);
    else {
        const double hi = id == 0 ? hi0 : (id == 1 ? hi1 : (id == 2 ? hi2 : hi3));
        const double lo = id == 0 ? lo0 : (id == 1 ? lo1 : (id == 2 ? lo2 : lo3));
        r = hi - ((t * (s1 + s2) - lo) - t);
    }
    return lm_copysign(r, x);
}
LM_API double lm_atan2(double y, double x) {
    const double pi = 3.1415926535897931160E+00, pi_lo = 1.2246467991473531772E-16, pi_o_2 = 1.5707963267948965580E+00, pi_o_4 = 7.8539816339744827900E-01;
    if (lm_isnan(x) || lm_isnan(y)) return lm_nan();
    const int sy = (int)(lm_bits(y) >> 63), sx = (int)(lm_bits(x) >> 63);
    if (y == 0.0) return sx ? (sy ? -pi : pi) : y;                         // atan2(+-0, +x) = +-0, atan2(+-0, -x) = +-pi
    if (x == 0.0) return sy ? -pi_o_2 : pi_o_2;
    if (lm_isinf(x)) {
        if (lm_isinf(y)) return sx ? (sy ? -3.0 * pi_o_4 : 3.0 * pi_o_4) : (sy ? -pi_o_4 : pi_o_4);
        return sx ? (sy ? -pi : pi) : (sy ? -0.0 : 0.0);
    }
    if (lm_isinf(y)) return sy ? -pi_o_2 : pi_o_2;
    // exponent difference guards (|y/x| huge or tiny)
    const int ey = (int)((lm_bits(y) >> 52) & 0x7ff), ex = (int)((lm_bits(x) >> 52) & 0x7ff);
    double z;
    if (ey - ex > 64) z = pi_o_2 + 0.5 * pi_lo;
    else if (sx && ey - ex < -64) z = 0.0;
    else z = lm_atan(lm_abs(y / x));
    if (!sx) return sy ? -z : z;
    return sy ? (z - pi_lo) - pi : pi - (z - pi_lo);
}
LM_API double lm_acos(double x) {
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17, pi = 3.14159265358979311600e+00;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01, pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05, qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    if (lm_isnan(x)) return x;
    const double ax = lm_abs(x);
    if (ax > 1.0) return lm_nan();
    if (ax == 1.0) return x > 0.0 ? 0.0 : pi + 2.0 * pio2_lo;
    if (ax < 0.5) {
        if (ax < 1.3877787807814457e-17 /* 2^-56 */) return pio2_hi + pio2_lo;
        const double z = x * x;
        const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        return pio2_hi - (x - (pio2_lo - x * (p / q)));
    }
    if (x < 0.0) {
        const double z = (1.0 + x) * 0.5;
        const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const double s = LM_SQRT(z);
        const double w = (p / q) * s - pio2_lo;
        return pi - 2.0 * (s + w);
    }
    const double z = (1.0 - x) * 0.5;
    const double s = LM_SQRT(z);
    const double df = lm_from_bits(lm_bits(s) & 0xFFFFFFFF00000000ull);
    const double c = (z - df * df) / (s + df);
    const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const double w = (p / q) * s + c;
    return 2.0 * (df + w);
}
