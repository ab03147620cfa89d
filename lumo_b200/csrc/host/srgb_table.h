// The RGB -> spectrum coefficient table of the reference (`srgb.coeff`, read by src/tracer/color/spectrum/tables.rs:6-84):
// "SPEC", u32 resolution = 64, 64 f32 scale knots, then 3 x 64^3 x 3 f32 sigmoid-polynomial coefficients.  The 9.4 MB file
// is not part of the reference mount, so it is regenerated here with the published procedure it was made by
// (W. Jakob and J. Hanika, "A Low-Dimensional Function Space for Efficient Spectral Upsampling", 2019, and the optimiser
// published with it): for every grid column (largest channel l, the two other channels relative to it), a Gauss-Newton fit
// of the three coefficients in CIELAB with central-difference Jacobians, warm-started along the brightness axis from knot
// res/5 upwards and then downwards; scale knots = smoothstep(smoothstep(k / (res - 1))); CIE 1931 observer and D65 on the
// 95-sample, 5 nm grid of color/samples.rs refined three times, Simpson 3/8 weights.
// tests/test_spectrum.py checks lookups in the result against the reference's 33 known-answer vectors (spectrum_tests.rs:37-111)
// at the reference's own tolerance.
#pragma once
#include "spectra_data.h"
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace lumo_host {
namespace srgb_table {

static const int RES = 64, CIE_N = 95, FINE = (CIE_N - 1) * 3 + 1;
static const double LMIN = 360.0, LMAX = 830.0, GN_EPS = 1e-4;
static const double XYZ_TO_SRGB[3][3] = {{3.240479, -1.537150, -0.498535}, {-0.969256, 1.875991, 0.041556}, {0.055648, -0.204043, 1.057311}};
static const double SRGB_TO_XYZ[3][3] = {{0.412453, 0.357580, 0.180423}, {0.212671, 0.715160, 0.072169}, {0.019334, 0.119193, 0.950227}};

struct Tables { double lambda[FINE], rgb[3][FINE], white[3]; };

static inline double cie_interp(const double* data, double x) {
    x -= LMIN; x *= (CIE_N - 1) / (LMAX - LMIN);
    int offset = (int)x; if (offset < 0) offset = 0; if (offset > CIE_N - 2) offset = CIE_N - 2;
    const double w = x - offset;
    return (1.0 - w) * data[offset] + w * data[offset + 1];
}
static inline void init_tables(Tables& T, double illum_div) {
    std::memset(&T, 0, sizeof T);
    const double h = (LMAX - LMIN) / (FINE - 1);
    // the illuminant is normalised so that the white point has Y = 1 (the optimiser's tables carry that normalisation in the data)
    double ynorm = 0.0;
    for (int i = 0; i < FINE; i++) {
        const double lambda = LMIN + i * h;
        double weight = 3.0 / 8.0 * h;
        if (i == 0 || i == FINE - 1) {} else if ((i - 1) % 3 == 2) weight *= 2.0; else weight *= 3.0;
        ynorm += cie_interp(spectra::Y, lambda) * cie_interp(spectra::D65, lambda) * weight;
    }
    for (int i = 0; i < FINE; i++) {
        const double lambda = LMIN + i * h;
        const double xyz[3] = {cie_interp(spectra::X, lambda), cie_interp(spectra::Y, lambda), cie_interp(spectra::Z, lambda)};
        const double I = cie_interp(spectra::D65, lambda) / (illum_div > 0.0 ? illum_div : ynorm);
        double weight = 3.0 / 8.0 * h;
        if (i == 0 || i == FINE - 1) {} else if ((i - 1) % 3 == 2) weight *= 2.0; else weight *= 3.0;
        T.lambda[i] = lambda;
        for (int k = 0; k < 3; k++) for (int j = 0; j < 3; j++) T.rgb[k][i] += XYZ_TO_SRGB[k][j] * xyz[j] * I * weight;
        for (int k = 0; k < 3; k++) T.white[k] += xyz[k] * I * weight;
    }
}
static inline double sigmoid(double x) { return 0.5 * x / std::sqrt(1.0 + x * x) + 0.5; }
static inline double smoothstep(double x) { return x * x * (3.0 - 2.0 * x); }
static inline void cie_lab(const Tables& T, double* p) {
    double X = 0, Y = 0, Z = 0;
    for (int j = 0; j < 3; j++) { X += p[j] * SRGB_TO_XYZ[0][j]; Y += p[j] * SRGB_TO_XYZ[1][j]; Z += p[j] * SRGB_TO_XYZ[2][j]; }
    auto f = [](double t) { const double d = 6.0 / 29.0; return t > d * d * d ? std::cbrt(t) : t / (d * d * 3.0) + 4.0 / 29.0; };
    const double fx = f(X / T.white[0]), fy = f(Y / T.white[1]), fz = f(Z / T.white[2]);
    p[0] = 116.0 * fy - 16.0; p[1] = 500.0 * (fx - fy); p[2] = 200.0 * (fy - fz);
}
static inline void eval_residual(const Tables& T, const double* coeffs, const double* rgb, double* residual) {
    double out[3] = {0, 0, 0};
    for (int i = 0; i < FINE; i++) {
        const double lambda = (T.lambda[i] - LMIN) / (LMAX - LMIN);
        double x = 0.0;
        for (int k = 0; k < 3; k++) x = x * lambda + coeffs[k];
        const double s = sigmoid(x);
        for (int j = 0; j < 3; j++) out[j] += T.rgb[j][i] * s;
    }
    cie_lab(T, out);
    std::memcpy(residual, rgb, 3 * sizeof(double));
    cie_lab(T, residual);
    for (int j = 0; j < 3; j++) residual[j] -= out[j];
}
static inline void eval_jacobian(const Tables& T, const double* coeffs, const double* rgb, double jac[3][3]) {
    double r0[3], r1[3], tmp[3];
    for (int i = 0; i < 3; i++) {
        std::memcpy(tmp, coeffs, sizeof tmp); tmp[i] -= GN_EPS; eval_residual(T, tmp, rgb, r0);
        std::memcpy(tmp, coeffs, sizeof tmp); tmp[i] += GN_EPS; eval_residual(T, tmp, rgb, r1);
        for (int j = 0; j < 3; j++) jac[j][i] = (r1[j] - r0[j]) * 1.0 / (2.0 * GN_EPS);
    }
}
// LU decomposition with partial pivoting, as the published optimiser solves its 3x3 systems
static inline bool lup_decompose(double A[3][3], int* P, double tol) {
    for (int i = 0; i <= 3; i++) P[i] = i;
    for (int i = 0; i < 3; i++) {
        double maxA = 0.0; int imax = i;
        for (int k = i; k < 3; k++) { const double a = std::fabs(A[k][i]); if (a > maxA) { maxA = a; imax = k; } }
        if (maxA < tol) return false;
        if (imax != i) {
            const int j = P[i]; P[i] = P[imax]; P[imax] = j;
            for (int c = 0; c < 3; c++) { const double t = A[i][c]; A[i][c] = A[imax][c]; A[imax][c] = t; }
            P[3]++;
        }
        for (int j = i + 1; j < 3; j++) { A[j][i] /= A[i][i]; for (int k = i + 1; k < 3; k++) A[j][k] -= A[j][i] * A[i][k]; }
    }
    return true;
}
static inline void lup_solve(double A[3][3], const int* P, const double* b, double* x) {
    for (int i = 0; i < 3; i++) { x[i] = b[P[i]]; for (int k = 0; k < i; k++) x[i] -= A[i][k] * x[k]; }
    for (int i = 2; i >= 0; i--) { for (int k = i + 1; k < 3; k++) x[i] -= A[i][k] * x[k]; x[i] = x[i] / A[i][i]; }
}
static inline bool gauss_newton(const Tables& T, const double rgb[3], double coeffs[3], double clamp_max, bool clamp_abs, int it = 15) {
    for (int i = 0; i < it; i++) {
        double J[3][3], residual[3], x[3]; int P[4];
        eval_residual(T, coeffs, rgb, residual);
        eval_jacobian(T, coeffs, rgb, J);
        if (!lup_decompose(J, P, 1e-15)) return false;
        lup_solve(J, P, residual, x);
        double r = 0.0;
        for (int j = 0; j < 3; j++) { coeffs[j] -= x[j]; r += residual[j] * residual[j]; }
        const double mx = clamp_abs ? std::fmax(std::fmax(std::fabs(coeffs[0]), std::fabs(coeffs[1])), std::fabs(coeffs[2])) : std::fmax(std::fmax(coeffs[0], coeffs[1]), coeffs[2]);
        if (clamp_max > 0.0 && mx > clamp_max) for (int j = 0; j < 3; j++) coeffs[j] *= clamp_max / mx;
        if (r < 1e-6) break;
    }
    return true;
}
// out: RES scale knots, then 3 * RES^3 * 3 coefficients (the file's payload after "SPEC" + resolution)
// illum_div: what the D65 table is divided by; the published optimiser's data carries 1 / 10566.864005283874576 (white Y = 1 on the 5 nm grid)
static inline void generate(float* scale_out, float* data_out, int threads, double illum_div = 10566.864005283874576, double clamp_max = 200.0) {
    Tables T; init_tables(T, illum_div);
    double scale[RES];
    for (int k = 0; k < RES; k++) { scale[k] = smoothstep(smoothstep((double)k / (double)(RES - 1))); scale_out[k] = (float)scale[k]; }
    auto column = [&](int l, int j, int i) {
        const double x = (double)i / (double)(RES - 1), y = (double)j / (double)(RES - 1);
        const int start = RES / 5;
        auto store = [&](int k, const double* coeffs) {
            const double c0 = LMIN, c1 = 1.0 / (LMAX - LMIN);
            const double A = coeffs[0], B = coeffs[1], C = coeffs[2];
            const size_t idx = (((size_t)l * RES + k) * RES + j) * RES + i;
            data_out[3 * idx + 0] = (float)(A * (c1 * c1));
            data_out[3 * idx + 1] = (float)(B * c1 - 2.0 * A * c0 * (c1 * c1));
            data_out[3 * idx + 2] = (float)(C - B * c0 * c1 + A * ((c0 * c1) * (c0 * c1)));
        };
        double coeffs[3] = {0, 0, 0}, rgb[3];
        for (int k = start; k < RES; k++) {
            const double b = scale[k];
            rgb[l] = b; rgb[(l + 1) % 3] = x * b; rgb[(l + 2) % 3] = y * b;
            gauss_newton(T, rgb, coeffs, std::fabs(clamp_max), clamp_max < 0.0);
            store(k, coeffs);
        }
        coeffs[0] = coeffs[1] = coeffs[2] = 0.0;
        for (int k = start; k >= 0; k--) {
            const double b = scale[k];
            rgb[l] = b; rgb[(l + 1) % 3] = x * b; rgb[(l + 2) % 3] = y * b;
            gauss_newton(T, rgb, coeffs, std::fabs(clamp_max), clamp_max < 0.0);
            store(k, coeffs);
        }
    };
    const int n_cols = 3 * RES * RES;
    if (threads < 1) threads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back([&, t]() { for (int c = t; c < n_cols; c += threads) column(c / (RES * RES), (c / RES) % RES, c % RES); });
    for (auto& t : th) t.join();
}

}  // namespace srgb_table
}  // namespace lumo_host
