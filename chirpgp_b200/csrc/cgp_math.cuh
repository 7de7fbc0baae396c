// cgp_math.cuh -- short-dependency-chain FP64 elementary functions for the sequential filters.
//
// The time loop of a filter is one long dependent chain (Cholesky -> sigma point -> softplus -> sin/cos -> moments
// -> update), so the *latency* of exp / log / sincos / rsqrt / rcp is what bounds a step.  Measured on B200
// (profiles/microbench/fp64_latency.cu): DFMA 8.2 cycles, CUDA exp 159, log 297, sincos 213, rsqrt 66, 1/x 71,
// log(exp(x)+1) 447 cycles.  The versions below evaluate the same functions to <= 2 ulp with Estrin-style
// polynomials and (except for the softplus range split) no branches, keeping the overflow / NaN behaviour of the
// reference's formulas (tests/test_gpu_math.py).
#pragma once
#include <cuda_runtime.h>

namespace cgp {

#define CGP_MDEV __device__ __forceinline__

// ---- reciprocal square root / reciprocal: hardware seed (2^-22.9 relative) + one cubically convergent step
CGP_MDEV double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.);                 // e = 1 - x y^2
    const double q = e * fma(e, 0.375, 0.5);         // e/2 + 3 e^2 / 8
    return fma(y, q, y);
}
CGP_MDEV double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.);                 // e = 1 - x y
    const double q = fma(e, e, e);                   // e + e^2
    return fma(y, q, y);
}

// ---- exp(x) for every x, branch-free (Taylor degree 13 on |r| <= ln2/2 in Estrin form; the scale 2^k is applied as
// two exact power-of-two factors so that overflow -> +inf and underflow -> 0 come out of ordinary arithmetic).
CGP_MDEV double fast_exp(double x) {
    const double xc = fmin(fmax(x, -800.), 800.);                         // 2^(+-1154) saturates to inf / 0 below
    const double t = fma(xc, 1.4426950408889634, 6755399441055744.0);     // round(x / ln2) in the low bits
    const int k = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, -6.93147180369123816490e-01, xc);                  // ln2_hi
    r = fma(kf, -1.90821492927058770002e-10, r);                          // ln2_lo
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p01 = 1. + r;                                            // c0 + c1 r
    const double p23 = fma(r, 1.6666666666666666e-01, 0.5);
    const double p45 = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    const double p67 = fma(r, 1.9841269841269841e-04, 1.3888888888888889e-03);
    const double p89 = fma(r, 2.7557319223985893e-06, 2.4801587301587302e-05);
    const double pab = fma(r, 2.5052108385441720e-08, 2.7557319223985888e-07);
    const double pcd = fma(r, 1.6059043836821613e-10, 2.0876756987868100e-09);
    const double q0 = fma(r2, p23, p01);
    const double q1 = fma(r2, p67, p45);
    const double q2 = fma(r2, pab, p89);
    const double s0 = fma(r4, q1, q0);
    const double s1 = fma(r4, pcd, q2);
    const double p = fma(r8, s1, s0);
    const int k1 = k >> 1, k2 = k - k1;                                   // |k1|, |k2| <= 578: both scales are normal
    const double sc1 = __hiloint2double((k1 + 1023) << 20, 0);
    const double sc2 = __hiloint2double((k2 + 1023) << 20, 0);
    const double res = (p * sc1) * sc2;
    return (x != x) ? x : res;                                            // fmin/fmax drop NaN: put it back
}

// The same for -708 <= x <= 708 (finite, no overflow / underflow): no clamp, no NaN restore, one exact power-of-two scale.
// Bit-identical to fast_exp on that range ((p 2^k1) 2^k2 = p 2^k exactly while everything stays normal).
CGP_MDEV double fast_exp_inrange(double x) {
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    const int k = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, -6.93147180369123816490e-01, x);
    r = fma(kf, -1.90821492927058770002e-10, r);
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p01 = 1. + r;
    const double p23 = fma(r, 1.6666666666666666e-01, 0.5);
    const double p45 = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    const double p67 = fma(r, 1.9841269841269841e-04, 1.3888888888888889e-03);
    const double p89 = fma(r, 2.7557319223985893e-06, 2.4801587301587302e-05);
    const double pab = fma(r, 2.5052108385441720e-08, 2.7557319223985888e-07);
    const double pcd = fma(r, 1.6059043836821613e-10, 2.0876756987868100e-09);
    const double q0 = fma(r2, p23, p01);
    const double q1 = fma(r2, p67, p45);
    const double q2 = fma(r2, pab, p89);
    const double s0 = fma(r4, q1, q0);
    const double s1 = fma(r4, pcd, q2);
    const double p = fma(r8, s1, s0);
    return p * __hiloint2double((k + 1023) << 20, 0);
}

// ---- log(v) for normal positive v (fdlibm's e_log.c kernel with the division replaced by fast_rcp); +inf -> +inf,
// NaN -> NaN.  Used for log(exp(x) + 1), where v >= 1.
CGP_MDEV double fast_log_pos(double v) {
    int hx = __double2hiint(v);
    const int lx = __double2loint(v);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    const double mth = __hiloint2double(hx | (i ^ 0x3ff00000), lx);      // m in [sqrt(2)/2, sqrt(2))
    k += i >> 20;
    const double dk = (double)k;
    const double f = mth - 1.;
    const double s = f * fast_rcp(2. + f);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01),
                              6.666666666666735130e-01);
    const double R = t1 + t2;
    const double hfsq = 0.5 * f * f;
    const double res = fma(dk, 6.93147180369123816490e-01, -((hfsq - fma(s, hfsq + R, dk * 1.90821492927058770002e-10)) - f));
    return (v < 1.7976931348623157e308) ? res : v;                       // inf / NaN pass through
}

// series branch of fast_softplus (valid for 3 <= x <= 700)
CGP_MDEV double softplus_series(double x) {
    const double u = fast_exp_inrange(-x);
    const double u2 = u * u, u4 = u2 * u2, u8 = u4 * u4;
    const double p0 = fma(u, -0.5, 1.);
    const double p1 = fma(u, -0.25, 3.3333333333333331e-01);
    const double p2 = fma(u, -1.6666666666666666e-01, 0.2);
    const double p3 = fma(u, -0.125, 1.4285714285714285e-01);
    const double p4 = fma(u, -0.1, 1.1111111111111110e-01);
    const double p5 = fma(u, -8.3333333333333329e-02, 9.0909090909090912e-02);
    const double q0 = fma(u2, p1, p0);
    const double q1 = fma(u2, p3, p2);
    const double q2 = fma(u2, p5, p4);
    const double s = fma(u8, q2, fma(u4, q1, q0));
    return fma(u, s, x);
}
// ---- softplus g(x) = log(exp(x) + 1) (models.py:50).  For x >= 3:  g = x + log1p(u), u = exp(-x) <= 0.05, with
// the alternating series of log1p (12 terms, < 1e-17 truncation).  Else (and for x > 700, where the reference's
// naive form overflows to +inf, and NaN) the reference's literal formula.
CGP_MDEV double softplus_general(double x) { return fast_log_pos(fast_exp(x) + 1.); }   // the reference's formula
CGP_MDEV double fast_softplus(double x) {
    if (!(x >= 3. && x <= 700.)) return softplus_general(x);
    return softplus_series(x);
}
// Warp-uniform variant: every lane of the (fully active) warp takes the same side, so the loop body of the
// sequential filters keeps one straight-line fast path (no divergence bookkeeping).
CGP_MDEV double fast_softplus_warp(double x) {
    if (__all_sync(0xffffffffu, x >= 3. && x <= 700.)) return softplus_series(x);
    return softplus_general(x);
}
// softplus and its derivative sigmoid(x) = e^x / (e^x + 1) = 1 / (1 + e^-x): the two sides of the range split, each
// branch-free (callers that evaluate several arguments at once pick the side once and interleave the evaluations)
CGP_MDEV void softplus_sigmoid_general(double x, double &g, double &sg) {
    const double ex = fast_exp(x), d = ex + 1.;
    g = fast_log_pos(d);
    sg = (ex < 1.7976931348623157e308) ? ex * fast_rcp(d) : 1.;         // e^x / (e^x + 1); inf / inf would be NaN
}
CGP_MDEV void softplus_sigmoid_series(double x, double &g, double &sg) {   // 3 <= x <= 700
    const double u = fast_exp_inrange(-x);
    const double u2 = u * u, u4 = u2 * u2, u8 = u4 * u4;
    const double p0 = fma(u, -0.5, 1.);
    const double p1 = fma(u, -0.25, 3.3333333333333331e-01);
    const double p2 = fma(u, -1.6666666666666666e-01, 0.2);
    const double p3 = fma(u, -0.125, 1.4285714285714285e-01);
    const double p4 = fma(u, -0.1, 1.1111111111111110e-01);
    const double p5 = fma(u, -8.3333333333333329e-02, 9.0909090909090912e-02);
    const double q0 = fma(u2, p1, p0);
    const double q1 = fma(u2, p3, p2);
    const double q2 = fma(u2, p5, p4);
    const double s = fma(u8, q2, fma(u4, q1, q0));
    g = fma(u, s, x);
    sg = fast_rcp(1. + u);
}
CGP_MDEV bool softplus_in_series_range(double x) { return x >= 3. && x <= 700.; }
CGP_MDEV void fast_softplus_sigmoid(double x, double &g, double &sg) {
    if (!softplus_in_series_range(x)) {
        softplus_sigmoid_general(x, g, sg);
        return;
    }
    softplus_sigmoid_series(x, g, sg);
}

// ---- sincos(x): Cody-Waite reduction by pi/2 in three FMA steps (accurate for |x| <= 1e9), fdlibm kernel
// polynomials (|r| <= pi/4) in Estrin form, branch-free.  |x| > 1e9 (an angle of more than 10^8 turns per sample
// -- only reachable after the filter has diverged) and non-finite x give NaN.
CGP_MDEV void fast_sincos(double x, double *sn, double *cs) {
    x = (fabs(x) <= 1.0e9) ? x : __longlong_as_double(0x7ff8000000000000LL);
    const double t = fma(x, 6.36619772367581382433e-01, 6755399441055744.0);   // round(x * 2/pi)
    const int q = __double2loint(t);
    const double j = t - 6755399441055744.0;
    double r = fma(j, -1.5707963267948966e+00, x);
    r = fma(j, -6.1232339957367660e-17, r);
    r = fma(j, 1.4973849048591698e-33, r);
    const double z = r * r, z2 = z * z;
    // sin kernel: r + r^3 (S1 + z S2 + z^2 (S3 + z S4) + z^4 (S5 + z S6))
    const double s12 = fma(z, 8.33333333332248946124e-03, -1.66666666666666324348e-01);
    const double s34 = fma(z, 2.75573137070700676789e-06, -1.98412698298579493134e-04);
    const double s56 = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    const double sp = fma(z2, fma(z2, s56, s34), s12);
    const double sr = fma(r * z, sp, r);
    // cos kernel: 1 - z/2 + z^2 (C1 + z C2 + z^2 (C3 + z C4) + z^4 (C5 + z C6))
    const double c12 = fma(z, -1.38888888888741095749e-03, 4.16666666666666019037e-02);
    const double c34 = fma(z, -2.75573143513906633035e-07, 2.48015872894767294178e-05);
    const double c56 = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    const double cp = fma(z2, fma(z2, c56, c34), c12);
    const double cr = fma(z2, cp, fma(z, -0.5, 1.));
    const bool swap = q & 1;
    const double a = swap ? cr : sr;      // |sin| candidate
    const double b = swap ? sr : cr;      // |cos| candidate
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b : b;
}

}  // namespace cgp
