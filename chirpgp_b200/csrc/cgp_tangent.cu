// cgp_tangent.cu -- forward-mode (tangent) derivative kernels for the negative log-likelihood of EVERY filter on the path:
// ekf, sgp_filter, cd_ekf, cd_sgp_filter on the chirp-family models.  The reference differentiates these objectives with
// jax.grad through lax.scan (demos/ghfs_mle.py:54-61, demos/cd_ekfs_mle.py, demos/cd_ghfs_mle.py,
// tetralith/jobs/{ghfs,cd_ekfs,cd_ghfs}_mle.py); with a handful of hyper-parameters (six) forward mode needs no checkpoints,
// no second sweep and no per-step storage: one group of lanes carries the filter state AND its derivative along ONE
// parameter direction, so n_dir directions are n_dir independent groups running side by side (the single-chirp demos leave
// the GPU empty otherwise).
//
// The filter step is written once over dual numbers (value, derivative): Cholesky factor, sigma points, model mean /
// drift (closed-form Jacobians for the EKF variants, differentiated once more by the dual arithmetic), RK4 stages
// (quadratures.py:34-54), measurement update (filters_smoothers.py:55-68) and nll increment (:44-45).  The value parts
// follow the operation order of the primal kernels (cgp_kernels.cuh) with the same elementary functions (cgp_math.cuh), so
// nll agrees with cgp_<filter>_f64 to rounding; the derivative parts are exact derivatives of those formulas.
//
// Inputs per direction k: the tangents of the kernel inputs, d consts / d theta_k, d m0 / d theta_k, d P0 / d theta_k,
// d Qc / d theta_k, d Xi / d theta_k -- the small map theta -> (consts, m0, P0, Qc) stays in host autodiff
// (chirpgp_b200/mle.py: torch.func.jacfwd of the model builder), exactly as for the adjoint kernel of cgp_nll2.cu.
#include <string.h>
#include <type_traits>
#include "cgp_dispatch.cuh"

namespace cgp {
namespace tng {

struct Dual {
    double v, d;
};
CGP_DEV Dual mk(double v, double d = 0.) { return Dual{v, d}; }
CGP_DEV Dual operator+(Dual a, Dual b) { return Dual{a.v + b.v, a.d + b.d}; }
CGP_DEV Dual operator-(Dual a, Dual b) { return Dual{a.v - b.v, a.d - b.d}; }
CGP_DEV Dual operator-(Dual a) { return Dual{-a.v, -a.d}; }
CGP_DEV Dual operator*(Dual a, Dual b) { return Dual{a.v * b.v, fma(a.v, b.d, a.d * b.v)}; }
CGP_DEV Dual operator*(double a, Dual b) { return Dual{a * b.v, a * b.d}; }
CGP_DEV Dual operator*(Dual a, double b) { return Dual{a.v * b, a.d * b}; }
CGP_DEV Dual operator+(Dual a, double b) { return Dual{a.v + b, a.d}; }
CGP_DEV Dual operator-(double a, Dual b) { return Dual{a - b.v, -b.d}; }
// a b + c with the value rounded like fma(a, b, c)
CGP_DEV Dual dfma(Dual a, Dual b, Dual c) { return Dual{fma(a.v, b.v, c.v), fma(a.v, b.d, fma(a.d, b.v, c.d))}; }
CGP_DEV Dual dfma(double a, Dual b, Dual c) { return Dual{fma(a, b.v, c.v), fma(a, b.d, c.d)}; }
CGP_DEV Dual drcp(Dual a) {
    const double r = fast_rcp(a.v);
    return Dual{r, -(a.d * r) * r};
}
CGP_DEV Dual dsqrt(Dual a) {
    const double s = sqrt(a.v);
    return Dual{s, 0.5 * a.d / s};
}
CGP_DEV Dual drsqrt(Dual a) {                          // a^{-1/2}
    const double r = fast_rsqrt(a.v);
    return Dual{r, -0.5 * a.d * r * (r * r)};
}
CGP_DEV Dual dlog(Dual a) { return Dual{fast_log_pos(a.v), a.d * fast_rcp(a.v)}; }
CGP_DEV void dsoftplus_sigmoid(Dual x, Dual &g, Dual &sg) {
    double gv, s;
    fast_softplus_sigmoid(x.v, gv, s);
    g = Dual{gv, s * x.d};
    sg = Dual{s, (s * (1. - s)) * x.d};
}
CGP_DEV void dsincos(Dual x, Dual &sn, Dual &cs) {
    double s, c;
    fast_sincos(x.v, &s, &c);
    sn = Dual{s, c * x.d};
    cs = Dual{c, -s * x.d};
}
template <int G> CGP_DEV Dual group_sum(Dual a) { return Dual{group_allreduce<G>(a.v), group_allreduce<G>(a.d)}; }

enum { KIND_EKF = 0, KIND_SGP = 1, KIND_CD_EKF = 2, KIND_CD_SGP = 3 };

// ---- models over dual numbers -------------------------------------------------------------------------------------
template <int NH> struct LcdD {                          // models.py:295-309, :369-384 (ModelLCD)
    static constexpr int D = 2 * NH + 2, V = D - 2;
    Dual e, f00, f01, f10, f11, q, s00, s01, s11;
    double fs, dt;
    CGP_DEV void load(const double *c, const double *cd, double dt_) {
        auto at = [&](int i) { return Dual{c[i], cd ? cd[i] : 0.}; };
        e = at(0); f00 = at(1); f01 = at(2); f10 = at(3); f11 = at(4); q = at(5); s00 = at(6); s01 = at(7); s11 = at(8);
        fs = c[9]; dt = dt_;
    }
    CGP_DEV Dual sig(int r, int c) const {
        if (r == c) return r < V ? q : (r == V ? s00 : s11);
        return s01;
    }
    static CGP_DEV constexpr bool has_sig(int r, int c) { return (r == c) || (r == V && c == V + 1) || (r == V + 1 && c == V); }
    CGP_DEV void mean(const Dual (&u)[D], Dual (&m)[D]) const {
        Dual gv, sg;
        dsoftplus_sigmoid(u[V], gv, sg);
        const Dual w = (kTwoPi * gv) * fs;
        CGP_UNROLL for (int k = 0; k < NH; k++) {
            Dual sn, cs;
            dsincos((dt * (double)(k + 1)) * w, sn, cs);
            const Dual ce = cs * e, se = sn * e;
            m[2 * k] = dfma(-se, u[2 * k + 1], ce * u[2 * k]);
            m[2 * k + 1] = dfma(ce, u[2 * k + 1], se * u[2 * k]);
        }
        m[V] = dfma(f01, u[V + 1], f00 * u[V]);
        m[V + 1] = dfma(f11, u[V + 1], f10 * u[V]);
    }
    CGP_DEV void mean_jac(const Dual (&u)[D], Dual (&m)[D], Dual (&J)[D][D]) const {
        Dual gv, sg;
        dsoftplus_sigmoid(u[V], gv, sg);
        const Dual w = (kTwoPi * gv) * fs, dw = (kTwoPi * sg) * fs;
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) J[r][c] = mk(0.);
        CGP_UNROLL for (int k = 0; k < NH; k++) {
            const double dtk = dt * (double)(k + 1);
            Dual sn, cs;
            dsincos(dtk * w, sn, cs);
            const Dual ce = cs * e, se = sn * e, dth = dtk * dw;
            m[2 * k] = dfma(-se, u[2 * k + 1], ce * u[2 * k]);
            m[2 * k + 1] = dfma(ce, u[2 * k + 1], se * u[2 * k]);
            J[2 * k][2 * k] = ce;     J[2 * k][2 * k + 1] = -se;
            J[2 * k + 1][2 * k] = se; J[2 * k + 1][2 * k + 1] = ce;
            J[2 * k][V] = -m[2 * k + 1] * dth;
            J[2 * k + 1][V] = m[2 * k] * dth;
        }
        m[V] = dfma(f01, u[V + 1], f00 * u[V]);
        m[V + 1] = dfma(f11, u[V + 1], f10 * u[V]);
        J[V][V] = f00; J[V][V + 1] = f01; J[V + 1][V] = f10; J[V + 1][V + 1] = f11;
    }
};
template <int NH> struct SdeD {                          // models.py:104-110, :164-168 (ModelSDE)
    static constexpr int D = 2 * NH + 2, V = D - 2;
    Dual lam, g2, tg;
    double fs;
    CGP_DEV void load(const double *c, const double *cd) {
        lam = Dual{c[0], cd ? cd[0] : 0.}; g2 = Dual{c[1], cd ? cd[1] : 0.}; tg = Dual{c[2], cd ? cd[2] : 0.};
        fs = c[3];
    }
    CGP_DEV void drift_w(Dual w, const Dual (&u)[D], Dual (&a)[D]) const {
        CGP_UNROLL for (int k = 0; k < NH; k++) {
            const Dual wk = w * (double)(k + 1);
            a[2 * k] = dfma(-wk, u[2 * k + 1], -lam * u[2 * k]);
            a[2 * k + 1] = dfma(-lam, u[2 * k + 1], wk * u[2 * k]);
        }
        a[V] = u[V + 1];
        a[V + 1] = dfma(-tg, u[V + 1], -g2 * u[V]);
    }
    CGP_DEV void drift(const Dual (&u)[D], Dual (&a)[D]) const {
        Dual gv, sg;
        dsoftplus_sigmoid(u[V], gv, sg);
        drift_w((kTwoPi * gv) * fs, u, a);
    }
    CGP_DEV void drift_jac(const Dual (&u)[D], Dual (&a)[D], Dual (&J)[D][D]) const {
        Dual gv, sg;
        dsoftplus_sigmoid(u[V], gv, sg);
        const Dual w = (kTwoPi * gv) * fs, dw = (kTwoPi * sg) * fs;
        drift_w(w, u, a);
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) J[r][c] = mk(0.);
        CGP_UNROLL for (int k = 0; k < NH; k++) {
            const Dual wk = w * (double)(k + 1), dwk = dw * (double)(k + 1);
            J[2 * k][2 * k] = -lam;     J[2 * k][2 * k + 1] = -wk;
            J[2 * k + 1][2 * k] = wk;   J[2 * k + 1][2 * k + 1] = -lam;
            J[2 * k][V] = -dwk * u[2 * k + 1];
            J[2 * k + 1][V] = dwk * u[2 * k];
        }
        J[V][V + 1] = mk(1.);
        J[V + 1][V] = -g2;
        J[V + 1][V + 1] = -tg;
    }
};

// ---- pieces of the filters over dual numbers ----------------------------------------------------------------------
template <int D> CGP_DEV void chol_d(const Dual (&P)[NSym<D>::value], Dual (&L)[NSym<D>::value]) {     // chol_lower_sym_rsqrt
    CGP_UNROLL for (int j = 0; j < D; j++) {
        Dual s = P[sidx(j, j)];
        CGP_UNROLL for (int k = 0; k < j; k++) s = dfma(-L[sidx(j, k)], L[sidx(j, k)], s);
        const Dual r = drsqrt(s);
        L[sidx(j, j)] = s * r;
        CGP_UNROLL for (int i = j + 1; i < D; i++) {
            Dual t = P[sidx(i, j)];
            CGP_UNROLL for (int k = 0; k < j; k++) t = dfma(-L[sidx(i, k)], L[sidx(j, k)], t);
            L[sidx(i, j)] = t * r;
        }
    }
}
template <int D>
CGP_DEV void sigma_point(const Dual (&m)[D], const Dual (&L)[NSym<D>::value], const double *__restrict__ xi, Dual (&chi)[D]) {
    CGP_UNROLL for (int r = 0; r < D; r++) {
        Dual s = L[sidx(r, 0)] * __ldg(xi);
        CGP_UNROLL for (int c = 1; c <= r; c++) s = dfma(__ldg(xi + c), L[sidx(r, c)], s);
        chi[r] = m[r] + s;
    }
}
// filters_smoothers.py:55-68 on packed covariances; returns the nll increment (:44-45: sc = sqrt(S), (log(2 pi sc^2) + r^2/sc^2)/2)
template <int D>
CGP_DEV Dual update_d(const Dual (&mp)[D], const Dual (&Pp)[NSym<D>::value], const double (&H)[D], Dual Xi, double y, Dual (&mf)[D],
                      Dual (&Pf)[NSym<D>::value]) {
    Dual PH[D];
    CGP_UNROLL for (int i = 0; i < D; i++) {
        Dual s = Pp[sidx(i, 0)] * H[0];
        CGP_UNROLL for (int j = 1; j < D; j++) s = dfma(H[j], Pp[sidx(i, j)], s);
        PH[i] = s;
    }
    Dual S = PH[0] * H[0];
    CGP_UNROLL for (int j = 1; j < D; j++) S = dfma(H[j], PH[j], S);
    S = S + Xi;
    const Dual rS = drcp(S);
    Dual pred = mp[0] * H[0];
    CGP_UNROLL for (int i = 1; i < D; i++) pred = dfma(H[i], mp[i], pred);
    const Dual r = y - pred;
    Dual K[D];
    CGP_UNROLL for (int i = 0; i < D; i++) K[i] = PH[i] * rS;
    CGP_UNROLL for (int i = 0; i < D; i++) mf[i] = dfma(K[i], r, mp[i]);
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j <= i; j++) Pf[sidx(i, j)] = dfma(-K[i], PH[j], Pp[sidx(i, j)]);
    const Dual sc = dsqrt(S), sc2 = sc * sc;
    return dfma(r * r, drcp(sc2), dlog(kTwoPi * sc2)) * 0.5;
}
template <int D, class Ode> CGP_DEV void rk4_d(Ode &&ode, Dual (&m)[D], Dual (&P)[NSym<D>::value], double dt) {    // quadratures.py:34-54
    constexpr int NS = NSym<D>::value;
    Dual km[D], kP[NS], am[D], aP[NS], tm[D], tP[NS];
    ode(m, P, km, kP);
    CGP_UNROLL for (int i = 0; i < D; i++) { am[i] = km[i]; tm[i] = m[i] + (dt * km[i]) * 0.5; }
    CGP_UNROLL for (int i = 0; i < NS; i++) { aP[i] = kP[i]; tP[i] = P[i] + (dt * kP[i]) * 0.5; }
    ode(tm, tP, km, kP);
    CGP_UNROLL for (int i = 0; i < D; i++) { am[i] = am[i] + 2. * km[i]; tm[i] = m[i] + (dt * km[i]) * 0.5; }
    CGP_UNROLL for (int i = 0; i < NS; i++) { aP[i] = aP[i] + 2. * kP[i]; tP[i] = P[i] + (dt * kP[i]) * 0.5; }
    ode(tm, tP, km, kP);
    CGP_UNROLL for (int i = 0; i < D; i++) { am[i] = am[i] + 2. * km[i]; tm[i] = m[i] + dt * km[i]; }
    CGP_UNROLL for (int i = 0; i < NS; i++) { aP[i] = aP[i] + 2. * kP[i]; tP[i] = P[i] + dt * kP[i]; }
    ode(tm, tP, km, kP);
    constexpr double kSixth = 1. / 6.;
    CGP_UNROLL for (int i = 0; i < D; i++) m[i] = m[i] + (dt * (am[i] + km[i])) * kSixth;
    CGP_UNROLL for (int i = 0; i < NS; i++) P[i] = P[i] + (dt * (aP[i] + kP[i])) * kSixth;
}

struct TangentIO {
    const double *ys;
    int n_dir;
    const double *consts_dot; int64_t consts_dot_stride;      // [B|1, n_dir, NC]
    const double *m0_dot;     int64_t m0_dot_stride;          // [B|1, n_dir, d]
    const double *P0_dot;     int64_t P0_dot_stride;          // [B|1, n_dir, d, d]
    const double *Qc_dot;     int64_t Qc_dot_stride;          // [B|1, n_dir, d, d]   (CD filters)
    const double *Xi_dot;                                     // [n_dir] or NULL
    double *nll;                                              // [B]
    double *nll_dot;                                          // [B, n_dir]
};

template <int D> CGP_DEV void load_sym_d(const double *v, const double *d, Dual (&P)[NSym<D>::value]) {
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++)
        P[sidx(r, c)] = Dual{v[r * D + c], d ? 0.5 * (d[r * D + c] + d[c * D + r]) : 0.};
}

// One group of G lanes per (problem, direction).  G = 1 for the EKF variants (no sigma points), 16 / 32 for the sigma-point
// filters (points dealt round-robin, partial sums combined by a butterfly all-reduce so that every lane keeps the
// bit-identical replica of the state).
template <int NH, int KIND, int G>
__global__ void __launch_bounds__(128) tangent_kernel(const CgpProblem p, const TangentIO io) {
    constexpr int D = 2 * NH + 2, NS = NSym<D>::value;
    constexpr bool CD = (KIND == KIND_CD_EKF || KIND == KIND_CD_SGP);
    using Model = typename std::conditional<CD, SdeD<NH>, LcdD<NH>>::type;
    int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    const bool live = gid < p.B * io.n_dir;                 // surplus groups of the last warp redo the last problem and
    if (!live) gid = p.B * io.n_dir - 1;                    // discard it (the shuffles below are warp-wide)
    const int64_t b = gid / io.n_dir;
    const int k = (int)(gid % io.n_dir);
    Model mdl;
    if constexpr (CD) {
        mdl.load(p.consts + b * p.consts_stride, io.consts_dot ? io.consts_dot + b * io.consts_dot_stride + (int64_t)k * CGP_NC_SDE : nullptr);
    } else {
        mdl.load(p.consts + b * p.consts_stride, io.consts_dot ? io.consts_dot + b * io.consts_dot_stride + (int64_t)k * CGP_NC_LCD : nullptr, p.dt);
    }
    Dual m[D], P[NS], Qc[NS];
    {
        const double *m0 = p.m0 + b * p.m0_stride;
        const double *m0d = io.m0_dot ? io.m0_dot + b * io.m0_dot_stride + (int64_t)k * D : nullptr;
        CGP_UNROLL for (int i = 0; i < D; i++) m[i] = Dual{m0[i], m0d ? m0d[i] : 0.};
        load_sym_d<D>(p.P0 + b * p.P0_stride, io.P0_dot ? io.P0_dot + b * io.P0_dot_stride + (int64_t)k * D * D : nullptr, P);
        if constexpr (CD)
            load_sym_d<D>(p.Qc + b * p.Qc_stride, io.Qc_dot ? io.Qc_dot + b * io.Qc_dot_stride + (int64_t)k * D * D : nullptr, Qc);
    }
    double H[D];
    CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
    const Dual Xi = Dual{p.Xi, io.Xi_dot ? io.Xi_dot[k] : 0.};
    const double *__restrict__ y = io.ys + (b / p.ys_repeat) * p.T;
    const double *__restrict__ sw = p.sig_w, *__restrict__ sxi = p.sig_xi;
    const int n = p.n_sigma;
    const double dt = p.dt;
    Dual acc = mk(0.);
    for (int64_t t = 0; t < p.T; t++) {
        const double yt = __ldg(y + t);
        Dual mp[D], Pp[NS];
        if constexpr (KIND == KIND_EKF) {                                        // filters_smoothers.py:255-257
            Dual J[D][D], JP[D][D];
            mdl.mean_jac(m, mp, J);
            CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j < D; j++) {
                Dual s = J[i][0] * P[sidx(0, j)];
                CGP_UNROLL for (int q = 1; q < D; q++) s = dfma(J[i][q], P[sidx(q, j)], s);
                JP[i][j] = s;
            }
            CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j <= i; j++) {
                Dual s = JP[i][0] * J[j][0];
                CGP_UNROLL for (int q = 1; q < D; q++) s = dfma(JP[i][q], J[j][q], s);
                Pp[sidx(i, j)] = Model::has_sig(i, j) ? s + mdl.sig(i, j) : s;
            }
        } else if constexpr (KIND == KIND_SGP) {                                 // :88-121
            Dual L[NS], am[D], aP[NS];
            chol_d<D>(P, L);
            CGP_UNROLL for (int i = 0; i < D; i++) am[i] = mk(0.);
            CGP_UNROLL for (int i = 0; i < NS; i++) aP[i] = mk(0.);
            for (int i = lane; i < n; i += G) {
                Dual chi[D], ev[D];
                sigma_point<D>(m, L, sxi + (int64_t)i * D, chi);
                mdl.mean(chi, ev);
                const double w = __ldg(sw + i);
                CGP_UNROLL for (int r = 0; r < D; r++) am[r] = dfma(w, ev[r], am[r]);
                CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++) {
                    Dual v = ev[r] * ev[c];
                    if (Model::has_sig(r, c)) v = v + mdl.sig(r, c);
                    aP[sidx(r, c)] = dfma(w, v, aP[sidx(r, c)]);
                }
            }
            CGP_UNROLL for (int r = 0; r < D; r++) mp[r] = group_sum<G>(am[r]);
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++)
                Pp[sidx(r, c)] = group_sum<G>(aP[sidx(r, c)]) - mp[r] * mp[c];
        } else if constexpr (KIND == KIND_CD_EKF) {                              // :384-385 + RK4
            CGP_UNROLL for (int i = 0; i < D; i++) mp[i] = m[i];
            CGP_UNROLL for (int i = 0; i < NS; i++) Pp[i] = P[i];
            rk4_d<D>([&](const Dual (&mm)[D], const Dual (&PP)[NS], Dual (&dm)[D], Dual (&dP)[NS]) {
                Dual J[D][D], X[D][D];
                mdl.drift_jac(mm, dm, J);
                CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j < D; j++) {
                    Dual s = J[i][0] * PP[sidx(0, j)];
                    CGP_UNROLL for (int q = 1; q < D; q++) s = dfma(J[i][q], PP[sidx(q, j)], s);
                    X[i][j] = s;
                }
                CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++)
                    dP[sidx(r, c)] = (X[c][r] + X[r][c]) + Qc[sidx(r, c)];
            }, mp, Pp, dt);
        } else {                                                                 // :124-137 + RK4
            CGP_UNROLL for (int i = 0; i < D; i++) mp[i] = m[i];
            CGP_UNROLL for (int i = 0; i < NS; i++) Pp[i] = P[i];
            rk4_d<D>([&](const Dual (&mm)[D], const Dual (&PP)[NS], Dual (&dm)[D], Dual (&dP)[NS]) {
                Dual L[NS], am[D], aQ[D][D];
                chol_d<D>(PP, L);
                CGP_UNROLL for (int i = 0; i < D; i++) am[i] = mk(0.);
                CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) aQ[r][c] = mk(0.);
                for (int i = lane; i < n; i += G) {
                    Dual chi[D], f[D];
                    sigma_point<D>(mm, L, sxi + (int64_t)i * D, chi);
                    mdl.drift(chi, f);
                    const double w = __ldg(sw + i);
                    CGP_UNROLL for (int r = 0; r < D; r++) am[r] = dfma(w, f[r], am[r]);
                    CGP_UNROLL for (int r = 0; r < D; r++) {
                        const Dual dr = chi[r] - mm[r];
                        CGP_UNROLL for (int c = 0; c < D; c++) aQ[r][c] = dfma(w, dr * f[c], aQ[r][c]);
                    }
                }
                CGP_UNROLL for (int r = 0; r < D; r++) dm[r] = group_sum<G>(am[r]);
                CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) aQ[r][c] = group_sum<G>(aQ[r][c]);
                CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++)
                    dP[sidx(r, c)] = (aQ[r][c] + aQ[c][r]) + Qc[sidx(r, c)];
            }, mp, Pp, dt);
        }
        acc = acc + update_d<D>(mp, Pp, H, Xi, yt, m, P);
    }
    if (lane == 0 && live) {
        if (k == 0 && io.nll) io.nll[b] = acc.v;
        io.nll_dot[b * io.n_dir + k] = acc.d;
    }
}

template <int NH, int KIND, int G>
static int launch_one(const CgpProblem &p, const TangentIO &io, cudaStream_t s) {
    const int64_t threads = p.B * io.n_dir * G;
    const unsigned grid = (unsigned)ceil_div(threads, 128);
    tangent_kernel<NH, KIND, G><<<grid, 128, 0, s>>>(p, io);
    return check_launch();
}
template <int NH> static int launch_nh(const CgpProblem &p, int kind, const TangentIO &io, cudaStream_t s) {
    switch (kind) {
        case KIND_EKF: return launch_one<NH, KIND_EKF, 1>(p, io, s);
        case KIND_CD_EKF: return launch_one<NH, KIND_CD_EKF, 1>(p, io, s);
        case KIND_SGP: return p.n_sigma > 16 ? launch_one<NH, KIND_SGP, 32>(p, io, s) : launch_one<NH, KIND_SGP, 16>(p, io, s);
        case KIND_CD_SGP: return p.n_sigma > 16 ? launch_one<NH, KIND_CD_SGP, 32>(p, io, s) : launch_one<NH, KIND_CD_SGP, 16>(p, io, s);
        default: return CGP_ERR_BAD_ARG;
    }
}

}  // namespace tng
}  // namespace cgp

using namespace cgp;

extern "C" int cgp_filter_nll_tangent_f64(const char *filter, const CgpProblem *p, const double *ys, int n_dir,
                                          const double *consts_dot, int64_t consts_dot_stride, const double *m0_dot,
                                          int64_t m0_dot_stride, const double *P0_dot, int64_t P0_dot_stride,
                                          const double *Qc_dot, int64_t Qc_dot_stride, const double *Xi_dot, double *nll,
                                          double *nll_dot, void *stream) {
    if (!filter || !p || !ys || !nll_dot || n_dir < 1 || p->B < 1 || p->T < 1 || !p->consts || !p->m0 || !p->P0 || !p->H ||
        p->ys_repeat < 1)
        return CGP_ERR_BAD_ARG;
    int kind;
    if (!strcmp(filter, "ekf")) kind = tng::KIND_EKF;
    else if (!strcmp(filter, "sgp_filter")) kind = tng::KIND_SGP;
    else if (!strcmp(filter, "cd_ekf")) kind = tng::KIND_CD_EKF;
    else if (!strcmp(filter, "cd_sgp_filter")) kind = tng::KIND_CD_SGP;
    else return CGP_ERR_BAD_ARG;
    const bool cd = kind == tng::KIND_CD_EKF || kind == tng::KIND_CD_SGP;
    const bool sg = kind == tng::KIND_SGP || kind == tng::KIND_CD_SGP;
    if (p->model != (cd ? CGP_MODEL_SDE : CGP_MODEL_LCD)) return CGP_ERR_UNSUPPORTED;
    if (p->d != 2 * p->num_harmonics + 2) return CGP_ERR_BAD_ARG;
    if (cd && !p->Qc) return CGP_ERR_BAD_ARG;
    if (sg && (p->n_sigma < 1 || !p->sig_w || !p->sig_xi)) return CGP_ERR_BAD_ARG;
    tng::TangentIO io{ys, n_dir, consts_dot, consts_dot_stride, m0_dot, m0_dot_stride, P0_dot, P0_dot_stride,
                      Qc_dot, Qc_dot_stride, Xi_dot, nll, nll_dot};
    cudaStream_t s = (cudaStream_t)stream;
    switch (p->num_harmonics) {
        case 1: return tng::launch_nh<1>(*p, kind, io, s);
        case 2: return tng::launch_nh<2>(*p, kind, io, s);
        case 3: return tng::launch_nh<3>(*p, kind, io, s);
        default: return CGP_ERR_UNSUPPORTED;
    }
}
