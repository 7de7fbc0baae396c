// cgp_kernels.cuh -- kernel templates of the chirpgp_b200 hot path (sm_100a, FP64 SIMT).
//
// Work decomposition (see DESIGN.md "Kernels"):
//   * sequential filters without sigma points (kf / ekf / cd_ekf): one THREAD owns one chirp and walks the
//     time loop with mean and covariance in registers;
//   * sigma-point filters (sgp_filter / cd_sgp_filter) and the sequential CD sigma-point smoother: a GROUP of
//     G lanes (8/16/32) owns one chirp; every lane keeps a replica of mean/covariance/Cholesky factor in
//     registers, sigma points are dealt round-robin to lanes and the weighted sums are combined with a
//     butterfly __shfl_xor all-reduce (every lane gets the bit-identical sum, so replicas never diverge);
//   * discrete smoothers (rts / eks / sgp_smoother) are split into a TIME-PARALLEL gain kernel (one thread per
//     (chirp, step): smoother gain G_k, predicted mean/cov -- these depend on the filtering result only) and
//     a sequential sweep kernel.  The reference's step (filters_smoothers.py:83-84)
//         ms = mf + G (ms' - mp),   Ps = Pf + G (Ps' - Pp) G^T
//     is regrouped so that everything that does not depend on (ms', Ps') is formed in the time-parallel half:
//         c = mf - G mp,   C = Pf - G Pp G^T = Pf - G D^T   (G Pp = D),      ms = c + G ms',   Ps = C + G Ps' G^T.
//     The workspace record [G | c | C packed] is 30 doubles at d = 4 (it was [G | mp | Pp], 36) and the sweep no longer
//     reads (mf, Pf): 400 instead of 608 bytes per step through the sweep.  C is the covariance of x_k given x_{k+1}
//     (positive semi-definite), so Ps is a sum of two PSD terms; the regrouped recursion differs from the reference's
//     order by ~1e-13 over T = 3141 (profiles/scripts/regroup_noise.py), three orders below the reference's own
//     summation-order noise.
#pragma once
#include "cgp_device.cuh"

namespace cgp {

struct FilterIO {
    const double *__restrict__ ys;
    double *__restrict__ mfs;
    double *__restrict__ Pfs;
    double *__restrict__ nell;
    int nell_last_only;
    double *__restrict__ ws = nullptr;     // fused filter + smoother gains: SmootherIO::ws records, filled by the filter
};
struct SmootherIO {
    const double *__restrict__ mfs;
    const double *__restrict__ Pfs;
    double *__restrict__ mss;
    double *__restrict__ Pss;
    double *__restrict__ ws;     // per (chirp, step): [G d*d | c d | C packed lower d(d+1)/2 | pad to an even count]
};

// doubles per smoother record (even, so that records stay 16-byte aligned): d = 4: 30, d = 6: 64, d = 8: 108
__host__ __device__ constexpr int ws_record_doubles(int d) { return (d * d + d + d * (d + 1) / 2 + 1) & ~1; }
template <int D> CGP_DEV constexpr int ws_record() { return ws_record_doubles(D); }

// ================================================================================================ thread-per-chirp filters
// kf (filters_smoothers.py:145-184) and ekf (:222-264): Model = ModelLinearDisc<D> | ModelLCD<NH>
// WIDE: outputs are 32-byte aligned and d % 4 == 0 -> every lane writes whole 32-byte sectors (STG.256)
template <class Model, bool WIDE = false>
__global__ void __launch_bounds__(128) ekf_thread_kernel(const CgpProblem p, const FilterIO io) {
    constexpr int D = Model::D;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    double m[D], P[D][D], H[D];
    load_vec<D>(p.m0 + b * p.m0_stride, m);
    load_mat<D>(p.P0 + b * p.P0_stride, P);
    CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
    const double *__restrict__ y = io.ys + (b / p.ys_repeat) * p.T;
    const int64_t T = p.T;
    const bool store = io.mfs != nullptr;
    double acc = 0.;
    NellRowWriter nellw;
    if (io.nell && !io.nell_last_only) nellw.init(io.nell, b, T);
    double ynext = __ldg(y);
    for (int64_t t = 0; t < T; t++) {
        const double yt = ynext;
        if (t + 1 < T) ynext = __ldg(y + t + 1);
        double mp[D], J[D][D], JP[D][D], Pp[D][D];
        mdl.mean_jac(m, mp, J);                       // :255-256
        jmul<Model, D>(J, P, JP);
        mul_jt<Model, D>(JP, J, Pp);                  // :257
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++)
            if (Model::has_sig(r, c)) Pp[r][c] += mdl.sig(r, c);
        acc = acc + linear_update<D>(mp, Pp, H, p.Xi, yt, m, P);
        if (store) {
            gstore_vec<WIDE, D>(io.mfs + (b * T + t) * D, m);
            gstore_mat<WIDE, D>(io.Pfs + (b * T + t) * (D * D), P);
        }
        if (io.nell && !io.nell_last_only) nellw.put(t, T, acc);
    }
    if (io.nell && io.nell_last_only) io.nell[b] = acc;
}

// ekf_for_kpt (filters_smoothers.py:267-314) with the measurement function of build_kpt_chirp_model (models.py:572-578):
// state x = [omega, a_1 .. a_NH, phase], d = NH + 2; linear prediction with (F, Sigma) (consts = [F | Sigma]);
//   h(x) = sum_k a_k sin(k g(x_0 + x_{d-1})),   H = jacfwd(h)(mp) in closed form:
//   dh/dx_0 = dh/dx_{d-1} = g'(x_0 + x_{d-1}) sum_k k a_k cos(k phi),   dh/da_k = sin(k phi).       One thread per chirp.
template <int NH>
__global__ void __launch_bounds__(128) ekf_kpt_thread_kernel(const CgpProblem p, const FilterIO io) {
    constexpr int D = NH + 2;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    ModelLinearDisc<D> mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    double m[D], P[D][D];
    load_vec<D>(p.m0 + b * p.m0_stride, m);
    load_mat<D>(p.P0 + b * p.P0_stride, P);
    const double *__restrict__ y = io.ys + (b / p.ys_repeat) * p.T;
    const int64_t T = p.T;
    const bool store = io.mfs != nullptr;
    double acc = 0.;
    NellRowWriter nellw;
    if (io.nell && !io.nell_last_only) nellw.init(io.nell, b, T);
    double ynext = __ldg(y);
    for (int64_t t = 0; t < T; t++) {
        const double yt = ynext;
        if (t + 1 < T) ynext = __ldg(y + t + 1);
        double mp[D], FP[D][D], Pp[D][D], H[D];
        matvec<D>(mdl.F, m, mp);                      // _linear_predict :48-52
        matmul<D>(mdl.F, P, FP);
        matmul_nt<D>(FP, mdl.F, Pp);
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) Pp[r][c] += mdl.Sg[r][c];
        double phi, dphi;
        softplus_and_sigmoid(mp[0] + mp[D - 1], phi, dphi);
        double pred = 0., dsum = 0.;
        CGP_UNROLL for (int k = 1; k <= NH; k++) {
            double sn, cs;
            fast_sincos(phi * (double)k, &sn, &cs);
            pred = (k == 1) ? mp[k] * sn : fma(mp[k], sn, pred);
            dsum = (k == 1) ? mp[k] * (cs * (double)k) : fma(mp[k], cs * (double)k, dsum);
            H[k] = sn;
        }
        H[0] = dsum * dphi;
        H[D - 1] = H[0];
        acc = acc + nonlinear_update<D>(mp, Pp, H, pred, p.Xi, yt, m, P);
        if (store) {
            gstore_vec_auto<D>(io.mfs + (b * T + t) * D, m);
            gstore_mat_auto<D>(io.Pfs + (b * T + t) * (D * D), P);
        }
        if (io.nell && !io.nell_last_only) nellw.put(t, T, acc);
    }
    if (io.nell && io.nell_last_only) io.nell[b] = acc;
}

// ---- continuous-discrete pieces on packed-symmetric covariances --------------------------------
// rhs of the CD-EKF moment ODE (filters_smoothers.py:384-385): dm = a(m), dP = P J^T + J P + b b^T.
// With P exactly symmetric (P J^T)_rc == (J P)_cr bit for bit, so one product X = J P suffices.
template <class Model>
CGP_DEV void cd_ekf_ode(const Model &mdl, const double (&Qc)[NSym<Model::D>::value], const double (&m)[Model::D],
                        const double (&P)[NSym<Model::D>::value], double (&dm)[Model::D],
                        double (&dP)[NSym<Model::D>::value]) {
    constexpr int D = Model::D;
    double J[D][D], X[D][D];
    mdl.drift_jac(m, dm, J);
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j < D; j++) {
        double s = J[i][0] * P[sidx(0, j)];
        CGP_UNROLL for (int k = 1; k < D; k++) s = fma(J[i][k], P[sidx(k, j)], s);
        X[i][j] = s;
    }
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++)
        dP[sidx(r, c)] = (X[c][r] + X[r][c]) + Qc[sidx(r, c)];
}

// One classic RK4 step of size dt on the pair (m, P) (quadratures.py:34-54 / :57-81): exactly one step per
// measurement interval, 4 rhs evaluations.  `ode(m, P, dm, dP)` is any callable.
template <int D, class Ode>
CGP_DEV void rk4_step(Ode &&ode, double (&m)[D], double (&P)[NSym<D>::value], double dt) {
    constexpr int NS = NSym<D>::value;
    double km[D], kP[NS], am[D], aP[NS], tm[D], tP[NS];
    ode(m, P, km, kP);
    CGP_UNROLL for (int i = 0; i < D; i++) { am[i] = km[i]; tm[i] = m[i] + dt * km[i] * 0.5; }
    CGP_UNROLL for (int i = 0; i < NS; i++) { aP[i] = kP[i]; tP[i] = P[i] + dt * kP[i] * 0.5; }
    ode(tm, tP, km, kP);
    CGP_UNROLL for (int i = 0; i < D; i++) { am[i] = am[i] + 2 * km[i]; tm[i] = m[i] + dt * km[i] * 0.5; }
    CGP_UNROLL for (int i = 0; i < NS; i++) { aP[i] = aP[i] + 2 * kP[i]; tP[i] = P[i] + dt * kP[i] * 0.5; }
    ode(tm, tP, km, kP);
    CGP_UNROLL for (int i = 0; i < D; i++) { am[i] = am[i] + 2 * km[i]; tm[i] = m[i] + dt * km[i]; }
    CGP_UNROLL for (int i = 0; i < NS; i++) { aP[i] = aP[i] + 2 * kP[i]; tP[i] = P[i] + dt * kP[i]; }
    ode(tm, tP, km, kP);
    // dt (k1 + 2 k2 + 2 k3 + k4) / 6 with the division by 6 as a multiplication by the rounded reciprocal (<= 1 ulp)
    constexpr double kSixth = 1. / 6.;
    CGP_UNROLL for (int i = 0; i < D; i++) m[i] = m[i] + dt * (am[i] + km[i]) * kSixth;
    CGP_UNROLL for (int i = 0; i < NS; i++) P[i] = P[i] + dt * (aP[i] + kP[i]) * kSixth;
}

// cd_ekf (filters_smoothers.py:352-397): Model = ModelLinearSDE<D> | ModelSDE<NH>
template <class Model>
__global__ void __launch_bounds__(128) cd_ekf_thread_kernel(const CgpProblem p, const FilterIO io) {
    constexpr int D = Model::D, NS = NSym<D>::value;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride);
    double m[D], P[NS], Qc[NS], H[D];
    load_vec<D>(p.m0 + b * p.m0_stride, m);
    load_sym<D>(p.P0 + b * p.P0_stride, P);
    load_sym<D>(p.Qc + b * p.Qc_stride, Qc);
    CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
    const double *__restrict__ y = io.ys + (b / p.ys_repeat) * p.T;
    const int64_t T = p.T;
    const bool store = io.mfs != nullptr;
    const double dt = p.dt;
    double acc = 0.;
    NellRowWriter nellw;
    if (io.nell && !io.nell_last_only) nellw.init(io.nell, b, T);
    double ynext = __ldg(y);
    for (int64_t t = 0; t < T; t++) {
        const double yt = ynext;
        if (t + 1 < T) ynext = __ldg(y + t + 1);
        rk4_step<D>([&](const double (&mm)[D], const double (&PP)[NS], double (&dm)[D], double (&dP)[NS]) {
            cd_ekf_ode<Model>(mdl, Qc, mm, PP, dm, dP);
        }, m, P, dt);
        double mf[D], Pf[NS];
        acc = acc + linear_update_sym<D>(m, P, H, p.Xi, yt, mf, Pf);
        CGP_UNROLL for (int i = 0; i < D; i++) m[i] = mf[i];
        CGP_UNROLL for (int i = 0; i < NS; i++) P[i] = Pf[i];
        if (store) {
            gstore_vec_auto<D>(io.mfs + (b * T + t) * D, m);
            gstore_sym_auto<D>(io.Pfs + (b * T + t) * (D * D), P);
        }
        if (io.nell && !io.nell_last_only) nellw.put(t, T, acc);
    }
    if (io.nell && io.nell_last_only) io.nell[b] = acc;
}

// ================================================================================================ sigma-point pieces
// Sigma-point moments of the discretised model (filters_smoothers.py:88-121) computed cooperatively by G lanes:
//   mp = sum_i w_i f(chi_i),  Pp = sum_i w_i (f f^T + Sigma) - mp mp^T   (uncentred form, :120),
//   optionally Dx = sum_i w_i chi_i f_i^T - m mp^T (:525).            chi_i = m + L xi_i (quadratures.py:201).
// SHARE (Gauss-Hermite tables, dimension 0 fastest, quadratures.py:181-189): points i, i + nb, ..., i + (p-1) nb
// (nb = n / p) differ only in the LAST coordinate of xi, and L is lower triangular, so their chi[0..D-2] -- in
// particular chi[V] -- are bit-identical: the transcendental part of the model (softplus, sin, cos) is
// evaluated once per base index instead of once per point.  Results are unchanged.
template <class Model, int G, int P, bool CROSS>
CGP_DEV void sgp_moments(const Model &mdl, const double *__restrict__ sw, const double *__restrict__ sxi, int n,
                         int lane, const double (&m)[Model::D], const double (&Pc)[NSym<Model::D>::value],
                         double (&mp)[Model::D], double (&Pp)[NSym<Model::D>::value], double (&Dx)[Model::D][Model::D],
                         double *gsm = nullptr) {
    constexpr int D = Model::D, NS = NSym<D>::value;
    double L[NS];
    chol_lower_sym_rsqrt<D>(Pc, L);
    double am[D], aP[NS];
    CGP_UNROLL for (int i = 0; i < D; i++) am[i] = 0.;
    CGP_UNROLL for (int i = 0; i < NS; i++) aP[i] = 0.;
    if (CROSS) { CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) Dx[r][c] = 0.; }

    auto accumulate = [&](int i, const double (&chi)[D], const double (&ev)[D]) {
        const double w = __ldg(sw + i);
        CGP_UNROLL for (int r = 0; r < D; r++) am[r] = fma(w, ev[r], am[r]);
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++) {
            double v = ev[r] * ev[c];
            if (Model::has_sig(r, c)) v += mdl.sig(r, c);
            aP[sidx(r, c)] = fma(w, v, aP[sidx(r, c)]);
        }
        if (CROSS) {
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++)
                Dx[r][c] = fma(w, chi[r] * ev[c], Dx[r][c]);
        }
    };
    auto chi_row = [&](int i, int r) {
        double s = L[sidx(r, 0)] * __ldg(sxi + i * D);
        CGP_UNROLL for (int c = 1; c <= r; c++) s = fma(L[sidx(r, c)], __ldg(sxi + i * D + c), s);
        return m[r] + s;
    };
    if constexpr (P > 0) {
        // Gauss-Hermite table with P nodes per dimension: points base + c nb (c < P) share chi[0..D-2] and with it
        // the transcendental part and the chirp rows ev[0..V-1] of the model; only chi[D-1] and the Matern rows
        // ev[V], ev[V+1] differ.  The weighted sums over the P points are therefore formed from
        //   W = sum_c w_c,  S_t = sum_c w_c ev[V+t]_c,  X = sum_c w_c chi[D-1]_c
        // and the P x (3 or 2) genuinely quadratic terms: the same sums as the reference's einsum, associated
        // differently (differences at rounding level, far below the algorithm's own summation-order noise).
        constexpr int V = D - 2;
        const int nb = n / P;
        for (int base = lane; base < nb; base += G) {
            double chi[D], ev[D];
            CGP_UNROLL for (int r = 0; r < D; r++) chi[r] = chi_row(base, r);
            const typename Model::Trig trig = mdl.prep(chi);
            mdl.mean_with(trig, chi, ev);
            double W = 0., S0 = 0., S1 = 0., X = 0., q00 = 0., q10 = 0., q11 = 0., x0 = 0., x1 = 0.;
            CGP_UNROLL for (int c = 0; c < P; c++) {
                const int i = base + c * nb;
                if (c > 0) { chi[D - 1] = chi_row(i, D - 1); mdl.mean_tail(chi, ev); }
                const double w = __ldg(sw + i);
                W += w;
                S0 = fma(w, ev[V], S0);
                S1 = fma(w, ev[V + 1], S1);
                q00 = fma(w, ev[V] * ev[V] + mdl.sig(V, V), q00);
                q10 = fma(w, ev[V + 1] * ev[V] + mdl.sig(V + 1, V), q10);
                q11 = fma(w, ev[V + 1] * ev[V + 1] + mdl.sig(V + 1, V + 1), q11);
                if (CROSS) {
                    X = fma(w, chi[D - 1], X);
                    x0 = fma(w, chi[D - 1] * ev[V], x0);
                    x1 = fma(w, chi[D - 1] * ev[V + 1], x1);
                }
            }
            CGP_UNROLL for (int r = 0; r < V; r++) am[r] = fma(W, ev[r], am[r]);
            am[V] += S0; am[V + 1] += S1;
            CGP_UNROLL for (int r = 0; r < V; r++) CGP_UNROLL for (int c = 0; c <= r; c++) {
                double v = ev[r] * ev[c];
                if (Model::has_sig(r, c)) v += mdl.sig(r, c);
                aP[sidx(r, c)] = fma(W, v, aP[sidx(r, c)]);
            }
            CGP_UNROLL for (int c = 0; c < V; c++) {
                aP[sidx(V, c)] = fma(ev[c], S0, aP[sidx(V, c)]);
                aP[sidx(V + 1, c)] = fma(ev[c], S1, aP[sidx(V + 1, c)]);
            }
            aP[sidx(V, V)] += q00; aP[sidx(V + 1, V)] += q10; aP[sidx(V + 1, V + 1)] += q11;
            if (CROSS) {
                CGP_UNROLL for (int r = 0; r < D - 1; r++) {
                    CGP_UNROLL for (int c = 0; c < V; c++) Dx[r][c] = fma(W, chi[r] * ev[c], Dx[r][c]);
                    Dx[r][V] = fma(chi[r], S0, Dx[r][V]);
                    Dx[r][V + 1] = fma(chi[r], S1, Dx[r][V + 1]);
                }
                CGP_UNROLL for (int c = 0; c < V; c++) Dx[D - 1][c] = fma(ev[c], X, Dx[D - 1][c]);
                Dx[D - 1][V] += x0; Dx[D - 1][V + 1] += x1;
            }
        }
    } else {
        for (int i = lane; i < n; i += G) {
            double chi[D], ev[D];
            CGP_UNROLL for (int r = 0; r < D; r++) chi[r] = chi_row(i, r);
            mdl.mean(chi, ev);
            accumulate(i, chi, ev);
        }
    }
    if constexpr (G > 1 && !CROSS) {
        if (gsm) {                                     // many values: transpose through shared memory
            double a[D + NS];
            CGP_UNROLL for (int r = 0; r < D; r++) a[r] = am[r];
            CGP_UNROLL for (int i = 0; i < NS; i++) a[D + i] = aP[i];
            group_sum_smem<D + NS, G>(a, gsm, lane);
            CGP_UNROLL for (int r = 0; r < D; r++) mp[r] = a[r];
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++)
                Pp[sidx(r, c)] = a[D + sidx(r, c)] - mp[r] * mp[c];
            return;
        }
    }
    CGP_UNROLL for (int r = 0; r < D; r++) mp[r] = group_allreduce<G>(am[r]);
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++)
        Pp[sidx(r, c)] = group_allreduce<G>(aP[sidx(r, c)]) - mp[r] * mp[c];
    if (CROSS) {
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++)
            Dx[r][c] = group_allreduce<G>(Dx[r][c]) - m[r] * mp[c];
    }
}

// rhs of the CD sigma-point moment ODE (filters_smoothers.py:124-137), G lanes cooperating:
//   dm = sum_i w_i a(chi_i),  Q = sum_i w_i (chi_i - m) a(chi_i)^T,  dP = Q + Q^T + b b^T.
template <class Model, int G, int P>
CGP_DEV void cd_sgp_ode(const Model &mdl, const double *__restrict__ sw, const double *__restrict__ sxi, int n,
                        int lane, const double (&Qc)[NSym<Model::D>::value], const double (&m)[Model::D],
                        const double (&Pc)[NSym<Model::D>::value], double (&dm)[Model::D],
                        double (&dP)[NSym<Model::D>::value], double *gsm = nullptr) {
    constexpr int D = Model::D, NS = NSym<D>::value;
    double L[NS];
    chol_lower_sym_rsqrt<D>(Pc, L);
    double am[D], aQ[D][D];
    CGP_UNROLL for (int i = 0; i < D; i++) am[i] = 0.;
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) aQ[r][c] = 0.;
    auto make_chi = [&](int i, double (&chi)[D]) {
        CGP_UNROLL for (int r = 0; r < D; r++) {
            double s = L[sidx(r, 0)] * __ldg(sxi + i * D);
            CGP_UNROLL for (int c = 1; c <= r; c++) s = fma(L[sidx(r, c)], __ldg(sxi + i * D + c), s);
            chi[r] = m[r] + s;
        }
    };
    auto accumulate = [&](int i, const double (&chi)[D], const double (&f)[D]) {
        const double w = __ldg(sw + i);
        CGP_UNROLL for (int r = 0; r < D; r++) am[r] = fma(w, f[r], am[r]);
        CGP_UNROLL for (int r = 0; r < D; r++) {
            const double dr = chi[r] - m[r];
            CGP_UNROLL for (int c = 0; c < D; c++) aQ[r][c] = fma(w, dr * f[c], aQ[r][c]);
        }
    };
    if constexpr (P > 0 && !Model::kLinear) {
        const int nb = n / P;
        for (int base = lane; base < nb; base += G) {
            double chi[D], f[D];
            make_chi(base, chi);
            const double w = mdl.omega(chi[Model::V]);
            mdl.drift_w(w, chi, f);
            accumulate(base, chi, f);
            CGP_UNROLL for (int c = 1; c < P; c++) {
                const int i = base + c * nb;
                make_chi(i, chi);
                mdl.drift_w(w, chi, f);
                accumulate(i, chi, f);
            }
        }
    } else {
        for (int i = lane; i < n; i += G) {
            double chi[D], f[D];
            make_chi(i, chi);
            mdl.drift(chi, f);
            accumulate(i, chi, f);
        }
    }
    bool summed = false;
    if constexpr (G > 1) {
        if (gsm) {
            double a[D + D * D];
            CGP_UNROLL for (int r = 0; r < D; r++) a[r] = am[r];
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) a[D + r * D + c] = aQ[r][c];
            group_sum_smem<D + D * D, G>(a, gsm, lane);
            CGP_UNROLL for (int r = 0; r < D; r++) dm[r] = a[r];
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) aQ[r][c] = a[D + r * D + c];
            summed = true;
        }
    }
    if (!summed) {
        CGP_UNROLL for (int r = 0; r < D; r++) dm[r] = group_allreduce<G>(am[r]);
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) aQ[r][c] = group_allreduce<G>(aQ[r][c]);
    }
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++)
        dP[sidx(r, c)] = (aQ[r][c] + aQ[c][r]) + Qc[sidx(r, c)];
}

// ================================================================================================ group-per-chirp filters
// Launch shape of the group-per-chirp kernels: the shared-memory reduction is used when a lane holds >= 24 partial
// sums (d >= 6); the block shrinks to 64 threads when 128 would need more than 40 KB of static shared memory.
template <class Model, int G, bool CD> struct GroupCfg {
    static constexpr int D = Model::D;
    static constexpr int NA = CD ? D + D * D : D + NSym<D>::value;
    static constexpr bool kSmem = G > 1 && NA >= 24;
    static constexpr int kPerGroup = GroupSmem<NA, G>::kDoubles;
    static constexpr bool kFits128 = !kSmem || (128 / G) * kPerGroup * 8 <= 40960;
    static constexpr bool kFits64 = !kSmem || (64 / G) * kPerGroup * 8 <= 40960;
    static constexpr int kBlock = kFits128 ? 128 : ((kFits64 || G > 32) ? 64 : 32);        // d = 12: one 32-lane block
    static constexpr int kGroups = kBlock / G;
};

// sgp_filter (filters_smoothers.py:446-490) and cd_sgp_filter (:534-582)
template <class Model, int G, int P, bool CD>
__global__ void __launch_bounds__(GroupCfg<Model, G, CD>::kBlock) sgp_filter_kernel(const CgpProblem p, const FilterIO io) {
    constexpr int D = Model::D, NS = NSym<D>::value;
    using Cfg = GroupCfg<Model, G, CD>;
    __shared__ double gsm_all[Cfg::kSmem ? Cfg::kGroups * Cfg::kPerGroup : 1];
    __shared__ double nl_all[G > 1 ? Cfg::kBlock : 1];            // nll increments of G consecutive steps, per group
    double *gsm = Cfg::kSmem ? gsm_all + (threadIdx.x / G) * Cfg::kPerGroup : nullptr;
    double *nl = nl_all + (G > 1 ? (threadIdx.x / G) * G : 0);
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    const bool active = gid < p.B;
    const int64_t b = active ? gid : p.B - 1;        // idle groups shadow the last chirp (shuffles need all lanes)
    Model mdl;
    if constexpr (CD) mdl.load(p.consts + b * p.consts_stride);
    else mdl.load(p.consts + b * p.consts_stride, p.dt);
    double m[D], Pc[NS], H[D], Qc[NS];
    load_vec<D>(p.m0 + b * p.m0_stride, m);
    load_sym<D>(p.P0 + b * p.P0_stride, Pc);
    if constexpr (CD) load_sym<D>(p.Qc + b * p.Qc_stride, Qc);
    CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
    const double *__restrict__ y = io.ys + (b / p.ys_repeat) * p.T;
    const int64_t T = p.T;
    const bool store = io.mfs != nullptr && active && lane == 0;
    const bool store_nell = io.nell != nullptr && active && lane == 0;
    const int n = p.n_sigma;
    const double *__restrict__ sw = p.sig_w;
    const double *__restrict__ sxi = p.sig_xi;
    const double dt = p.dt;
    double acc = 0., Sk = 1., rk = 0.;
    double ynext = __ldg(y);
    for (int64_t t = 0; t < T; t++) {
        const double yt = ynext;
        if (t + 1 < T) ynext = __ldg(y + t + 1);
        double mp[D], Pp[NS];
        if constexpr (CD) {
            CGP_UNROLL for (int i = 0; i < D; i++) mp[i] = m[i];
            CGP_UNROLL for (int i = 0; i < NS; i++) Pp[i] = Pc[i];
            rk4_step<D>([&](const double (&mm)[D], const double (&PP)[NS], double (&dm)[D], double (&dP)[NS]) {
                cd_sgp_ode<Model, G, P>(mdl, sw, sxi, n, lane, Qc, mm, PP, dm, dP, gsm);
            }, mp, Pp, dt);
        } else {
            double dummy[D][D];
            sgp_moments<Model, G, P, false>(mdl, sw, sxi, n, lane, m, Pc, mp, Pp, dummy, gsm);
        }
        double S, resid;
        linear_update_fast<D, false>(mp, Pp, H, p.Xi, yt, m, Pc, S, resid);
        if (store) {
            gstore_vec_auto<D>(io.mfs + (b * T + t) * D, m);
            gstore_sym_auto<D>(io.Pfs + (b * T + t) * (D * D), Pc);
        }
        if constexpr (G == 1) {
            acc = acc + nll_increment(S, resid);
            if (store_nell && !io.nell_last_only) io.nell[b * T + t] = acc;
        } else {
            // lane (t mod G) keeps (S, r) of step t; every G steps the increments are evaluated in SIMD (one log / sqrt /
            // divide per lane instead of one per step) and accumulated in the reference's sequential order
            const int slot = (int)(t % G);
            if (lane == slot) { Sk = S; rk = resid; }
            if (slot == G - 1 || t == T - 1) {
                const int cnt = slot + 1;
                nl[lane] = lane < cnt ? nll_increment(Sk, rk) : 0.;
                __syncwarp();
                if (lane == 0) {
                    double c = acc;
                    for (int j = 0; j < cnt; j++) { c = c + nl[j]; nl[j] = c; }
                }
                __syncwarp();
                acc = nl[cnt - 1];
                if (io.nell != nullptr && active && !io.nell_last_only && lane < cnt) io.nell[b * T + (t - slot) + lane] = nl[lane];
                __syncwarp();
            }
        }
    }
    if (store_nell && io.nell_last_only) io.nell[b] = acc;
}

// ================================================================================================ discrete smoothers
// One lane writes R doubles (R even, dst 16-byte aligned) in whole 32-byte sectors wherever dst allows: a record that starts
// in the middle of a sector goes out as 16 bytes + 256-bit stores + the rest (records of 30 doubles alternate between the two).
template <int R> CGP_DEV void gstore_flat(double *__restrict__ dst, const double (&v)[R]) {
    static_assert(R % 2 == 0, "even record");
    auto st16 = [](double *q, double a, double b) { *reinterpret_cast<double2 *>(q) = make_double2(a, b); };
    if ((reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
        CGP_UNROLL for (int i = 0; i + 4 <= R; i += 4) stg256(dst + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
        if constexpr (R % 4 != 0) st16(dst + R - 2, v[R - 2], v[R - 1]);
    } else {
        st16(dst, v[0], v[1]);
        CGP_UNROLL for (int i = 2; i + 4 <= R; i += 4) stg256(dst + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
        if constexpr ((R - 2) % 4 != 0) st16(dst + R - 2, v[R - 2], v[R - 1]);
    }
}

// Shared tail of the gain kernels: G = (cho_solve(chol(Pp), DT))^T (filters_smoothers.py:81-82), then the workspace
// record [G | c | C] of this (chirp, step): c = mf - G mp, C = Pf - G D^T (lower triangle of Pf, as cho_factor would read).
template <int D>
CGP_DEV void gain_and_store(const double (&DT)[D][D], const double (&mp)[D], const double (&Pp)[D][D], const double (&mf)[D],
                            const double (&Pfl)[NSym<D>::value], double *__restrict__ rec) {
    constexpr int NS = NSym<D>::value, R = ws_record<D>();
    double L[D][D], rinv[D], Gm[D][D];
    chol_lower_rsqrt<D>(Pp, L, rinv);
    CGP_UNROLL for (int c = 0; c < D; c++) {          // column c of X = Pp^{-1} DT is row c of G^T ... G = X^T
        double col[D];
        CGP_UNROLL for (int i = 0; i < D; i++) col[i] = DT[i][c];
        chol_solve_vec_rinv<D>(L, rinv, col);
        CGP_UNROLL for (int i = 0; i < D; i++) Gm[c][i] = col[i];
    }
    double cv[D], Cs[NS];
    CGP_UNROLL for (int r = 0; r < D; r++) {
        double sacc = mf[r];
        CGP_UNROLL for (int k = 0; k < D; k++) sacc = fma(-Gm[r][k], mp[k], sacc);
        cv[r] = sacc;
    }
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int q = 0; q <= r; q++) {
        double sacc = Pfl[sidx(r, q)];
        CGP_UNROLL for (int k = 0; k < D; k++) sacc = fma(-Gm[r][k], DT[k][q], sacc);      // (G D^T)_rq, D_qk = DT_kq
        Cs[sidx(r, q)] = sacc;
    }
    if constexpr (D <= 6) {
        if ((reinterpret_cast<uintptr_t>(rec) & 15u) == 0) {
            double v[R];
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int k = 0; k < D; k++) v[r * D + k] = Gm[r][k];
            CGP_UNROLL for (int r = 0; r < D; r++) v[D * D + r] = cv[r];
            CGP_UNROLL for (int i = 0; i < NS; i++) v[D * D + D + i] = Cs[i];
            CGP_UNROLL for (int i = D * D + D + NS; i < R; i++) v[i] = 0.;
            gstore_flat<R>(rec, v);
            return;
        }
    }
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int k = 0; k < D; k++) rec[r * D + k] = Gm[r][k];
    CGP_UNROLL for (int r = 0; r < D; r++) rec[D * D + r] = cv[r];
    CGP_UNROLL for (int i = 0; i < NS; i++) rec[D * D + D + i] = Cs[i];
}

// rts (filters_smoothers.py:187-219) / eks (:317-349) gains: one thread per (chirp, step k), k in [0, T-2].
template <class Model>
__global__ void __launch_bounds__(128) eks_gain_kernel(const CgpProblem p, const SmootherIO io) {
    constexpr int D = Model::D;
    const int64_t Tm1 = p.T - 1;
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= p.B * Tm1) return;
    const int64_t b = item / Tm1, t = item - b * Tm1;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    double mf[D], Pf[D][D];
    load_vec<D>(io.mfs + (b * p.T + t) * D, mf);
    load_mat<D>(io.Pfs + (b * p.T + t) * (D * D), Pf);
    double mp[D], J[D][D], DT[D][D], Pp[D][D];
    mdl.mean_jac(mf, mp, J);                          // :342-343
    jmul<Model, D>(J, Pf, DT);                        // DT = J Pf (:345)
    mul_jt<Model, D>(DT, J, Pp);                      // :344
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++)
        if (Model::has_sig(r, c)) Pp[r][c] += mdl.sig(r, c);
    double Pfl[NSym<D>::value];
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++) Pfl[sidx(r, c)] = Pf[r][c];
    gain_and_store<D>(DT, mp, Pp, mf, Pfl, io.ws + (b * p.T + t) * ws_record<D>());
}

// rts / eks in ONE pass, one thread per chirp, no workspace: for k = T-2 .. 0 the gain is formed from (mf_k, Pf_k)
// (:342-345, :81-82) and applied at once (:83-84).  For batches large enough to fill the GPU with one thread per chirp (the
// CRLB job runs 10^6 chirps, tetralith/jobs/crlb_ekf.py:59): DRAM traffic is the algorithmic 2 * 8 (d + d^2) bytes per step
// instead of ~3.3x that through the [G | mp | Pp] workspace of the two-kernel path.  (mf, Pf) of the next step are
// fetched while the current one is evaluated.
template <class Model, bool WIDE = false>
__global__ void __launch_bounds__(128) eks_onepass_thread_kernel(const CgpProblem p, const SmootherIO io) {
    constexpr int D = Model::D;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const int64_t T = p.T;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    double ms[D], Ps[D][D];
    gload_vec<WIDE, D>(io.mfs + (b * T + T - 1) * D, ms);
    gload_mat<WIDE, D>(io.Pfs + (b * T + T - 1) * (D * D), Ps);
    gstore_vec<WIDE, D>(io.mss + (b * T + T - 1) * D, ms);
    gstore_mat<WIDE, D>(io.Pss + (b * T + T - 1) * (D * D), Ps);
    if (T < 2) return;
    double mfn[D], Pfn[D][D];
    gload_vec<WIDE, D>(io.mfs + (b * T + T - 2) * D, mfn);
    gload_mat<WIDE, D>(io.Pfs + (b * T + T - 2) * (D * D), Pfn);
    for (int64_t t = T - 2; t >= 0; t--) {
        double mf[D], Pf[D][D];
        CGP_UNROLL for (int r = 0; r < D; r++) {
            mf[r] = mfn[r];
            CGP_UNROLL for (int c = 0; c < D; c++) Pf[r][c] = Pfn[r][c];
        }
        if (t > 0) {
            gload_vec<WIDE, D>(io.mfs + (b * T + t - 1) * D, mfn);
            gload_mat<WIDE, D>(io.Pfs + (b * T + t - 1) * (D * D), Pfn);
        }
        double mp[D], J[D][D], DT[D][D], Pp[D][D];
        mdl.mean_jac(mf, mp, J);                          // :342-343
        jmul<Model, D>(J, Pf, DT);                        // DT = J Pf (:345)
        mul_jt<Model, D>(DT, J, Pp);                      // :344
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++)
            if (Model::has_sig(r, c)) Pp[r][c] += mdl.sig(r, c);
        double L[D][D], rinv[D], Gm[D][D];
        chol_lower_rsqrt<D>(Pp, L, rinv);
        CGP_UNROLL for (int c = 0; c < D; c++) {          // G = (Pp^{-1} DT)^T, as gain_and_store
            double col[D];
            CGP_UNROLL for (int i = 0; i < D; i++) col[i] = DT[i][c];
            chol_solve_vec_rinv<D>(L, rinv, col);
            CGP_UNROLL for (int i = 0; i < D; i++) Gm[c][i] = col[i];
        }
        double dm[D], dP[D][D], t1[D][D], t2[D][D], gm[D];
        CGP_UNROLL for (int r = 0; r < D; r++) dm[r] = ms[r] - mp[r];
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) dP[r][c] = Ps[r][c] - Pp[r][c];
        matvec<D>(Gm, dm, gm);
        CGP_UNROLL for (int r = 0; r < D; r++) ms[r] = mf[r] + gm[r];
        matmul<D>(Gm, dP, t1);
        matmul_nt<D>(t1, Gm, t2);
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) Ps[r][c] = Pf[r][c] + t2[r][c];
        gstore_vec<WIDE, D>(io.mss + (b * T + t) * D, ms);
        gstore_mat<WIDE, D>(io.Pss + (b * T + t) * (D * D), Ps);
    }
}

// sgp_smoother gains (filters_smoothers.py:520-527): one thread per (chirp, step), all sigma points serially.
template <class Model, int P>
__global__ void __launch_bounds__(128) sgp_gain_kernel(const CgpProblem p, const SmootherIO io) {
    constexpr int D = Model::D, NS = NSym<D>::value;
    const int64_t Tm1 = p.T - 1;
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= p.B * Tm1) return;
    const int64_t b = item / Tm1, t = item - b * Tm1;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    double mf[D], Pf[NS];
    load_vec<D>(io.mfs + (b * p.T + t) * D, mf);
    load_sym<D>(io.Pfs + (b * p.T + t) * (D * D), Pf);
    double mp[D], Pps[NS], Dx[D][D];
    sgp_moments<Model, 1, P, true>(mdl, p.sig_w, p.sig_xi, p.n_sigma, 0, mf, Pf, mp, Pps, Dx);
    double DT[D][D], Pp[D][D];
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) DT[r][c] = Dx[c][r];
    sym_to_full<D>(Pps, Pp);
    gain_and_store<D>(DT, mp, Pp, mf, Pf, io.ws + (b * p.T + t) * ws_record<D>());
}

// Sequential sweep (filters_smoothers.py:83-84 regrouped, scan :218/:348/:530, stacking :140-142): one thread per chirp.
//   ms = c + G ms';   Ps = C + G Ps' G^T
// The record of the next step is fetched into registers while the current step is evaluated.
template <int D>
__global__ void __launch_bounds__(64) smoother_sweep_kernel(const CgpProblem p, const SmootherIO io) {
    constexpr int R = ws_record<D>(), NS = NSym<D>::value;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const int64_t T = p.T;
    double ms[D], Ps[D][D];
    load_vec<D>(io.mfs + (b * T + T - 1) * D, ms);
    load_mat<D>(io.Pfs + (b * T + T - 1) * (D * D), Ps);
    gstore_vec_auto<D>(io.mss + (b * T + T - 1) * D, ms);
    gstore_mat_auto<D>(io.Pss + (b * T + T - 1) * (D * D), Ps);
    if (T < 2) return;
    double Gn[D][D], cn[D], Cn[NS];
    auto fetch = [&](int64_t t) {
        const double *rec = io.ws + (b * T + t) * R;
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) Gn[r][c] = rec[r * D + c];
        CGP_UNROLL for (int i = 0; i < D; i++) cn[i] = rec[D * D + i];
        CGP_UNROLL for (int i = 0; i < NS; i++) Cn[i] = rec[D * D + D + i];
    };
    fetch(T - 2);
    for (int64_t t = T - 2; t >= 0; t--) {
        double Gm[D][D], cv[D], Cs[NS];
        CGP_UNROLL for (int r = 0; r < D; r++) {
            cv[r] = cn[r];
            CGP_UNROLL for (int c = 0; c < D; c++) Gm[r][c] = Gn[r][c];
        }
        CGP_UNROLL for (int i = 0; i < NS; i++) Cs[i] = Cn[i];
        if (t > 0) fetch(t - 1);
        double t1[D][D], t2[D][D], gm[D];
        matvec<D>(Gm, ms, gm);
        CGP_UNROLL for (int r = 0; r < D; r++) ms[r] = cv[r] + gm[r];
        matmul<D>(Gm, Ps, t1);
        matmul_nt<D>(t1, Gm, t2);
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) Ps[r][c] = Cs[sidx(r, c)] + t2[r][c];
        gstore_vec_auto<D>(io.mss + (b * T + t) * D, ms);
        gstore_mat_auto<D>(io.Pss + (b * T + t) * (D * D), Ps);
    }
}

// ================================================================================================ CD smoothers
// cd_eks (filters_smoothers.py:400-443): thread per chirp, RK4 backwards with dt <- -dt (:423).
//   rhs (:427-432): gamma = b b^T;  M = J_a(m) + (Pf^{-1} gamma^T)^T;  dm = a(m) + gamma Pf^{-1} (m - mf);
//                   dP = M P + P M^T - gamma.            chol(Pf) and Pf^{-1} gamma are hoisted out of the 4 stages.
template <class Model>
__global__ void __launch_bounds__(64) cd_eks_thread_kernel(const CgpProblem p, const SmootherIO io) {
    constexpr int D = Model::D, NS = NSym<D>::value;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const int64_t T = p.T;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride);
    double Qc[NS], Qf[D][D];
    load_sym<D>(p.Qc + b * p.Qc_stride, Qc);
    sym_to_full<D>(Qc, Qf);
    double ms[D], Ps[NS];
    load_vec<D>(io.mfs + (b * T + T - 1) * D, ms);
    load_sym<D>(io.Pfs + (b * T + T - 1) * (D * D), Ps);
    gstore_vec_auto<D>(io.mss + (b * T + T - 1) * D, ms);
    gstore_sym_auto<D>(io.Pss + (b * T + T - 1) * (D * D), Ps);
    const double ndt = -p.dt;
    for (int64_t t = T - 2; t >= 0; t--) {
        double mf[D], Pf[D][D], Lf[D][D], X[D][D];
        load_vec<D>(io.mfs + (b * T + t) * D, mf);
        load_mat<D>(io.Pfs + (b * T + t) * (D * D), Pf);
        chol_lower<D>(Pf, Lf);
        chol_solve_mat<D>(Lf, Qf, X);                 // Pf^{-1} gamma^T (gamma symmetric)
        rk4_step<D>([&](const double (&mm)[D], const double (&PP)[NS], double (&dm)[D], double (&dP)[NS]) {
            double J[D][D], a[D], z[D], Y[D][D];
            mdl.drift_jac(mm, a, J);
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) J[r][c] = J[r][c] + X[c][r];
            CGP_UNROLL for (int i = 0; i < D; i++) z[i] = mm[i] - mf[i];
            chol_solve_vec<D>(Lf, z);
            CGP_UNROLL for (int r = 0; r < D; r++) {
                double s = Qf[r][0] * z[0];
                CGP_UNROLL for (int k = 1; k < D; k++) s = fma(Qf[r][k], z[k], s);
                dm[r] = a[r] + s;
            }
            CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j < D; j++) {
                double s = J[i][0] * PP[sidx(0, j)];
                CGP_UNROLL for (int k = 1; k < D; k++) s = fma(J[i][k], PP[sidx(k, j)], s);
                Y[i][j] = s;
            }
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++)
                dP[sidx(r, c)] = (Y[r][c] + Y[c][r]) - Qc[sidx(r, c)];
        }, ms, Ps, ndt);
        gstore_vec_auto<D>(io.mss + (b * T + t) * D, ms);
        gstore_sym_auto<D>(io.Pss + (b * T + t) * (D * D), Ps);
    }
}

// cd_sgp_smoother (filters_smoothers.py:585-632): group of G lanes per chirp.
//   rhs (:615-621): Gm = Pf^{-1} gamma;  (_m, _P) = cd_sgp_common(m, P);
//                   dm = _m + Gm^T (m - mf);   dP = _P + Gm^T P + P Gm - 2 gamma.
template <class Model, int G, int P>
__global__ void __launch_bounds__(GroupCfg<Model, G, true>::kBlock) cd_sgp_smoother_kernel(const CgpProblem p, const SmootherIO io) {
    constexpr int D = Model::D, NS = NSym<D>::value;
    using Cfg = GroupCfg<Model, G, true>;
    __shared__ double gsm_all[Cfg::kSmem ? Cfg::kGroups * Cfg::kPerGroup : 1];
    double *gsm = Cfg::kSmem ? gsm_all + (threadIdx.x / G) * Cfg::kPerGroup : nullptr;
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    const bool active = gid < p.B;
    const int64_t b = active ? gid : p.B - 1;
    const bool store = active && lane == 0;
    const int64_t T = p.T;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride);
    double Qc[NS], Qf[D][D];
    load_sym<D>(p.Qc + b * p.Qc_stride, Qc);
    sym_to_full<D>(Qc, Qf);
    double ms[D], Ps[NS];
    load_vec<D>(io.mfs + (b * T + T - 1) * D, ms);
    load_sym<D>(io.Pfs + (b * T + T - 1) * (D * D), Ps);
    if (store) {
        gstore_vec_auto<D>(io.mss + (b * T + T - 1) * D, ms);
        gstore_sym_auto<D>(io.Pss + (b * T + T - 1) * (D * D), Ps);
    }
    const int n = p.n_sigma;
    const double ndt = -p.dt;
    for (int64_t t = T - 2; t >= 0; t--) {
        double mf[D], Pf[D][D], Lf[D][D], Gm[D][D];
        load_vec<D>(io.mfs + (b * T + t) * D, mf);
        load_mat<D>(io.Pfs + (b * T + t) * (D * D), Pf);
        chol_lower<D>(Pf, Lf);
        chol_solve_mat<D>(Lf, Qf, Gm);                // Gm = Pf^{-1} gamma
        rk4_step<D>([&](const double (&mm)[D], const double (&PP)[NS], double (&dm)[D], double (&dP)[NS]) {
            double _m[D], _P[NS], W[D][D];
            cd_sgp_ode<Model, G, P>(mdl, p.sig_w, p.sig_xi, n, lane, Qc, mm, PP, _m, _P, gsm);
            CGP_UNROLL for (int r = 0; r < D; r++) {
                double s = Gm[0][r] * (mm[0] - mf[0]);
                CGP_UNROLL for (int k = 1; k < D; k++) s = fma(Gm[k][r], mm[k] - mf[k], s);
                dm[r] = _m[r] + s;
            }
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) {      // W = Gm^T P
                double s = Gm[0][r] * PP[sidx(0, c)];
                CGP_UNROLL for (int k = 1; k < D; k++) s = fma(Gm[k][r], PP[sidx(k, c)], s);
                W[r][c] = s;
            }
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++)
                dP[sidx(r, c)] = ((_P[sidx(r, c)] + W[r][c]) + W[c][r]) - 2 * Qc[sidx(r, c)];
        }, ms, Ps, ndt);
        if (store) {
            gstore_vec_auto<D>(io.mss + (b * T + t) * D, ms);
            gstore_sym_auto<D>(io.Pss + (b * T + t) * (D * D), Ps);
        }
    }
}

}  // namespace cgp
