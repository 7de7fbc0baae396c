// cgp_fast.cuh -- tuned kernels for the headline path (Gauss-Hermite sigma-point filter + smoother on the
// chirp-family LCD models).  Same arithmetic as the generic kernels in cgp_kernels.cuh up to rounding:
//   * Cholesky columns are scaled by rsqrt(pivot) instead of divided by sqrt(pivot)      (<= 2 ulp per entry)
//   * the Kalman gain uses one reciprocal of S instead of d divisions                     (<= 1 ulp per entry)
//   * sigma-point partial sums are combined through shared memory in a fixed tree order instead of a
//     shuffle butterfly (a different, still deterministic, summation order).
#pragma once
#include "cgp_kernels.cuh"

namespace cgp {

// ------------------------------------------------------------------------------------------------ warp-per-chirp sigma-point filters
// All 32 lanes' partial sums a[0..NA) are combined through shared memory in a fixed tree order; every lane gets
// the same totals.  red: [NA][kRedPitch] doubles (16-byte aligned rows: the partials are read back with 16-byte loads, and
// pitch 34 keeps the 8 rows of a quarter-warp on different bank groups), res: [NA rounded up to even] doubles (16-byte aligned).
constexpr int kRedPitch = 34;
// rows of `red`: NA rounded up to the number of lanes that read (16 or 32) -- the lanes without a sum read a row of their own
// (never written, never used) instead of row 0, which would put two different addresses on the bank group of lanes 0 / 8.
template <int NA> struct RedRows { static constexpr int value = NA <= 16 ? 16 : ((NA + 31) / 32) * 32; };
template <int NA>
CGP_DEV void warp_sum_smem(const double (&a)[NA], double (*red)[kRedPitch], double *res, int lane, double (&tot)[NA]) {
    constexpr int KP = (NA <= 16) ? 16 : 32;          // lanes per "half"
    constexpr int HS = 32 / KP;                       // halves: each sums 32 / HS partials
    constexpr int CNT = 32 / HS;
    CGP_UNROLL for (int k = 0; k < NA; k++) red[k][lane] = a[k];
    __syncwarp();
    const int k = lane % KP, h = lane / KP;
    CGP_UNROLL for (int k0 = 0; k0 < NA; k0 += KP) {
        const int kk = k0 + k;
        const bool ok = kk < NA;
        double v[CNT];
        CGP_UNROLL for (int j = 0; j < CNT; j += 2) {
            const double2 x = *reinterpret_cast<const double2 *>(&red[kk][h * CNT + j]);
            v[j] = x.x; v[j + 1] = x.y;
        }
        CGP_UNROLL for (int w2 = 1; w2 < CNT; w2 <<= 1)
            CGP_UNROLL for (int j = 0; j + w2 < CNT; j += 2 * w2) v[j] += v[j + w2];
        double sacc = v[0];
        if (HS == 2) sacc += __shfl_xor_sync(0xffffffffu, sacc, 16);
        if (ok && h == 0) res[kk] = sacc;
    }
    __syncwarp();
    CGP_UNROLL for (int k2 = 0; k2 < NA; k2 += 2) {
        if (k2 + 1 < NA) {
            const double2 v = *reinterpret_cast<const double2 *>(&res[k2]);
            tot[k2] = v.x; tot[k2 + 1] = v.y;
        } else {
            tot[k2] = res[k2];
        }
    }
}

// NB <= 8 partial sums e[] per lane -> totals at dst[0..NB) (shared memory).  red2 is [NB][36]: FOUR lanes per sum, lane
// 4 k + h adds the partials h, h + 4, ..., h + 28 of sum k in a fixed tree and two shuffle steps join the four quarters
// (pitch 36 = 4 mod 16 keeps the 16 lanes of a half-warp on 16 different banks).
// The caller has stored red2[k][lane] = e[k] and synchronised the warp.
constexpr int kSmallSumPitch = 36;
template <int NB> CGP_DEV void small_sums_tail(double (*red2)[kSmallSumPitch], double *dst, int lane) {
    static_assert(NB <= 8, "four lanes per sum");
    const int k = lane >> 2, h = lane & 3;
    const bool ok = k < NB;
    double v[8];
    CGP_UNROLL for (int j = 0; j < 8; j++) v[j] = red2[ok ? k : 0][h + 4 * j];
    CGP_UNROLL for (int w2 = 1; w2 < 8; w2 <<= 1)
        CGP_UNROLL for (int j = 0; j + w2 < 8; j += 2 * w2) v[j] += v[j + w2];
    double sacc = v[0];
    sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
    sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
    if (ok && h == 0) dst[k] = sacc;
}

// Per-lane slice of a Gauss-Hermite table with P nodes per dimension (quadratures.py:157-196, dimension 0 fastest):
// lane l owns base index l < nb = P^(D-1); its P points l + c nb share xi[0..D-2] and differ in the last coordinate.
template <int D, int P> struct GhLane {
    double xb[D - 1], wl[P], xlast[P], Wl;
    CGP_DEV void load(const CgpProblem &p, int lane) {
        int nb = 1;
        CGP_UNROLL for (int i = 0; i < D - 1; i++) nb *= P;
        const bool has = lane < nb;
        CGP_UNROLL for (int r = 0; r < D - 1; r++) xb[r] = has ? p.sig_xi[lane * D + r] : 0.;
        Wl = 0.;
        CGP_UNROLL for (int c = 0; c < P; c++) {
            wl[c] = has ? p.sig_w[lane + c * nb] : 0.;
            xlast[c] = p.sig_xi[(c * nb) * D + (D - 1)];
            Wl += wl[c];
        }
    }
    // chi[0..D-2] of the lane's base index and the partial dot product of the last row
    CGP_DEV void points(const double (&m)[D], const double (&L)[NSym<D>::value], double (&chi)[D], double &slast) const {
        CGP_UNROLL for (int r = 0; r < D - 1; r++) {
            double s = L[sidx(r, 0)] * xb[0];
            CGP_UNROLL for (int c = 1; c <= r; c++) s = fma(L[sidx(r, c)], xb[c], s);
            chi[r] = m[r] + s;
        }
        slast = L[sidx(D - 1, 0)] * xb[0];
        CGP_UNROLL for (int c = 1; c < D - 1; c++) slast = fma(L[sidx(D - 1, c)], xb[c], slast);
    }
};

// Discrete-time prediction (filters_smoothers.py:88-121) for ModelLCD<NH>: one warp, lane = base index.
template <int NH, int P, int DBG = 0> struct GhPredictLCD {
    using Model = ModelLCD<NH>;
    static constexpr int D = Model::D, V = Model::V, NS = NSym<D>::value, NA = D + NS;
    Model mdl;
    GhLane<D, P> tab;
    CGP_DEV void load(const CgpProblem &p, int64_t b, int lane) {
        mdl.load(p.consts + b * p.consts_stride, p.dt);
        tab.load(p, lane);
    }
    // Cross-covariance of the smoother (filters_smoothers.py:525), D = sum_i w_i chi_i mu_i^T - m mp^T = L E with
    // E = sum_i w_i xi_i mu_i^T (chi_i = m + L xi_i).  Which of the D*D entries of E need the quadrature:
    //   * the Matern rows mu[V], mu[V+1] are LINEAR in chi, so their columns of D are F P exactly in exact arithmetic (the
    //     rule integrates polynomials of degree <= 2 exactly: sum w xi = 0, sum w xi xi^T = I); they are formed from the
    //     filtering covariance when the gain is evaluated (gain_record);
    //   * the chirp rows mu[0..V-1] do not depend on chi[D-1], hence not on xi[D-1]: E[D-1][0..V-1] = 0.
    // Left: NE = (D-1) V sums, E[c][q] = sum_l xb_l[c] W_l ev_l[q] over the base indices l (the reference's 81-term sums up
    // to rounding: ~1e-16 relative, six orders below the algorithm's summation-order noise, tests/test_noise_floor.py).
    static constexpr int NE = (D - 1) * V;
    static CGP_DEV void cross_partials(const GhLane<D, P> &tab, const double (&ev)[V], double (&ec)[NE]) {
        CGP_UNROLL for (int c = 0; c < D - 1; c++)
            CGP_UNROLL for (int q = 0; q < V; q++) ec[c * V + q] = tab.xb[c] * (tab.Wl * ev[q]);
    }
    CGP_DEV void predict(double (*red)[kRedPitch], double *res, int lane, const double (&m)[D], const double (&Pc)[NS], double (&mp)[D],
                         double (&Pp)[NS]) const {
        predict_impl<false>(red, res, nullptr, lane, m, Pc, mp, Pp);
    }
    // EXPORT: the cross sums are formed by another warp (cgp_duo.cuh); this lane leaves ev[0..V-1] at xop[0..V-1][lane].
    template <bool EXPORT>
    CGP_DEV void predict_impl(double (*red)[kRedPitch], double *res, double (*xop)[33], int lane, const double (&m)[D],
                              const double (&Pc)[NS], double (&mp)[D], double (&Pp)[NS]) const {
        // The last pivot L[D-1][D-1] is needed by chi[D-1] only, i.e. after the transcendental chain that starts from chi[V]:
        // its rsqrt is issued behind the softplus branch, into the latency shadows of sincos (same values, ~60 cycles off the chain).
        double L[NS], piv;
        if constexpr (DBG == 3) { CGP_UNROLL for (int i = 0; i < NS; i++) L[i] = Pc[i]; piv = 1.; }
        else chol_lower_sym_rsqrt_head<D>(Pc, L, piv);
        double chi[D], slast;
        tab.points(m, L, chi, slast);
        typename Model::Trig trig;
        if constexpr (DBG == 2) { CGP_UNROLL for (int k = 0; k < NH; k++) { trig.c[k] = 0.99 + 1e-3 * chi[V]; trig.s[k] = 0.05; } }
        else trig = mdl.template prep_v<true>(chi[V]);
        if constexpr (DBG != 3) L[sidx(D - 1, D - 1)] = piv * fast_rsqrt(piv);
        // weighted sums over this lane's P points: only chi[D-1] and the Matern rows ev[V], ev[V+1] differ between
        // them, so the sums factor through W = sum w_c, S_t = sum w_c ev[V+t]_c and three quadratic terms
        double a[NA];
        {
            double ev[D], S0 = 0., S1 = 0., q00 = 0., q10 = 0., q11 = 0.;
            CGP_UNROLL for (int c = 0; c < P; c++) {
                chi[D - 1] = m[D - 1] + fma(L[sidx(D - 1, D - 1)], tab.xlast[c], slast);
                if (c == 0) mdl.mean_with(trig, chi, ev); else mdl.mean_tail(chi, ev);
                const double w = tab.wl[c];
                S0 = fma(w, ev[V], S0);
                S1 = fma(w, ev[V + 1], S1);
                q00 = fma(w, ev[V] * ev[V] + mdl.sig(V, V), q00);
                q10 = fma(w, ev[V + 1] * ev[V] + mdl.sig(V + 1, V), q10);
                q11 = fma(w, ev[V + 1] * ev[V + 1] + mdl.sig(V + 1, V + 1), q11);
            }
            CGP_UNROLL for (int r = 0; r < V; r++) a[r] = tab.Wl * ev[r];
            a[V] = S0; a[V + 1] = S1;
            CGP_UNROLL for (int r = 0; r < V; r++) CGP_UNROLL for (int q = 0; q <= r; q++) {
                double v = ev[r] * ev[q];
                if (Model::has_sig(r, q)) v += mdl.sig(r, q);
                a[D + sidx(r, q)] = tab.Wl * v;
            }
            CGP_UNROLL for (int q = 0; q < V; q++) {
                a[D + sidx(V, q)] = ev[q] * S0;
                a[D + sidx(V + 1, q)] = ev[q] * S1;
            }
            a[D + sidx(V, V)] = q00; a[D + sidx(V + 1, V)] = q10; a[D + sidx(V + 1, V + 1)] = q11;
            if constexpr (EXPORT) {
                CGP_UNROLL for (int q = 0; q < V; q++) xop[q][lane] = ev[q];
            }
        }
        double tot[NA];
        if constexpr (DBG == 1) { CGP_UNROLL for (int k = 0; k < NA; k++) tot[k] = a[k] * 27.; }
        else warp_sum_smem<NA>(a, red, res, lane, tot);
        CGP_UNROLL for (int r = 0; r < D; r++) mp[r] = tot[r];
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int q = 0; q <= r; q++)
            Pp[sidx(r, q)] = fma(-mp[r], mp[q], tot[D + sidx(r, q)]);
    }
};

// rhs of the continuous-discrete sigma-point moment ODE (filters_smoothers.py:124-137) for ModelSDE<NH>, one warp:
//   dm = sum_i w_i a(chi_i),   Q = sum_i w_i (chi_i - m) a(chi_i)^T,   dP = Q + Q^T + b b^T.
// The drift's chirp rows depend on chi[0..V] only (shared by the lane's P points); a[V] = chi[D-1] and
// a[V+1] = -gamma^2 chi[V] - 2 gamma chi[D-1] vary with the last coordinate.
template <int NH, int P> struct GhRhsSDE {
    using Model = ModelSDE<NH>;
    static constexpr int D = Model::D, V = Model::V, NS = NSym<D>::value, NA = D + D * D;
    Model mdl;
    GhLane<D, P> tab;
    double Qc[NS];
    CGP_DEV void load(const CgpProblem &p, int64_t b, int lane) {
        mdl.load(p.consts + b * p.consts_stride);
        tab.load(p, lane);
        load_sym<D>(p.Qc + b * p.Qc_stride, Qc);
    }
    CGP_DEV void rhs(double (*red)[kRedPitch], double *res, int lane, const double (&m)[D], const double (&Pc)[NS], double (&dm)[D],
                     double (&dP)[NS]) const {
        double L[NS], piv;                              // last pivot deferred behind the softplus branch, as in GhPredictLCD
        chol_lower_sym_rsqrt_head<D>(Pc, L, piv);
        double chi[D], slast;
        tab.points(m, L, chi, slast);
        const double w = (kTwoPi * fast_softplus_warp(chi[V])) * mdl.fs;
        L[sidx(D - 1, D - 1)] = piv * fast_rsqrt(piv);
        double f[D];
        chi[D - 1] = 0.;
        mdl.drift_w(w, chi, f);                         // f[0..V-1] final; f[V], f[V+1] recomputed per point below
        double S2 = 0., S3 = 0., DL = 0., q2 = 0., q3 = 0.;
        CGP_UNROLL for (int c = 0; c < P; c++) {
            const double cl = m[D - 1] + fma(L[sidx(D - 1, D - 1)], tab.xlast[c], slast);
            const double f2 = cl, f3 = fma(-mdl.tg, cl, -mdl.g2 * chi[V]);
            const double dl = cl - m[D - 1], wc = tab.wl[c];
            S2 = fma(wc, f2, S2);
            S3 = fma(wc, f3, S3);
            DL = fma(wc, dl, DL);
            q2 = fma(wc, dl * f2, q2);
            q3 = fma(wc, dl * f3, q3);
        }
        double a[NA];
        CGP_UNROLL for (int r = 0; r < V; r++) a[r] = tab.Wl * f[r];
        a[V] = S2; a[V + 1] = S3;
        CGP_UNROLL for (int r = 0; r < D - 1; r++) {
            const double dr = chi[r] - m[r];
            CGP_UNROLL for (int c = 0; c < V; c++) a[D + r * D + c] = tab.Wl * (dr * f[c]);
            a[D + r * D + V] = dr * S2;
            a[D + r * D + V + 1] = dr * S3;
        }
        CGP_UNROLL for (int c = 0; c < V; c++) a[D + (D - 1) * D + c] = DL * f[c];
        a[D + (D - 1) * D + V] = q2;
        a[D + (D - 1) * D + V + 1] = q3;
        double tot[NA];
        warp_sum_smem<NA>(a, red, res, lane, tot);
        CGP_UNROLL for (int r = 0; r < D; r++) dm[r] = tot[r];
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++)
            dP[sidx(r, c)] = (tot[D + r * D + c] + tot[D + c * D + r]) + Qc[sidx(r, c)];
    }
};

// Smoother workspace record [G | c | C] of one step (filters_smoothers.py:520-527, :81-84; layout: cgp_kernels.cuh) for
// ModelLCD<NH> from
//   Ein  (D-1) x V cross sums (GhPredictLCD::cross_partials),
//   tot  D + NSym moment totals (sum_i w_i mu_i | sum_i w_i (mu_i mu_i^T + Sigma), packed lower),
//   mPq  [m (D) | P packed (NSym)]: the filtering mean and covariance the prediction started from,
//   f    the Matern transition block [f00, f01, f10, f11] (models.py:61-73).
// L = chol(Pq);  D[:, q] = L E[:, q] for the chirp columns,  D[:, V+t] = f_t0 P[:, V] + f_t1 P[:, V+1] for the linear ones;
// mp = tot[0..D),  Pp = tot - mp mp^T;  G = D Pp^{-1} (row r of G solves Pp g = D_r^T);  c = m - G mp;  C = P - G D^T.
// `rec` may alias Ein / tot (they are read before anything is written); mPq must not overlap rec.
template <int NH>
CGP_DEV void gain_record(const double *Ein, const double *tot, const double *mPq, const double (&f)[4], double *rec) {
    constexpr int D = 2 * NH + 2, V = D - 2, NS = NSym<D>::value, DD = D * D, NE = (D - 1) * V;
    double Dx[D][D], mp[D], Lq[NS], rinv[D];
    {
        double E[NE], Pq[NS], L[NS];
        load_vec<NS>(mPq + D, Pq);
        chol_lower_sym_rsqrt<D>(Pq, L);
        load_vec<NE>(Ein, E);
        CGP_UNROLL for (int r = 0; r < D; r++) {
            CGP_UNROLL for (int q = 0; q < V; q++) {             // chirp columns of row r of D = L E  (E[D-1][.] = 0)
                double sacc = L[sidx(r, 0)] * E[q];
                CGP_UNROLL for (int k = 1; k <= r && k < D - 1; k++) sacc = fma(L[sidx(r, k)], E[k * V + q], sacc);
                Dx[r][q] = sacc;
            }
            Dx[r][V] = fma(f[1], Pq[sidx(r, V + 1)], f[0] * Pq[sidx(r, V)]);
            Dx[r][V + 1] = fma(f[3], Pq[sidx(r, V + 1)], f[2] * Pq[sidx(r, V)]);
        }
    }
    load_vec<D>(tot, mp);
    {
        double tp[NS], Pp[NS];
        load_vec<NS>(tot + D, tp);
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int q = 0; q <= r; q++)
            Pp[sidx(r, q)] = fma(-mp[r], mp[q], tp[sidx(r, q)]);              // as GhPredictLCD::predict_impl
        CGP_UNROLL for (int j = 0; j < D; j++) {                 // chol_lower_sym_rsqrt, keeping 1 / L_jj
            double sacc = Pp[sidx(j, j)];
            CGP_UNROLL for (int k = 0; k < j; k++) sacc = fma(-Lq[sidx(j, k)], Lq[sidx(j, k)], sacc);
            const double r = fast_rsqrt(sacc);
            rinv[j] = r;
            Lq[sidx(j, j)] = sacc * r;
            CGP_UNROLL for (int i = j + 1; i < D; i++) {
                double tacc = Pp[sidx(i, j)];
                CGP_UNROLL for (int k = 0; k < j; k++) tacc = fma(-Lq[sidx(i, k)], Lq[sidx(j, k)], tacc);
                Lq[sidx(i, j)] = tacc * r;
            }
        }
    }
    double cv[D];
    CGP_UNROLL for (int r = 0; r < D; r++) {
        double z[D];
        CGP_UNROLL for (int i = 0; i < D; i++) {
            double sacc = Dx[r][i];
            CGP_UNROLL for (int k = 0; k < i; k++) sacc = fma(-Lq[sidx(i, k)], z[k], sacc);
            z[i] = sacc * rinv[i];
        }
        CGP_UNROLL for (int i = D - 1; i >= 0; i--) {
            double sacc = z[i];
            CGP_UNROLL for (int k = i + 1; k < D; k++) sacc = fma(-Lq[sidx(k, i)], z[k], sacc);
            z[i] = sacc * rinv[i];
        }
        store_vec<D>(rec + r * D, z);
        double sacc = mPq[r];
        CGP_UNROLL for (int k = 0; k < D; k++) sacc = fma(-z[k], mp[k], sacc);
        cv[r] = sacc;
        CGP_UNROLL for (int q = 0; q <= r; q++) {                // C_rq = P_rq - (G D^T)_rq
            double cacc = mPq[D + sidx(r, q)];
            CGP_UNROLL for (int k = 0; k < D; k++) cacc = fma(-z[k], Dx[q][k], cacc);
            rec[DD + D + sidx(r, q)] = cacc;
        }
    }
    store_vec<D>(rec + DD, cv);
}

// sgp_filter (filters_smoothers.py:446-490, Pred = GhPredictLCD) / cd_sgp_filter (:534-582, Pred = GhRhsSDE + RK4) with a
// Gauss-Hermite table whose P^(D-1) base indices fit one warp.  ONE WARP PER CHIRP; mean, covariance and Cholesky
// factor are replicated in every lane.
//
// Output staging: lane 0 drops (m, P) of each step into a 32-step shared-memory ring and lane (t mod 32) keeps (S, r)
// of step t; every 32 steps the warp evaluates the 32 nll increments in SIMD (one log / sqrt / div per lane instead of
// one per step on the critical path), accumulates them in the reference's sequential order, and writes mfs / Pfs /
// nell with coalesced 16-byte stores.
template <class Pred, bool CD, bool H_E1>
__global__ void __launch_bounds__(32) gh_warp_filter_kernel(const CgpProblem p, const FilterIO io) {
    constexpr int D = Pred::D, NS = NSym<D>::value, NA = Pred::NA, DD = D * D, REC = D + DD;
    __shared__ __align__(16) double red[RedRows<NA>::value][kRedPitch];
    __shared__ __align__(16) double res[(NA + 1) & ~1];
    __shared__ __align__(16) double ring[32][REC];
    __shared__ double nl[32];
    const int lane = threadIdx.x;
    const int64_t b = blockIdx.x;
    Pred pred;
    pred.load(p, b, lane);
    double m[D], Pc[NS], H[D];
    load_vec<D>(p.m0 + b * p.m0_stride, m);
    load_sym<D>(p.P0 + b * p.P0_stride, Pc);
    CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
    const double *__restrict__ y = io.ys + (b / p.ys_repeat) * p.T;
    const int64_t T = p.T;
    const bool store_state = io.mfs != nullptr;
    const bool store_nell = io.nell != nullptr;
    const double dt = p.dt;
    double carry = 0.;                 // cumulative nll up to the last flushed step
    double Sk = 1., rk = 0.;           // (S, r) of the step this lane is responsible for
    double yv = (lane < T) ? __ldg(y + lane) : 0.;      // 32 measurements per load, broadcast by shuffle
    // Blocks of 32 steps: the inner loop is the chain and nothing else (32-bit counter, no end-of-block tests, the flush code
    // out of its instruction stream); the next block's measurements are in flight while the block runs.
    for (int64_t t0 = 0; t0 < T; t0 += 32) {
        const int n = (T - t0 < 32) ? (int)(T - t0) : 32;
        const double ynext = (t0 + 32 + lane < T) ? __ldg(y + t0 + 32 + lane) : 0.;
        for (int slot = 0; slot < n; slot++) {
            const double yt = __shfl_sync(0xffffffffu, yv, slot);
            double mp[D], Pp[NS];
            if constexpr (CD) {
                CGP_UNROLL for (int i = 0; i < D; i++) mp[i] = m[i];
                CGP_UNROLL for (int i = 0; i < NS; i++) Pp[i] = Pc[i];
                rk4_step<D>([&](const double (&mm)[D], const double (&PP)[NS], double (&dm)[D], double (&dP)[NS]) {
                    pred.rhs(red, res, lane, mm, PP, dm, dP);
                }, mp, Pp, dt);
            } else {
                pred.predict(red, res, lane, m, Pc, mp, Pp);
            }
            // ---- measurement update (filters_smoothers.py:55-68)
            double S, resid;
            linear_update_fast<D, H_E1>(mp, Pp, H, p.Xi, yt, m, Pc, S, resid);
            if (lane == slot) { Sk = S; rk = resid; }
            if (store_state && lane == 0) {
                store_vec<D>(&ring[slot][0], m);
                store_sym<D>(&ring[slot][D], Pc);
            }
        }
        yv = ynext;
        // ---- per block: nll increments in SIMD, sequential accumulation, coalesced stores
        nl[lane] = lane < n ? nll_increment(Sk, rk) : 0.;
        __syncwarp();
        if (lane == 0) {
            double c = carry;
            for (int j = 0; j < n; j++) { c = c + nl[j]; nl[j] = c; }     // reference order: n_ell = n_ell + inc
        }
        __syncwarp();
        carry = nl[n - 1];
        if (store_nell && !io.nell_last_only && lane < n) io.nell[b * T + t0 + lane] = nl[lane];
        if (store_state) {
            double2 *dm = reinterpret_cast<double2 *>(io.mfs + (b * T + t0) * D);
            for (int i = lane; i < n * (D / 2); i += 32)
                dm[i] = *reinterpret_cast<const double2 *>(&ring[i / (D / 2)][2 * (i % (D / 2))]);
            double2 *dP = reinterpret_cast<double2 *>(io.Pfs + (b * T + t0) * DD);
            for (int i = lane; i < n * (DD / 2); i += 32)
                dP[i] = *reinterpret_cast<const double2 *>(&ring[i / (DD / 2)][D + 2 * (i % (DD / 2))]);
        }
        __syncwarp();
    }
    if (store_nell && io.nell_last_only && lane == 0) io.nell[b] = carry;
}

// cd_sgp_smoother (filters_smoothers.py:585-632) for ModelSDE<NH> with a Gauss-Hermite table: one warp per chirp, RK4
// backwards in time.  rhs (:615-621): Gm = Pf^{-1} gamma (hoisted out of the 4 stages), (_m, _P) = cd_sgp_common(m, P),
// dm = _m + Gm^T (m - mf),  dP = _P + Gm^T P + P Gm - 2 gamma.   Results go through a 32-step ring like the filter.
// Everything that depends on the FILTERING result only -- the loads of (mf_k, Pf_k), chol(Pf_k) and the four solves of
// Gm_k -- is taken off the sequential chain: once per 32-step block lane j does it for step j of the block (SIMD over
// time) and leaves [mf | Gm] in shared memory, where the time loop picks it up as a broadcast read.
template <int NH, int P>
__global__ void __launch_bounds__(32) cd_ghs_warp_kernel(const CgpProblem p, const SmootherIO io) {
    using Rhs = GhRhsSDE<NH, P>;
    constexpr int D = Rhs::D, NS = NSym<D>::value, NA = Rhs::NA, DD = D * D, REC = D + DD;
    constexpr int PROW = (((D + DD) / 2) % 2 == 1) ? D + DD : D + DD + 2;      // odd number of 16-byte units per row
    __shared__ __align__(16) double red[RedRows<NA>::value][kRedPitch];
    __shared__ __align__(16) double res[(NA + 1) & ~1];
    __shared__ __align__(16) double ring[32][REC];
    __shared__ __align__(16) double pre[32][PROW];                             // per step of the block: mf | Gm
    const int lane = threadIdx.x;
    const int64_t b = blockIdx.x;
    const int64_t T = p.T;
    Rhs rhs;
    rhs.load(p, b, lane);
    double Qf[D][D];
    sym_to_full<D>(rhs.Qc, Qf);
    double ms[D], Ps[NS];
    load_vec<D>(io.mfs + (b * T + T - 1) * D, ms);
    load_sym<D>(io.Pfs + (b * T + T - 1) * DD, Ps);
    if (lane == 0) {
        store_vec<D>(io.mss + (b * T + T - 1) * D, ms);
        store_sym<D>(io.Pss + (b * T + T - 1) * DD, Ps);
    }
    const double ndt = -p.dt;
    // walk backwards; ring slot s holds step t with (t & 31) == s; flush when a 32-aligned block is complete
    for (int64_t t = T - 2; t >= 0; t--) {
        if (t == T - 2 || (t & 31) == 31) {            // first step of a 32-aligned block: lane j prepares step (t & ~31) + j
            const int64_t tj = (t & ~(int64_t)31) + lane;
            if (tj <= t) {
                double mfj[D], Pf[D][D], Lf[D][D], rinv[D];
                load_vec<D>(io.mfs + (b * T + tj) * D, mfj);
                load_mat<D>(io.Pfs + (b * T + tj) * DD, Pf);
                chol_lower_rsqrt<D>(Pf, Lf, rinv);
                store_vec<D>(&pre[lane][0], mfj);
                CGP_UNROLL for (int c = 0; c < D; c++) {       // Gm = Pf^{-1} gamma, column by column; stored as Gm^T rows
                    double col[D];
                    CGP_UNROLL for (int i = 0; i < D; i++) col[i] = Qf[i][c];
                    chol_solve_vec_rinv<D>(Lf, rinv, col);
                    store_vec<D>(&pre[lane][D + c * D], col);
                }
            }
            __syncwarp();
        }
        double mf[D], Gm[D][D];
        {
            const double *row = &pre[t & 31][0];
            load_vec<D>(row, mf);
            CGP_UNROLL for (int c = 0; c < D; c++) {
                double col[D];
                load_vec<D>(row + D + c * D, col);
                CGP_UNROLL for (int i = 0; i < D; i++) Gm[i][c] = col[i];
            }
        }
        rk4_step<D>([&](const double (&mm)[D], const double (&PP)[NS], double (&dm)[D], double (&dP)[NS]) {
            double _m[D], _P[NS], W[D][D];
            rhs.rhs(red, res, lane, mm, PP, _m, _P);
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
                dP[sidx(r, c)] = ((_P[sidx(r, c)] + W[r][c]) + W[c][r]) - 2 * rhs.Qc[sidx(r, c)];
        }, ms, Ps, ndt);
        const int slot = (int)(t & 31);
        if (lane == 0) {
            store_vec<D>(&ring[slot][0], ms);
            store_sym<D>(&ring[slot][D], Ps);
        }
        if (slot == 0 || t == 0) {
            // steps [t, hi] are in the ring (hi = last step of this 32-block that is < T-1)
            const int64_t hi = ((t | 31) < T - 2) ? (t | 31) : (T - 2);
            const int n = (int)(hi - t + 1);
            __syncwarp();
            double2 *dm2 = reinterpret_cast<double2 *>(io.mss + (b * T + t) * D);
            for (int i = lane; i < n * (D / 2); i += 32)
                dm2[i] = *reinterpret_cast<const double2 *>(&ring[(slot + i / (D / 2)) & 31][2 * (i % (D / 2))]);
            double2 *dP2 = reinterpret_cast<double2 *>(io.Pss + (b * T + t) * DD);
            for (int i = lane; i < n * (DD / 2); i += 32)
                dP2[i] = *reinterpret_cast<const double2 *>(&ring[(slot + i / (DD / 2)) & 31][D + 2 * (i % (DD / 2))]);
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------ smoother sweep, warp per chirp
// Sequential part of rts / eks / sgp_smoother (filters_smoothers.py:83-84):
//     ms = mf + G (ms - mp),   Ps = Pf + G (Ps - Pp) G^T        for k = T-2 .. 0.
// One warp owns one chirp.  Tiles of TS consecutive [G | c | C] records (cgp_kernels.cuh) are staged in shared memory with
// cp.async (double buffered), the recursion  ms = c + G ms',  Ps = C + G Ps' G^T  runs with ONE MATRIX ENTRY PER LANE
// (operands exchanged through shared memory), and the TS results are written back with coalesced stores.
template <int D> struct SweepCfg {
    static constexpr int R = ws_record<D>();                  // workspace record
    static constexpr int TS = (D <= 4) ? 16 : 8;
    static constexpr int OUT = D + D * D;                     // ms | Ps per step
    static constexpr int TILE_DOUBLES = TS * R;
    static constexpr int WARPS = 1;
    static constexpr size_t smem_bytes() {
        return sizeof(double) * (2 * TILE_DOUBLES + TS * OUT + 2 * D * D + 2 * D);
    }
};

CGP_DEV void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
CGP_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> CGP_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int D>
__global__ void __launch_bounds__(32) smoother_sweep_warp_kernel(const CgpProblem p, const SmootherIO io) {
    using Cfg = SweepCfg<D>;
    constexpr int R = Cfg::R, TS = Cfg::TS, DD = D * D, OUT = Cfg::OUT, TILE = Cfg::TILE_DOUBLES;
    constexpr int EPL = (DD + 31) / 32;                    // matrix entries per lane
    // shared-memory map (doubles): two input tiles, the output tile, X = Ps', T1 = G X, the mean ms'
    constexpr int O_OUT = 2 * TILE, O_X = O_OUT + TS * OUT, O_T1 = O_X + DD, O_DM = O_T1 + DD;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x;
    const int64_t b = blockIdx.x;
    const int64_t T = p.T;
    const double *__restrict__ ws = io.ws + b * T * R;
    const double *__restrict__ mfs = io.mfs + b * T * D;
    const double *__restrict__ Pfs = io.Pfs + b * T * DD;
    double *__restrict__ mss = io.mss + b * T * D;
    double *__restrict__ Pss = io.Pss + b * T * DD;

    // entry e of the covariance handled by this lane (clamped duplicates keep every lane busy: no divergence)
    int er[EPL], ec[EPL], es[EPL];
    bool own[EPL];
    CGP_UNROLL for (int q = 0; q < EPL; q++) {
        const int e = lane + 32 * q;
        own[q] = e < DD;
        const int ee = own[q] ? e : e % DD;
        er[q] = ee / D; ec[q] = ee % D;
        const int rr = er[q] > ec[q] ? er[q] : ec[q], cc = er[q] > ec[q] ? ec[q] : er[q];
        es[q] = rr * (rr + 1) / 2 + cc;                    // position of (r, c) in the packed lower triangle C
    }
    const int mr = lane % D;                               // mean component handled by this lane
    // last step: copy the filter result (filters_smoothers.py:140-142)
    double Pcur[EPL], mcur;
    CGP_UNROLL for (int q = 0; q < EPL; q++) {
        Pcur[q] = Pfs[(T - 1) * DD + er[q] * D + ec[q]];
        if (own[q]) Pss[(T - 1) * DD + er[q] * D + ec[q]] = Pcur[q];
    }
    mcur = mfs[(T - 1) * D + mr];
    if (lane < D) mss[(T - 1) * D + mr] = mcur;
    if (T < 2) return;

    auto issue_tile = [&](int buf, int64_t lo, int n) {
        double *dst = smem + buf * TILE;
        const double *s0 = ws + lo * R;
        for (int i = lane; i < n * R / 2; i += 32) cp_async16(dst + 2 * i, s0 + 2 * i);
        cp_async_commit();
    };
    int64_t hi = T - 1;                                    // steps [lo, hi) of the current tile, walking backwards
    int buf = 0;
    {
        const int n = (int)(hi < TS ? hi : TS);
        issue_tile(0, hi - n, n);
    }
    while (hi > 0) {
        const int n = (int)(hi < TS ? hi : TS);
        const int64_t lo = hi - n;
        if (lo > 0) {
            const int nn = (int)(lo < TS ? lo : TS);
            issue_tile(buf ^ 1, lo - nn, nn);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const int o_ws = buf * TILE;
        for (int j = n - 1; j >= 0; j--) {
            const int oG = o_ws + j * R, oc = oG + DD, oC = oc + D;
            // X = Ps' ; the mean ms'
            CGP_UNROLL for (int q = 0; q < EPL; q++) smem[O_X + er[q] * D + ec[q]] = Pcur[q];
            smem[O_DM + mr] = mcur;
            __syncwarp();
            // T1 = G X ; ms = c + G ms'
            CGP_UNROLL for (int q = 0; q < EPL; q++) {
                double s = smem[oG + er[q] * D] * smem[O_X + ec[q]];
                CGP_UNROLL for (int k = 1; k < D; k++) s = fma(smem[oG + er[q] * D + k], smem[O_X + k * D + ec[q]], s);
                smem[O_T1 + er[q] * D + ec[q]] = s;
            }
            {
                double s = smem[oG + mr * D] * smem[O_DM];
                CGP_UNROLL for (int k = 1; k < D; k++) s = fma(smem[oG + mr * D + k], smem[O_DM + k], s);
                mcur = smem[oc + mr] + s;
                smem[O_OUT + j * OUT + mr] = mcur;
            }
            __syncwarp();
            // Ps = C + T1 G^T
            CGP_UNROLL for (int q = 0; q < EPL; q++) {
                double s = smem[O_T1 + er[q] * D] * smem[oG + ec[q] * D];
                CGP_UNROLL for (int k = 1; k < D; k++) s = fma(smem[O_T1 + er[q] * D + k], smem[oG + ec[q] * D + k], s);
                Pcur[q] = smem[oC + es[q]] + s;
                smem[O_OUT + j * OUT + D + er[q] * D + ec[q]] = Pcur[q];
            }
            __syncwarp();
        }
        // coalesced write-back of the n finished steps
        for (int i = lane; i < n * D; i += 32) mss[lo * D + i] = smem[O_OUT + (i / D) * OUT + (i % D)];
        for (int i = lane; i < n * DD; i += 32) Pss[lo * DD + i] = smem[O_OUT + (i / DD) * OUT + D + (i % DD)];
        __syncwarp();
        hi = lo;
        buf ^= 1;
    }
}

// ------------------------------------------------------------------------------------------------ CD-EKF / CD-EKS, 16 lanes per chirp
// cd_ekf (filters_smoothers.py:352-397) and cd_eks (:400-443) for the chirp SDE (d = 4) when the batch is too small to
// fill the GPU with one thread per chirp: a HALF-WARP owns one chirp, lane (i, j) = (l / 4, l % 4) holds the covariance
// entry P_ij, the mean is replicated.  X = J P needs column j of P (shuffles from lanes (k, j)) and the sparse row i of
// the drift Jacobian; dP = X + X^T + b b^T needs X_ji (one shuffle).  Exactly the thread-per-chirp arithmetic, spread
// over 16 lanes; the 16 entries of a step are written with one coalesced 128-byte store.
struct HalfWarp {
    int l, i, j, base;          // lane within the half-warp, row, column, first lane of the half within the warp
    CGP_DEV explicit HalfWarp(int lane) : l(lane & 15), i((lane & 15) >> 2), j(lane & 3), base(lane & 16) {}
    CGP_DEV double get(double v, int src) const { return __shfl_sync(0xffffffffu, v, base + src); }
    CGP_DEV double sum16(double v) const {
        CGP_UNROLL for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        return v;
    }
};

// row i of the drift Jacobian of the chirp SDE (models.py:104-110): J = [[-lam, -w, -w' u1, 0], [w, -lam, w' u0, 0],
// [0, 0, 0, 1], [0, 0, -gamma^2, -2 gamma]], for given w = 2 pi g(V) fs and w' = 2 pi g'(V) fs
CGP_DEV void chirp_drift_and_jrow_w(const ModelSDE<1> &mdl, const HalfWarp &hw, const double (&m)[4], double w, double dw,
                                    double (&a)[4], double (&jr)[4]) {
    mdl.drift_w(w, m, a);
    const bool r0 = hw.i == 0, r1 = hw.i == 1, r2 = hw.i == 2;
    jr[0] = r0 ? -mdl.lam : (r1 ? w : 0.);
    jr[1] = r0 ? -w : (r1 ? -mdl.lam : 0.);
    jr[2] = r0 ? -dw * m[1] : (r1 ? dw * m[0] : (r2 ? 0. : -mdl.g2));
    jr[3] = (r0 || r1) ? 0. : (r2 ? 1. : -mdl.tg);
}
CGP_DEV void chirp_drift_and_jrow(const ModelSDE<1> &mdl, const HalfWarp &hw, const double (&m)[4], double (&a)[4], double (&jr)[4]) {
    double gv, sg;
    softplus_and_sigmoid(m[2], gv, sg);
    chirp_drift_and_jrow_w(mdl, hw, m, (kTwoPi * gv) * mdl.fs, (kTwoPi * sg) * mdl.fs, a, jr);
}
// select without a branch (nested ?: over registers are sometimes compiled to BSSY / BRA / BSYNC, ~35 cycles each on a
// single-warp dependency chain)
CGP_DEV double selp(bool c, double a, double b) {
    double r;
    asm("{ .reg .pred p; setp.ne.s32 p, %3, 0; selp.f64 %0, %1, %2, p; }" : "=d"(r) : "d"(a), "d"(b), "r"((int)c));
    return r;
}
// Per-lane constants of row i of the chirp SDE's drift Jacobian and the row itself for given (w, w'): two selects on the
// operands that vary instead of rebuilding the row by seven.
struct ChirpJRow {
    bool r0, r1, r01, j0, j1, j2;
    double c0k, c1k, c2k, c3k, sgn;
    CGP_DEV ChirpJRow(const ModelSDE<1> &mdl, const HalfWarp &hw)
        : r0(hw.i == 0), r1(hw.i == 1), r01(hw.i < 2), j0(hw.j == 0), j1(hw.j == 1), j2(hw.j == 2) {
        c0k = r0 ? -mdl.lam : 0.;
        c1k = r1 ? -mdl.lam : 0.;
        c2k = hw.i == 3 ? -mdl.g2 : 0.;
        c3k = r01 ? 0. : (hw.i == 2 ? 1. : -mdl.tg);
        sgn = r0 ? -1. : 1.;
    }
    CGP_DEV void row(const double (&m)[4], double w, double dw, double (&jr)[4]) const {
        jr[0] = selp(r1, w, c0k);
        jr[1] = selp(r0, -w, c1k);
        jr[2] = selp(r01, (sgn * dw) * selp(r0, m[1], m[0]), c2k);
        jr[3] = c3k;
    }
    CGP_DEV double by_col(const double (&v)[4]) const { return selp(j0, v[0], selp(j1, v[1], selp(j2, v[2], v[3]))); }
};
// The (V, V') block of the chirp SDE is linear and never sees the oscillator states, so in the FILTER (rhs = a(m)) the four
// RK4 stage arguments of V follow from (V, V') alone: the step's four softplus / sigmoid evaluations are independent of each
// other and of the covariance chain.  Lane (., j) of the half-warp evaluates stage j's (one evaluation per lane instead of
// four (branch + exp + series) in sequence on the chain -- a single warp pays ~3 cycles per instruction it issues, replicated
// or not) and the four (w, w') pairs are gathered by shuffle.  Same expressions as rk4_step_lane for the stage arguments.
// The side of the softplus range split is chosen per chirp (half-warp): a chirp's result does not depend on its neighbour.
CGP_DEV void chirp_stage_frequencies(const ModelSDE<1> &mdl, const HalfWarp &hw, const ChirpJRow &jc, double v, double vd,
                                     double dt, double (&w)[4], double (&dw)[4]) {
    double x[4];
    x[0] = v;
    double kv = vd, kd = fma(-mdl.tg, vd, -mdl.g2 * v);
    double tv = v + dt * kv * 0.5, td = vd + dt * kd * 0.5;
    x[1] = tv;
    kv = td; kd = fma(-mdl.tg, td, -mdl.g2 * tv);
    tv = v + dt * kv * 0.5; td = vd + dt * kd * 0.5;
    x[2] = tv;
    kv = td;
    x[3] = v + dt * kv;
    const double xs = jc.by_col(x);
    const unsigned bal = __ballot_sync(0xffffffffu, softplus_in_series_range(xs));
    double gv, sg;
    if (((bal >> hw.base) & 0xffffu) == 0xffffu) softplus_sigmoid_series(xs, gv, sg);
    else softplus_sigmoid_general(xs, gv, sg);
    const double ws = (kTwoPi * gv) * mdl.fs, dws = (kTwoPi * sg) * mdl.fs;
    CGP_UNROLL for (int q = 0; q < 4; q++) { w[q] = hw.get(ws, q); dw[q] = hw.get(dws, q); }
}
// (A P)_ij for the lane's (i, j), with `arow` = row i of A and P distributed one entry per lane
CGP_DEV double row_times_P(const HalfWarp &hw, const double (&arow)[4], double Pe) {
    double s = arow[0] * hw.get(Pe, hw.j);
    CGP_UNROLL for (int k = 1; k < 4; k++) s = fma(arow[k], hw.get(Pe, 4 * k + hw.j), s);
    return s;
}
template <class Ode> CGP_DEV void rk4_step_lane(Ode &&ode, double (&m)[4], double &Pe, double dt) {   // ode(stage, m, P, dm, dP)
    double km[4], kP, am[4], aP, tm[4], tP;
    ode(0, m, Pe, km, kP);
    CGP_UNROLL for (int q = 0; q < 4; q++) { am[q] = km[q]; tm[q] = m[q] + dt * km[q] * 0.5; }
    aP = kP; tP = Pe + dt * kP * 0.5;
    ode(1, tm, tP, km, kP);
    CGP_UNROLL for (int q = 0; q < 4; q++) { am[q] = am[q] + 2 * km[q]; tm[q] = m[q] + dt * km[q] * 0.5; }
    aP = aP + 2 * kP; tP = Pe + dt * kP * 0.5;
    ode(2, tm, tP, km, kP);
    CGP_UNROLL for (int q = 0; q < 4; q++) { am[q] = am[q] + 2 * km[q]; tm[q] = m[q] + dt * km[q]; }
    aP = aP + 2 * kP; tP = Pe + dt * kP;
    ode(3, tm, tP, km, kP);
    constexpr double kSixth = 1. / 6.;
    CGP_UNROLL for (int q = 0; q < 4; q++) m[q] = m[q] + dt * (am[q] + km[q]) * kSixth;
    Pe = Pe + dt * (aP + kP) * kSixth;
}

template <int NH, bool H_E1>   // NH == 1: the chirp SDE (d = 4 -> 16 covariance entries -> half a warp); H_E1: H is exactly e_1
__global__ void __launch_bounds__(32) cd_ekf_lane_kernel(const CgpProblem p, const FilterIO io) {
    static_assert(NH == 1, "16 lanes per chirp need d == 4");
    using Model = ModelSDE<1>;
    constexpr int D = 4, DD = 16, TB = 16;                   // time is walked in blocks of 16 steps (one nll slot per lane)
    __shared__ double nl[2][TB];
    const int lane = threadIdx.x;
    const HalfWarp hw(lane);
    const int half = lane >> 4;
    const int64_t gid = (int64_t)blockIdx.x * 2 + half;
    const bool active = gid < p.B;
    const int64_t b = active ? gid : p.B - 1;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride);
    const ChirpJRow jc(mdl, hw);
    double m[D], H[D];
    load_vec<D>(p.m0 + b * p.m0_stride, m);
    CGP_UNROLL for (int q = 0; q < D; q++) H[q] = p.H[q];
    // exactly symmetric covariance: read the lower triangle only (what the packed-symmetric kernels do)
    const int ii = hw.i > hw.j ? hw.i : hw.j, jj = hw.i > hw.j ? hw.j : hw.i;
    double Pe = (p.P0 + b * p.P0_stride)[ii * D + jj];
    const double Qe = (p.Qc + b * p.Qc_stride)[ii * D + jj];
    const double hj = jc.by_col(H);
    const double *__restrict__ y = io.ys + (b / p.ys_repeat) * p.T;
    const int64_t T = p.T;
    const bool store_state = io.mfs != nullptr && active;
    const bool store_m = store_state && hw.l < D;
    const bool store_nell = io.nell != nullptr && active;
    const double dt = p.dt, Xi = p.Xi;
    double *pP = store_state ? io.Pfs + b * T * DD + hw.l : nullptr;
    double *pm = store_state ? io.mfs + b * T * D + hw.j : nullptr;
    double carry = 0.;
    double ynext = (hw.l < T) ? __ldg(y + hw.l) : 0.;
    for (int64_t t0 = 0; t0 < T; t0 += TB) {
        const int n = (int)(T - t0 < TB ? T - t0 : TB);
        const double ycur = ynext;
        if (t0 + TB < T) ynext = (t0 + TB + hw.l < T) ? __ldg(y + t0 + TB + hw.l) : 0.;   // next block's samples in flight
        double Sk = 1., rk = 0.;
        // 32-bit inner loop that holds the chain and nothing else (no flush test, no 64-bit trip count)
        for (int s = 0; s < n; s++) {
            const double yt = hw.get(ycur, s);
            double wst[4], dwst[4];
            chirp_stage_frequencies(mdl, hw, jc, m[2], m[3], dt, wst, dwst);
            rk4_step_lane([&](int stage, const double (&mm)[D], double PP, double (&dm)[D], double &dP) {
                double jr[D];
                mdl.drift_w(wst[stage], mm, dm);
                jc.row(mm, wst[stage], dwst[stage], jr);
                const double X = row_times_P(hw, jr, PP);
                const double Xt = hw.get(X, 4 * hw.j + hw.i);
                dP = (Xt + X) + Qe;                                  // P J^T + J P + b b^T  (filters_smoothers.py:385)
            }, m, Pe, dt);
            // ---- measurement update (filters_smoothers.py:55-68): c = P h, S = h^T c + Xi
            double cr, cj, c[D], S, pred;
            if constexpr (H_E1) {
                // H = e_1: P h = column 1 of P, h^T P h = P_11, h^T m = m_1 (what the sums below give, adding exact zeros)
                CGP_UNROLL for (int q = 0; q < D; q++) c[q] = hw.get(Pe, 4 * q + 1);
                cr = hw.get(Pe, 4 * hw.i + 1);
                cj = hw.get(Pe, 4 * hw.j + 1);
                S = c[1] + Xi;
                pred = m[1];
            } else {
                cr = Pe * hj;                                        // row sums: c_i = sum_j P_ij h_j, in every lane of row i
                cr += __shfl_xor_sync(0xffffffffu, cr, 1);
                cr += __shfl_xor_sync(0xffffffffu, cr, 2);
                CGP_UNROLL for (int q = 0; q < D; q++) c[q] = hw.get(cr, 4 * q);
                cj = hw.get(cr, 4 * hw.j);
                S = H[0] * c[0];
                CGP_UNROLL for (int q = 1; q < D; q++) S = fma(H[q], c[q], S);
                S += Xi;
                pred = H[0] * m[0];
                CGP_UNROLL for (int q = 1; q < D; q++) pred = fma(H[q], m[q], pred);
            }
            const double rS = fast_rcp(S), resid = yt - pred;
            CGP_UNROLL for (int q = 0; q < D; q++) m[q] = fma(c[q] * rS, resid, m[q]);
            Pe = fma(-((cr * rS) * (cj * rS)), S, Pe);               // P - K K^T S, K = c / S
            Sk = selp(hw.l == s, S, Sk);
            rk = selp(hw.l == s, resid, rk);
            if (store_state) {
                *pP = Pe;
                if (store_m) *pm = jc.by_col(m);
                pP += DD;
                pm += D;
            }
        }
        nl[half][hw.l] = hw.l < n ? nll_increment(Sk, rk) : 0.;
        __syncwarp();
        if (hw.l == 0) {
            double cc = carry;
            for (int q = 0; q < n; q++) { cc = cc + nl[half][q]; nl[half][q] = cc; }
        }
        __syncwarp();
        carry = nl[half][n - 1];
        if (store_nell && !io.nell_last_only && hw.l < n) io.nell[b * T + t0 + hw.l] = nl[half][hw.l];
        __syncwarp();
    }
    if (store_nell && io.nell_last_only && hw.l == 0) io.nell[b] = carry;
}

// ------------------------------------------------------------------------------------------------ discrete EKF, 16 lanes per chirp
// ekf (filters_smoothers.py:222-264) for the chirp LCD model (d = 4) when the batch cannot fill the GPU with one thread per chirp
// (config 1: a single chirp): the layout of cd_ekf_lane_kernel -- lane (i, j) of a half-warp holds P_ij (a full matrix here: J P J^T
// is symmetric only to rounding and the reference keeps it as it comes), the mean is replicated.  The closed-form Jacobian of the
// mean (ModelLCD::mean_jac: rotation block, its column V, Matern block) is evaluated once per step by every lane; the lane picks its
// rows i and j with selects; X = J P takes column j of P by shuffle, Pp = X J^T + Sigma takes row i of X by shuffle.  The
// measurement update follows linear_update: S from the column sums H^T Pp, the gain from the row sums Pp H.  ~300 instructions per
// step in the warp instead of ~600 in a single thread (a lone warp pays ~3 cycles per instruction it issues).
template <int NH, bool H_E1>      // H_E1: H is exactly e_1 (CgpProblem::h_unit_index == 1) -- S, the gain and the prediction are entries, not sums
__global__ void __launch_bounds__(32) ekf_lane_kernel(const CgpProblem p, const FilterIO io) {
    static_assert(NH == 1, "16 lanes per chirp need d == 4");
    using Model = ModelLCD<1>;
    constexpr int D = 4, DD = 16, TB = 16;
    __shared__ double nl[2][TB];
    const int lane = threadIdx.x;
    const HalfWarp hw(lane);
    const int half = lane >> 4;
    const int64_t gid = (int64_t)blockIdx.x * 2 + half;
    const bool active = gid < p.B;
    const int64_t b = active ? gid : p.B - 1;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    double m[D], H[D];
    load_vec<D>(p.m0 + b * p.m0_stride, m);
    CGP_UNROLL for (int q = 0; q < D; q++) H[q] = p.H[q];
    double Pe = (p.P0 + b * p.P0_stride)[hw.l];
    // per-lane constants: which rows of J the lane needs (i for J P, j for X J^T), its entry of Sigma, its entries of H
    const bool i0 = hw.i == 0, iu = hw.i < 2, i2 = hw.i == 2, j0 = hw.j == 0, ju = hw.j < 2, j2 = hw.j == 2, j1 = hw.j == 1;
    const double fi2 = i2 ? mdl.f00 : mdl.f10, fi3 = iu ? 0. : (i2 ? mdl.f01 : mdl.f11);
    const double fj2 = j2 ? mdl.f00 : mdl.f10, fj3 = ju ? 0. : (j2 ? mdl.f01 : mdl.f11);
    double sg_ij = 0.;
    if (hw.i == hw.j) sg_ij = iu ? mdl.q : (i2 ? mdl.s00 : mdl.s11);
    else if (!iu && !ju) sg_ij = mdl.s01;
    const double hj = selp(j0, H[0], selp(j1, H[1], selp(j2, H[2], H[3])));
    const double hi = selp(i0, H[0], selp(hw.i == 1, H[1], selp(i2, H[2], H[3])));
    const double *__restrict__ y = io.ys + (b / p.ys_repeat) * p.T;
    const int64_t T = p.T;
    const bool store_state = io.mfs != nullptr && active;
    const bool store_m = store_state && hw.l < D;
    const bool store_nell = io.nell != nullptr && active;
    const double Xi = p.Xi, dt1 = mdl.dt * 1.;                 // dt * (k + 1), k = 0 (mean_jac)
    double *pP = store_state ? io.Pfs + b * T * DD + hw.l : nullptr;
    double *pm = store_state ? io.mfs + b * T * D + hw.j : nullptr;
    double carry = 0.;
    double ynext = (hw.l < T) ? __ldg(y + hw.l) : 0.;
    for (int64_t t0 = 0; t0 < T; t0 += TB) {
        const int n = (int)(T - t0 < TB ? T - t0 : TB);
        const double ycur = ynext;
        if (t0 + TB < T) ynext = (t0 + TB + hw.l < T) ? __ldg(y + t0 + TB + hw.l) : 0.;
        double Sk = 1., rk = 0.;
        for (int s = 0; s < n; s++) {
            const double yt = hw.get(ycur, s);
            // ---- prediction: mean, Jacobian of the mean (:255-256), Pp = J P J^T + Sigma (:257)
            double gv, sg;
            softplus_and_sigmoid(m[2], gv, sg);
            const double w = (kTwoPi * gv) * mdl.fs, dw = (kTwoPi * sg) * mdl.fs;
            double sn, cs;
            fast_sincos(dt1 * w, &sn, &cs);
            const double ce = cs * mdl.e, se = sn * mdl.e, dth = dt1 * dw;
            double mp[D];
            mp[0] = fma(-se, m[1], ce * m[0]);
            mp[1] = fma(ce, m[1], se * m[0]);
            const double J02 = -mp[1] * dth, J12 = mp[0] * dth;
            mp[2] = fma(mdl.f01, m[3], mdl.f00 * m[2]);
            mp[3] = fma(mdl.f11, m[3], mdl.f10 * m[2]);
            double jr[D], jc[D];
            jr[0] = selp(iu, selp(i0, ce, se), 0.);
            jr[1] = selp(iu, selp(i0, -se, ce), 0.);
            jr[2] = selp(iu, selp(i0, J02, J12), fi2);
            jr[3] = fi3;
            jc[0] = selp(ju, selp(j0, ce, se), 0.);
            jc[1] = selp(ju, selp(j0, -se, ce), 0.);
            jc[2] = selp(ju, selp(j0, J02, J12), fj2);
            jc[3] = fj3;
            const double X = row_times_P(hw, jr, Pe);                 // (J P)_ij
            double Pp = hw.get(X, 4 * hw.i) * jc[0];                  // (X J^T)_ij = sum_k X_ik J_jk
            CGP_UNROLL for (int k = 1; k < D; k++) Pp = fma(hw.get(X, 4 * hw.i + k), jc[k], Pp);
            Pp += sg_ij;
            // ---- measurement update (filters_smoothers.py:55-68, linear_update): S = (H^T Pp) H + Xi, K = Pp H / S
            double S, c[D], cr, cj, pred;
            if constexpr (H_E1) {
                // H = e_1: H^T Pp H = Pp_11, Pp H = column 1 of Pp, H mp = mp_1 (what the sums below give, adding exact zeros)
                S = hw.get(Pp, 5) + Xi;
                CGP_UNROLL for (int q = 0; q < D; q++) c[q] = hw.get(Pp, 4 * q + 1);
                cr = hw.get(Pp, 4 * hw.i + 1);
                cj = hw.get(Pp, 4 * hw.j + 1);
                pred = mp[1];
            } else {
                double cc = hi * Pp;                                 // column sums over i
                cc += __shfl_xor_sync(0xffffffffu, cc, 4);
                cc += __shfl_xor_sync(0xffffffffu, cc, 8);
                cr = Pp * hj;                                        // row sums over j
                cr += __shfl_xor_sync(0xffffffffu, cr, 1);
                cr += __shfl_xor_sync(0xffffffffu, cr, 2);
                S = hw.get(cc, 0) * H[0];
                CGP_UNROLL for (int q = 1; q < D; q++) S = fma(hw.get(cc, q), H[q], S);
                S += Xi;
                CGP_UNROLL for (int q = 0; q < D; q++) c[q] = hw.get(cr, 4 * q);
                cj = hw.get(cr, 4 * hw.j);
                pred = H[0] * mp[0];
                CGP_UNROLL for (int q = 1; q < D; q++) pred = fma(H[q], mp[q], pred);
            }
            const double rS = fast_rcp(S), resid = yt - pred;
            CGP_UNROLL for (int q = 0; q < D; q++) m[q] = fma(c[q] * rS, resid, mp[q]);
            Pe = Pp - ((cr * rS) * (cj * rS)) * S;
            Sk = selp(hw.l == s, S, Sk);
            rk = selp(hw.l == s, resid, rk);
            if (store_state) {
                *pP = Pe;
                if (store_m) *pm = selp(j0, m[0], selp(j1, m[1], selp(j2, m[2], m[3])));
                pP += DD;
                pm += D;
            }
        }
        nl[half][hw.l] = hw.l < n ? nll_increment(Sk, rk) : 0.;
        __syncwarp();
        if (hw.l == 0) {
            double acc = carry;
            for (int q = 0; q < n; q++) { acc = acc + nl[half][q]; nl[half][q] = acc; }
        }
        __syncwarp();
        carry = nl[half][n - 1];
        if (store_nell && !io.nell_last_only && hw.l < n) io.nell[b * T + t0 + hw.l] = nl[half][hw.l];
        __syncwarp();
    }
    if (store_nell && io.nell_last_only && hw.l == 0) io.nell[b] = carry;
}

// cd_eks: rhs (filters_smoothers.py:427-432) gamma = b b^T, M = J_a(m) + (Pf^{-1} gamma)^T, dm = a(m) + gamma Pf^{-1} (m - mf),
// dP = M P + P M^T - gamma.  X = Pf^{-1} gamma depends on the filtering result only: once per 16-step block lane j of the
// half-warp loads (mf, Pf) of step j, factorises Pf and solves for X (SIMD over time), and the time loop reads [mf | X] back
// from shared memory.  With gamma symmetric, gamma Pf^{-1} = X^T, so dm = a + X^T (m - mf): a 4 x 4 product on the chain instead
// of the reference's two triangular solves per stage (same value up to rounding).
template <int NH>
__global__ void __launch_bounds__(32) cd_eks_lane_kernel(const CgpProblem p, const SmootherIO io) {
    static_assert(NH == 1, "16 lanes per chirp need d == 4");
    using Model = ModelSDE<1>;
    constexpr int D = 4, DD = 16, TB = 16, PROW = 22;      // [mf (4) | X (16)] + pad: 11 x 16 bytes per row
    __shared__ __align__(16) double pre[2][TB][PROW];
    const int lane = threadIdx.x;
    const HalfWarp hw(lane);
    const int half = lane >> 4;
    const int64_t gid = (int64_t)blockIdx.x * 2 + half;
    const bool active = gid < p.B;
    const int64_t b = active ? gid : p.B - 1;
    const int64_t T = p.T;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride);
    const ChirpJRow jc(mdl, hw);
    double Qf[D][D];
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) {
        const int rr = r > c ? r : c, cc = r > c ? c : r;
        Qf[r][c] = (p.Qc + b * p.Qc_stride)[rr * D + cc];
    }
    const int ii = hw.i > hw.j ? hw.i : hw.j, jj = hw.i > hw.j ? hw.j : hw.i;
    const double Qe = (p.Qc + b * p.Qc_stride)[ii * D + jj];
    const double *__restrict__ mfs = io.mfs + b * T * D;
    const double *__restrict__ Pfs = io.Pfs + b * T * DD;
    double ms[D];
    load_vec<D>(mfs + (T - 1) * D, ms);
    double Pe = Pfs[(T - 1) * DD + ii * D + jj];
    const bool store_m = active && hw.l < D;
    if (active) {
        io.Pss[(b * T + T - 1) * DD + hw.l] = Pe;
        if (store_m) io.mss[(b * T + T - 1) * D + hw.l] = mfs[(T - 1) * D + hw.l];
    }
    const double ndt = -p.dt;
    // steps T-2 ... 0 in 16-aligned blocks, walked downwards; per block lane j prepares step lo + j (SIMD over time), the
    // 32-bit inner loop holds the chain and nothing else
    for (int64_t hi = T - 1; hi > 0;) {                    // steps [lo, hi) of this block
        const int64_t lo = (hi - 1) & ~(int64_t)(TB - 1);
        const int n = (int)(hi - lo);
        if (hw.l < n) {
            const int64_t tj = lo + hw.l;
            double mfj[D], Pf[D][D], Lf[D][D], rinv[D];
            load_vec<D>(mfs + tj * D, mfj);
            load_mat<D>(Pfs + tj * DD, Pf);
            chol_lower_rsqrt<D>(Pf, Lf, rinv);
            store_vec<D>(&pre[half][hw.l][0], mfj);
            CGP_UNROLL for (int c = 0; c < D; c++) {       // column c of X = Pf^{-1} gamma
                double col[D];
                CGP_UNROLL for (int q = 0; q < D; q++) col[q] = Qf[q][c];
                chol_solve_vec_rinv<D>(Lf, rinv, col);
                store_vec<D>(&pre[half][hw.l][D + c * D], col);
            }
        }
        __syncwarp();
        double *pP = io.Pss + (b * T + hi - 1) * DD + hw.l;
        double *pm = io.mss + (b * T + hi - 1) * D + hw.j;
        for (int s = n - 1; s >= 0; s--) {
            double mf[D], xcol[D];                         // column i of X: row i of X^T = gamma Pf^{-1}, and the constant part of M's row i
            load_vec<D>(&pre[half][s][0], mf);
            load_vec<D>(&pre[half][s][D + hw.i * D], xcol);
            // (a straight-line variant of this step -- side of the softplus split chosen once per step so that two evaluations
            // could be in flight: V of stage s+1 does not depend on w of stage s -- was measured and dropped: ptxas does not
            // overlap them, 5.52 vs 5.43 ms, profiles/r2_cd_lane_kernels.txt)
            rk4_step_lane([&](int, const double (&mm)[D], double PP, double (&dm)[D], double &dP) {
                double jr[D], a[D], z[D], gv, sg;
                softplus_and_sigmoid(mm[2], gv, sg);
                const double w = (kTwoPi * gv) * mdl.fs, dw = (kTwoPi * sg) * mdl.fs;
                mdl.drift_w(w, mm, a);
                jc.row(mm, w, dw, jr);
                CGP_UNROLL for (int q = 0; q < D; q++) jr[q] = jr[q] + xcol[q];
                CGP_UNROLL for (int q = 0; q < D; q++) z[q] = mm[q] - mf[q];
                // dm_r = a_r + sum_k X_kr z_k: lane (i, .) forms component i with its column of X, the four are gathered
                double di = xcol[0] * z[0];
                CGP_UNROLL for (int k = 1; k < D; k++) di = fma(xcol[k], z[k], di);
                CGP_UNROLL for (int r = 0; r < D; r++) dm[r] = a[r] + hw.get(di, 4 * r);
                const double Y = row_times_P(hw, jr, PP);
                const double Yt = hw.get(Y, 4 * hw.j + hw.i);
                dP = (Y + Yt) - Qe;
            }, ms, Pe, ndt);
            if (active) {
                *pP = Pe;
                if (store_m) *pm = jc.by_col(ms);
            }
            pP -= DD;
            pm -= D;
        }
        __syncwarp();
        hi = lo;
    }
}

// ------------------------------------------------------------------------------------------------ smoother sweep, d = 4, 16 lanes per chirp
// Same recursion as smoother_sweep_warp_kernel, for d = 4: lane (i, j) of a half-warp holds Ps_ij, the two small matrix
// products exchange their operands with shuffles (no shared-memory round trips on the dependency chain), and the 16
// entries of a step are written with one coalesced 128-byte store.  Tiles of [G | c | C] records (30 doubles per step) are
// staged with cp.async, NSTAGE deep, per half-warp.
template <int TS_, int NSTAGE>
__global__ void __launch_bounds__(32) smoother_sweep_lane4_kernel(const CgpProblem p, const SmootherIO io) {
    constexpr int D = 4, DD = 16, R = ws_record<4>(), TS = TS_, TILE = TS * R;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x;
    const HalfWarp hw(lane);
    const int half = lane >> 4;
    const int64_t gid = (int64_t)blockIdx.x * 2 + half;
    const bool active = gid < p.B;
    const int64_t b = active ? gid : p.B - 1;
    const int64_t T = p.T;
    double *my = smem + half * (NSTAGE * TILE);
    const double *__restrict__ ws = io.ws + b * T * R;
    const double *__restrict__ mfs = io.mfs + b * T * D;
    const double *__restrict__ Pfs = io.Pfs + b * T * DD;
    double *__restrict__ mss = io.mss + b * T * D;
    double *__restrict__ Pss = io.Pss + b * T * DD;
    const int eC = DD + D + sidx(hw.i, hw.j);             // this lane's entry of the packed C within a record
    double Pe = Pfs[(T - 1) * DD + hw.l];
    double ms[D];
    CGP_UNROLL for (int q = 0; q < D; q++) ms[q] = mfs[(T - 1) * D + q];
    if (active) {
        Pss[(T - 1) * DD + hw.l] = Pe;
        if (hw.l < D) mss[(T - 1) * D + hw.l] = mfs[(T - 1) * D + hw.l];
    }
    if (T < 2) return;
    auto issue_tile = [&](int buf, int64_t lo, int n) {
        double *dst = my + buf * TILE;
        const double *s0 = ws + lo * R;
        for (int i = hw.l; i < n * R / 2; i += 16) cp_async16(dst + 2 * i, s0 + 2 * i);
        cp_async_commit();
    };
    // NSTAGE-deep cp.async pipeline over tiles walking backwards from step T-2; empty commit groups keep the group
    // count uniform at the tail
    int64_t hi = T - 1;                 // steps [lo, hi) of the tile being consumed
    int64_t next_hi = T - 1;            // upper end of the next tile to issue
    int buf = 0, ibuf = 0;
    CGP_UNROLL for (int st = 0; st < NSTAGE - 1; st++) {
        if (next_hi > 0) {
            const int n = (int)(next_hi < TS ? next_hi : TS);
            issue_tile(ibuf, next_hi - n, n);
            next_hi -= n;
        } else {
            cp_async_commit();
        }
        ibuf = (ibuf + 1) % NSTAGE;
    }
    while (hi > 0) {
        const int n = (int)(hi < TS ? hi : TS);
        const int64_t lo = hi - n;
        if (next_hi > 0) {
            const int nn = (int)(next_hi < TS ? next_hi : TS);
            issue_tile(ibuf, next_hi - nn, nn);
            next_hi -= nn;
        } else {
            cp_async_commit();
        }
        ibuf = (ibuf + 1) % NSTAGE;
        cp_async_wait<NSTAGE - 1>();
        __syncwarp();
        const double *tw = my + buf * TILE;
        double *pP = Pss + (lo + n - 1) * DD, *pm = mss + (lo + n - 1) * D;
        #pragma unroll 4
        for (int jj = n - 1; jj >= 0; jj--) {
            const double *Gm = tw + jj * R, *cv = Gm + DD;
            // rows i and j of the gain (each lane needs both)
            double gi[D], gj[D];
            CGP_UNROLL for (int k = 0; k < D; k++) { gi[k] = Gm[hw.i * D + k]; gj[k] = Gm[hw.j * D + k]; }
            double t1 = gi[0] * hw.get(Pe, hw.j);                                  // (G Ps')_ij = sum_k G_ik Ps'_kj
            CGP_UNROLL for (int k = 1; k < D; k++) t1 = fma(gi[k], hw.get(Pe, 4 * k + hw.j), t1);
            double t2 = hw.get(t1, 4 * hw.i) * gj[0];                              // (T1 G^T)_ij = sum_k T1_ik G_jk
            CGP_UNROLL for (int k = 1; k < D; k++) t2 = fma(hw.get(t1, 4 * hw.i + k), gj[k], t2);
            Pe = Gm[eC] + t2;
            // ms = c + G ms': lane (i, .) forms row i, the four rows are then gathered from lanes (k, 0)
            double msi = gi[0] * ms[0];
            CGP_UNROLL for (int k = 1; k < D; k++) msi = fma(gi[k], ms[k], msi);
            msi = cv[hw.i] + msi;
            CGP_UNROLL for (int k = 0; k < D; k++) ms[k] = hw.get(msi, 4 * k);
            if (active) {
                pP[hw.l] = Pe;
                if (hw.j == 0) pm[hw.i] = msi;
            }
            pP -= DD;
            pm -= D;
        }
        __syncwarp();
        hi = lo;
        buf = (buf + 1) % NSTAGE;
    }
}

// ------------------------------------------------------------------------------------------------ smoother sweep, d = 8, one warp per chirp
// Lane (i, jq) = (l / 4, l % 4) holds the two covariance entries (i, jq) and (i, jq + 4); the 8 x 8 products take their
// operands by shuffle: (G X)_ij needs column j of X (lanes (k, j % 4), slot j / 4), (T1 G^T)_ij needs row i of T1 (the
// four lanes of row i, both slots).  Gain rows come from the cp.async-staged tile of [G | c | C] records in shared memory.
template <int TS_, int NSTAGE>
__global__ void __launch_bounds__(32) smoother_sweep_lane8_kernel(const CgpProblem p, const SmootherIO io) {
    constexpr int D = 8, DD = 64, R = ws_record<8>(), TS = TS_, TILE = TS * R;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x, li = lane >> 2, jq = lane & 3;
    const int64_t b = blockIdx.x;
    const int64_t T = p.T;
    const double *__restrict__ ws = io.ws + b * T * R;
    const double *__restrict__ mfs = io.mfs + b * T * D;
    const double *__restrict__ Pfs = io.Pfs + b * T * DD;
    double *__restrict__ mss = io.mss + b * T * D;
    double *__restrict__ Pss = io.Pss + b * T * DD;
    const int ea = li * D + jq, eb = ea + 4;
    const int ca = DD + D + sidx(li, jq), cb = DD + D + sidx(li, jq + 4);      // the two entries of the packed C within a record
    double Pa = Pfs[(T - 1) * DD + ea], Pb = Pfs[(T - 1) * DD + eb];
    double ms[D];
    CGP_UNROLL for (int q = 0; q < D; q++) ms[q] = mfs[(T - 1) * D + q];
    Pss[(T - 1) * DD + ea] = Pa;
    Pss[(T - 1) * DD + eb] = Pb;
    if (lane < D) mss[(T - 1) * D + lane] = mfs[(T - 1) * D + lane];
    if (T < 2) return;
    auto issue_tile = [&](int buf, int64_t lo, int n) {
        double *dst = smem + buf * TILE;
        const double *s0 = ws + lo * R;
        for (int i = lane; i < n * R / 2; i += 32) cp_async16(dst + 2 * i, s0 + 2 * i);
        cp_async_commit();
    };
    int64_t hi = T - 1, next_hi = T - 1;
    int buf = 0, ibuf = 0;
    CGP_UNROLL for (int st = 0; st < NSTAGE - 1; st++) {
        if (next_hi > 0) {
            const int n = (int)(next_hi < TS ? next_hi : TS);
            issue_tile(ibuf, next_hi - n, n);
            next_hi -= n;
        } else {
            cp_async_commit();
        }
        ibuf = (ibuf + 1) % NSTAGE;
    }
    while (hi > 0) {
        const int n = (int)(hi < TS ? hi : TS);
        const int64_t lo = hi - n;
        if (next_hi > 0) {
            const int nn = (int)(next_hi < TS ? next_hi : TS);
            issue_tile(ibuf, next_hi - nn, nn);
            next_hi -= nn;
        } else {
            cp_async_commit();
        }
        ibuf = (ibuf + 1) % NSTAGE;
        cp_async_wait<NSTAGE - 1>();
        __syncwarp();
        const double *tw = smem + buf * TILE;
        double *pP = Pss + (lo + n - 1) * DD, *pm = mss + (lo + n - 1) * D;
        for (int jj = n - 1; jj >= 0; jj--) {
            const double *Gm = tw + jj * R, *cv = Gm + DD;
            double gi[D], ga[D], gb[D];
            CGP_UNROLL for (int k = 0; k < D; k++) { gi[k] = Gm[li * D + k]; ga[k] = Gm[jq * D + k]; gb[k] = Gm[(jq + 4) * D + k]; }
            double ta = gi[0] * __shfl_sync(0xffffffffu, Pa, jq), tb = gi[0] * __shfl_sync(0xffffffffu, Pb, jq);
            CGP_UNROLL for (int k = 1; k < D; k++) {
                ta = fma(gi[k], __shfl_sync(0xffffffffu, Pa, 4 * k + jq), ta);
                tb = fma(gi[k], __shfl_sync(0xffffffffu, Pb, 4 * k + jq), tb);
            }
            double ua = 0., ub = 0.;
            CGP_UNROLL for (int k = 0; k < 4; k++) {                  // T1_ik for k < 4 sits in slot a of lane (i, k)
                const double t1 = __shfl_sync(0xffffffffu, ta, 4 * li + k);
                ua = (k == 0) ? t1 * ga[0] : fma(t1, ga[k], ua);
                ub = (k == 0) ? t1 * gb[0] : fma(t1, gb[k], ub);
            }
            CGP_UNROLL for (int k = 0; k < 4; k++) {                  // k + 4: slot b
                const double t1 = __shfl_sync(0xffffffffu, tb, 4 * li + k);
                ua = fma(t1, ga[k + 4], ua);
                ub = fma(t1, gb[k + 4], ub);
            }
            Pa = Gm[ca] + ua;
            Pb = Gm[cb] + ub;
            double msi = gi[0] * ms[0];
            CGP_UNROLL for (int k = 1; k < D; k++) msi = fma(gi[k], ms[k], msi);
            msi = cv[li] + msi;
            CGP_UNROLL for (int k = 0; k < D; k++) ms[k] = __shfl_sync(0xffffffffu, msi, 4 * k);
            pP[ea] = Pa;
            pP[eb] = Pb;
            if (jq == 0) pm[li] = msi;
            pP -= DD;
            pm -= D;
        }
        __syncwarp();
        hi = lo;
        buf = (buf + 1) % NSTAGE;
    }
}

// ------------------------------------------------------------------------------------------------ cubature smoother gains, thread per (chirp, step)
// sgp_smoother's time-parallel half (filters_smoothers.py:520-527) for the spherical cubature rule (quadratures.py:139-150:
// points m +- sqrt(d) L e_j, equal weights) at larger d, where the generic kernel runs out of registers (d = 8: 36 + 36 + 64
// accumulators).  Uses the structure of the rule:
//   * chi_j+- = m +- s l_j needs only column j of L = chol(Pf): L lives in shared memory, one column is read per pair;
//   * the cross-covariance  D = sum_p w chi_p f_p^T - m mp^T  equals  s w sum_j l_j (f_j+ - f_j-)^T  (the m mp^T terms cancel
//     identically; evaluating this form avoids the reference's cancellation instead of reproducing it -- the difference is
//     the reference's own rounding noise, ~1e-13 relative), so  G = D Pp^{-1} = s w sum_j l_j z_j^T  with  Pp z_j = f_j+ - f_j-.
// D is accumulated in shared memory (rank-one updates l_j (f_j+ - f_j-)^T touch only rows r >= j), the rows of G are then
// solved in registers against chol(Pp).
template <int NH>
__global__ void __launch_bounds__(64) cubature_gain_kernel(const CgpProblem p, const SmootherIO io) {
    using Model = ModelLCD<NH>;
    constexpr int D = Model::D, NS = NSym<D>::value, BLK = 64;
    extern __shared__ __align__(16) double cg_smem[];
    double (*Ls)[BLK] = reinterpret_cast<double (*)[BLK]>(cg_smem);                    // chol(Pf), element-major: conflict-free
    double (*Ds)[BLK] = reinterpret_cast<double (*)[BLK]>(cg_smem + NS * BLK);         // s w sum_j l_j (f_j+ - f_j-)^T
    const int tid = threadIdx.x;
    const int64_t Tm1 = p.T - 1;
    const int64_t item = (int64_t)blockIdx.x * BLK + tid;
    if (item >= p.B * Tm1) return;
    const int64_t b = item / Tm1, t = item - b * Tm1;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    double m[D];
    load_vec<D>(io.mfs + (b * p.T + t) * D, m);
    double *__restrict__ rec = io.ws + (b * p.T + t) * ws_record<D>();      // [G | c | C]
    {
        double Pf[NS], L[NS];
        load_sym<D>(io.Pfs + (b * p.T + t) * (D * D), Pf);
        chol_lower_sym_rsqrt<D>(Pf, L);
        CGP_UNROLL for (int i = 0; i < NS; i++) Ls[i][tid] = L[i];
    }
    CGP_UNROLL for (int i = 0; i < D * D; i++) Ds[i][tid] = 0.;
    const double w = __ldg(p.sig_w);                         // equal weights 1 / (2 d)
    double am[D], aP[NS];
    CGP_UNROLL for (int i = 0; i < D; i++) am[i] = 0.;
    CGP_UNROLL for (int i = 0; i < NS; i++) aP[i] = 0.;
    auto accumulate = [&](const double (&ev)[D]) {
        CGP_UNROLL for (int r = 0; r < D; r++) am[r] = fma(w, ev[r], am[r]);
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++) {
            double v = ev[r] * ev[c];
            if (Model::has_sig(r, c)) v += mdl.sig(r, c);
            aP[sidx(r, c)] = fma(w, v, aP[sidx(r, c)]);
        }
    };
    CGP_UNROLL for (int j = 0; j < D; j++) {
        const double sp = __ldg(p.sig_xi + j * D + j), sn = __ldg(p.sig_xi + (D + j) * D + j);   // +sqrt(d), -sqrt(d)
        double lj[D], chi[D], evp[D], evn[D];
        CGP_UNROLL for (int r = j; r < D; r++) lj[r] = Ls[sidx(r, j)][tid];
        CGP_UNROLL for (int r = 0; r < D; r++) chi[r] = (r >= j) ? m[r] + lj[r] * sp : m[r];
        mdl.mean(chi, evp);
        accumulate(evp);
        CGP_UNROLL for (int r = 0; r < D; r++) chi[r] = (r >= j) ? m[r] + lj[r] * sn : m[r];
        mdl.mean(chi, evn);
        accumulate(evn);
        const double sw = sp * w;
        CGP_UNROLL for (int c = 0; c < D; c++) evp[c] = (evp[c] - evn[c]) * sw;
        CGP_UNROLL for (int r = j; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++)
            Ds[r * D + c][tid] = fma(lj[r], evp[c], Ds[r * D + c][tid]);
    }
    // chol(Pp) into the registers the accumulators leave behind
    double Lq[NS], rinv[D];
    {
        double Pps[NS];
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++) Pps[sidx(r, c)] = aP[sidx(r, c)] - am[r] * am[c];
        CGP_UNROLL for (int j = 0; j < D; j++) {             // chol_lower_sym_rsqrt, keeping 1 / L_jj
            double sacc = Pps[sidx(j, j)];
            CGP_UNROLL for (int k = 0; k < j; k++) sacc = fma(-Lq[sidx(j, k)], Lq[sidx(j, k)], sacc);
            const double r = fast_rsqrt(sacc);
            rinv[j] = r;
            Lq[sidx(j, j)] = sacc * r;
            CGP_UNROLL for (int i = j + 1; i < D; i++) {
                double tacc = Pps[sidx(i, j)];
                CGP_UNROLL for (int k = 0; k < j; k++) tacc = fma(-Lq[sidx(i, k)], Lq[sidx(j, k)], tacc);
                Lq[sidx(i, j)] = tacc * r;
            }
        }
    }
    // G = D Pp^{-1}: row r of G solves Pp g = (row r of D)^T;  c = m - G mp;  C = Pf - G D^T goes through the shared-memory
    // columns chol(Pf) has left (Ls), so that it can leave in whole sectors
    const double *__restrict__ Pfg = io.Pfs + (b * p.T + t) * (D * D);
    double cv[D];
    CGP_UNROLL for (int r = 0; r < D; r++) {
        double z[D];
        CGP_UNROLL for (int c = 0; c < D; c++) z[c] = Ds[r * D + c][tid];
        CGP_UNROLL for (int i = 0; i < D; i++) {
            double sacc = z[i];
            CGP_UNROLL for (int k = 0; k < i; k++) sacc = fma(-Lq[sidx(i, k)], z[k], sacc);
            z[i] = sacc * rinv[i];
        }
        CGP_UNROLL for (int i = D - 1; i >= 0; i--) {
            double sacc = z[i];
            CGP_UNROLL for (int k = i + 1; k < D; k++) sacc = fma(-Lq[sidx(k, i)], z[k], sacc);
            z[i] = sacc * rinv[i];
        }
        gstore_vec_auto<D>(rec + r * D, z);
        double sacc = m[r];
        CGP_UNROLL for (int k = 0; k < D; k++) sacc = fma(-z[k], am[k], sacc);
        cv[r] = sacc;
        CGP_UNROLL for (int q = 0; q <= r; q++) {
            double cacc = __ldg(Pfg + r * D + q);
            CGP_UNROLL for (int k = 0; k < D; k++) cacc = fma(-z[k], Ds[q * D + k][tid], cacc);
            Ls[sidx(r, q)][tid] = cacc;
        }
    }
    gstore_vec_auto<D>(rec + D * D, cv);
    {
        constexpr int NSP = ws_record<D>() - D * D - D;      // packed C + padding
        double Cs[NSP];
        CGP_UNROLL for (int i = 0; i < NSP; i++) Cs[i] = i < NS ? Ls[i < NS ? i : 0][tid] : 0.;
        gstore_vec_auto<NSP>(rec + D * D + D, Cs);
    }
}

}  // namespace cgp
