// cgp_fast.cuh -- tuned kernels for the headline path (Gauss-Hermite sigma-point filter + smoother on the
// chirp-family LCD models).  Same arithmetic as the generic kernels in cgp_kernels.cuh up to rounding:
//   * Cholesky columns are scaled by rsqrt(pivot) instead of divided by sqrt(pivot)      (<= 2 ulp per entry)
//   * the Kalman gain uses one reciprocal of S instead of d divisions                     (<= 1 ulp per entry)
//   * sigma-point partial sums are combined through shared memory in a fixed tree order instead of a
//     shuffle butterfly (a different, still deterministic, summation order).
#pragma once
#include "cgp_kernels.cuh"

namespace cgp {

// measurement update on packed-symmetric covariance with one reciprocal; H generic or the unit vector e_1
template <int D, bool H_E1>
CGP_DEV double linear_update_fast(const double (&mp)[D], const double (&Pp)[NSym<D>::value], const double (&H)[D], double Xi,
                                  double y, double (&mf)[D], double (&Pf)[NSym<D>::value]) {
    double PH[D], S, pred;
    if constexpr (H_E1) {
        CGP_UNROLL for (int i = 0; i < D; i++) PH[i] = Pp[sidx(i, 1)];
        S = PH[1] + Xi;
        pred = mp[1];
    } else {
        CGP_UNROLL for (int i = 0; i < D; i++) {
            double s = Pp[sidx(i, 0)] * H[0];
            CGP_UNROLL for (int j = 1; j < D; j++) s = fma(Pp[sidx(i, j)], H[j], s);
            PH[i] = s;
        }
        S = PH[0] * H[0];
        CGP_UNROLL for (int j = 1; j < D; j++) S = fma(PH[j], H[j], S);
        S += Xi;
        pred = H[0] * mp[0];
        CGP_UNROLL for (int i = 1; i < D; i++) pred = fma(H[i], mp[i], pred);
    }
    const double rS = 1. / S;
    double K[D];
    CGP_UNROLL for (int i = 0; i < D; i++) K[i] = PH[i] * rS;
    const double r = y - pred;
    CGP_UNROLL for (int i = 0; i < D; i++) mf[i] = fma(K[i], r, mp[i]);
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j <= i; j++)
        Pf[sidx(i, j)] = fma(-(K[i] * K[j]), S, Pp[sidx(i, j)]);
    const double sc = sqrt(S), sc2 = sc * sc;
    return (log(kTwoPi * sc2) + r * r / sc2) * 0.5;
}

// ------------------------------------------------------------------------------------------------ GH filter, warp per chirp
// sgp_filter (filters_smoothers.py:446-490) for ModelLCD<NH> with a Gauss-Hermite table of P nodes per dimension
// whose P^(D-1) base indices fit one warp.  Lane `l` owns base index l: its P points (l + c * nb) share
// chi[0..D-2] and the transcendental part of the model; all table entries the lane needs sit in registers.
template <int NH, int P>
__global__ void __launch_bounds__(128) ghf_filter_kernel(const CgpProblem p, const FilterIO io) {
    using Model = ModelLCD<NH>;
    constexpr int D = Model::D, V = Model::V, NS = NSym<D>::value, NA = D + NS;
    constexpr int PITCH = 33;
    constexpr int KP = (NA <= 16) ? 16 : 32;          // lanes per "half" in the shared-memory reduction
    constexpr int HS = 32 / KP;                       // halves: each sums 32 / HS partials
    constexpr int WARPS = 4;
    __shared__ double red[WARPS][NA][PITCH];
    __shared__ __align__(16) double res[WARPS][(NA + 1) & ~1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gid = (int64_t)blockIdx.x * WARPS + warp;
    const bool active = gid < p.B;
    const int64_t b = active ? gid : p.B - 1;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    double m[D], Pc[NS], H[D];
    load_vec<D>(p.m0 + b * p.m0_stride, m);
    load_sym<D>(p.P0 + b * p.P0_stride, Pc);
    bool h_e1 = true;
    CGP_UNROLL for (int i = 0; i < D; i++) { H[i] = p.H[i]; h_e1 = h_e1 && (H[i] == (i == 1 ? 1. : 0.)); }
    // per-lane table entries
    int nb = 1;
    CGP_UNROLL for (int i = 0; i < D - 1; i++) nb *= P;
    const bool has_pts = lane < nb;
    double xb[D - 1], wl[P], xlast[P];
    CGP_UNROLL for (int r = 0; r < D - 1; r++) xb[r] = has_pts ? p.sig_xi[lane * D + r] : 0.;
    CGP_UNROLL for (int c = 0; c < P; c++) {
        wl[c] = has_pts ? p.sig_w[lane + c * nb] : 0.;
        xlast[c] = p.sig_xi[(c * nb) * D + (D - 1)];
    }
    const double *__restrict__ y = io.ys + (b / p.ys_repeat) * p.T;
    const int64_t T = p.T;
    const bool store = io.mfs != nullptr && active && lane == 0;
    const bool store_nell = io.nell != nullptr && active && lane == 0;
    double acc_nll = 0.;
    double ynext = __ldg(y);
    for (int64_t t = 0; t < T; t++) {
        const double yt = ynext;
        if (t + 1 < T) ynext = __ldg(y + t + 1);
        // ---- sigma points of this lane
        double L[NS];
        chol_lower_sym_rsqrt<D>(Pc, L);
        double chi[D];
        CGP_UNROLL for (int r = 0; r < D - 1; r++) {
            double s = L[sidx(r, 0)] * xb[0];
            CGP_UNROLL for (int c = 1; c <= r; c++) s = fma(L[sidx(r, c)], xb[c], s);
            chi[r] = m[r] + s;
        }
        double slast = L[sidx(D - 1, 0)] * xb[0];
        CGP_UNROLL for (int c = 1; c < D - 1; c++) slast = fma(L[sidx(D - 1, c)], xb[c], slast);
        const typename Model::Trig trig = mdl.prep_v(chi[V]);
        double a[NA];
        CGP_UNROLL for (int i = 0; i < NA; i++) a[i] = 0.;
        CGP_UNROLL for (int c = 0; c < P; c++) {
            chi[D - 1] = m[D - 1] + fma(L[sidx(D - 1, D - 1)], xlast[c], slast);
            double ev[D];
            mdl.mean_with(trig, chi, ev);
            const double w = wl[c];
            CGP_UNROLL for (int r = 0; r < D; r++) a[r] = fma(w, ev[r], a[r]);
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int q = 0; q <= r; q++) {
                double v = ev[r] * ev[q];
                if (Model::has_sig(r, q)) v += mdl.sig(r, q);
                a[D + sidx(r, q)] = fma(w, v, a[D + sidx(r, q)]);
            }
        }
        // ---- combine the 32 lanes' partial sums through shared memory (fixed tree order)
        CGP_UNROLL for (int k = 0; k < NA; k++) red[warp][k][lane] = a[k];
        __syncwarp();
        double tot[NA];
        {
            const int k = lane % KP, h = lane / KP;
            constexpr int CNT = 32 / HS;
            CGP_UNROLL for (int k0 = 0; k0 < NA; k0 += KP) {
                const int kk = k0 + k;
                double v[CNT];
                const bool ok = kk < NA;
                CGP_UNROLL for (int j = 0; j < CNT; j++) v[j] = ok ? red[warp][ok ? kk : 0][h * CNT + j] : 0.;
                CGP_UNROLL for (int w2 = 1; w2 < CNT; w2 <<= 1)
                    CGP_UNROLL for (int j = 0; j + w2 < CNT; j += 2 * w2) v[j] += v[j + w2];
                double s = v[0];
                if (HS == 2) s += __shfl_xor_sync(0xffffffffu, s, 16);
                if (ok && h == 0) res[warp][kk] = s;
            }
        }
        __syncwarp();
        CGP_UNROLL for (int k = 0; k < NA; k += 2) {
            if (k + 1 < NA) {
                const double2 v = *reinterpret_cast<const double2 *>(&res[warp][k]);
                tot[k] = v.x; tot[k + 1] = v.y;
            } else {
                tot[k] = res[warp][k];
            }
        }
        double mp[D], Pp[NS];
        CGP_UNROLL for (int r = 0; r < D; r++) mp[r] = tot[r];
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int q = 0; q <= r; q++)
            Pp[sidx(r, q)] = fma(-mp[r], mp[q], tot[D + sidx(r, q)]);
        // ---- measurement update (filters_smoothers.py:55-68)
        double inc;
        if (h_e1) inc = linear_update_fast<D, true>(mp, Pp, H, p.Xi, yt, m, Pc);
        else inc = linear_update_fast<D, false>(mp, Pp, H, p.Xi, yt, m, Pc);
        acc_nll = acc_nll + inc;
        if (store) {
            store_vec<D>(io.mfs + (b * T + t) * D, m);
            store_sym<D>(io.Pfs + (b * T + t) * (D * D), Pc);
        }
        if (store_nell && !io.nell_last_only) io.nell[b * T + t] = acc_nll;
    }
    if (store_nell && io.nell_last_only) io.nell[b] = acc_nll;
}

// ------------------------------------------------------------------------------------------------ smoother sweep, warp per chirp
// Sequential part of rts / eks / sgp_smoother (filters_smoothers.py:83-84):
//     ms = mf + G (ms - mp),   Ps = Pf + G (Ps - Pp) G^T        for k = T-2 .. 0.
// One warp owns one chirp.  Tiles of TS consecutive steps ([G | mp | Pp] records of the gain kernel plus mf, Pf)
// are staged in shared memory with cp.async (double buffered), the recursion runs with ONE MATRIX ENTRY PER LANE
// (operands exchanged through shared memory), and the TS results are written back with coalesced 16-byte stores.
template <int D> struct SweepCfg {
    static constexpr int R = 2 * D * D + D;                   // workspace record
    static constexpr int TS = (D <= 4) ? 16 : 8;
    static constexpr int OUT = D + D * D;                     // ms | Ps per step
    static constexpr int TILE_DOUBLES = TS * (R + D + D * D); // ws + mf + Pf
    static constexpr int WARPS = 1;
    static constexpr size_t smem_bytes() {
        return sizeof(double) * WARPS * (2 * TILE_DOUBLES + TS * OUT + 2 * D * D + 2 * D);
    }
};

CGP_DEV void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
CGP_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> CGP_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int D>
__global__ void __launch_bounds__(32 * SweepCfg<D>::WARPS) smoother_sweep_warp_kernel(const CgpProblem p, const SmootherIO io) {
    using Cfg = SweepCfg<D>;
    constexpr int R = Cfg::R, TS = Cfg::TS, DD = D * D, OUT = Cfg::OUT, TILE = Cfg::TILE_DOUBLES;
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * Cfg::WARPS + warp;
    if (b >= p.B) return;                                  // whole warp exits together
    double *base = smem + (size_t)warp * (2 * TILE + TS * OUT + 2 * DD + 2 * D);
    double *tile[2] = {base, base + TILE};
    double *outb = base + 2 * TILE;                        // [TS][OUT]
    double *Xs = outb + TS * OUT;                          // Ps - Pp      (DD)
    double *T1 = Xs + DD;                                  // G (Ps - Pp)  (DD)
    double *ms_s = T1 + DD;                                // ms           (D)
    double *dm_s = ms_s + D;                               // ms - mp      (D)
    const int64_t T = p.T;
    const double *__restrict__ ws = io.ws + b * T * R;
    const double *__restrict__ mfs = io.mfs + b * T * D;
    const double *__restrict__ Pfs = io.Pfs + b * T * DD;
    double *__restrict__ mss = io.mss + b * T * D;
    double *__restrict__ Pss = io.Pss + b * T * DD;

    // last step: copy the filter result (filters_smoothers.py:140-142)
    for (int i = lane; i < D; i += 32) { const double v = mfs[(T - 1) * D + i]; mss[(T - 1) * D + i] = v; ms_s[i] = v; }
    for (int i = lane; i < DD; i += 32) { const double v = Pfs[(T - 1) * DD + i]; Pss[(T - 1) * DD + i] = v; T1[i] = v; }
    __syncwarp();
    if (T < 2) return;
    // tiles cover steps [lo, hi) going backwards from T-1 (exclusive)
    auto issue_tile = [&](int buf, int64_t lo, int n) {
        double *dst = tile[buf];
        const double *s0 = ws + lo * R;
        for (int i = lane; i < n * R / 2; i += 32) cp_async16(dst + 2 * i, s0 + 2 * i);
        const double *s1 = mfs + lo * D;
        double *d1 = dst + TS * R;
        for (int i = lane; i < n * D / 2; i += 32) cp_async16(d1 + 2 * i, s1 + 2 * i);
        const double *s2 = Pfs + lo * DD;
        double *d2 = d1 + TS * D;
        for (int i = lane; i < n * DD / 2; i += 32) cp_async16(d2 + 2 * i, s2 + 2 * i);
        cp_async_commit();
    };
    static_assert((R % 2 == 0) || (SweepCfg<D>::TS % 2 == 0), "16-byte copies need even element counts");
    int64_t hi = T - 1;
    int buf = 0;
    {
        const int n = (int)(hi < TS ? hi : TS);
        issue_tile(0, hi - n, n);
    }
    // current Ps lives in T1-slot "Pcur" registers: entry e of lane (e = lane, lane + 32, ...)
    constexpr int EPL = (DD + 31) / 32;                    // entries per lane
    double Pcur[EPL], mcur = 0.;
    CGP_UNROLL for (int q = 0; q < EPL; q++) { const int e = lane + 32 * q; Pcur[q] = e < DD ? T1[e] : 0.; }
    if (lane < D) mcur = ms_s[lane];
    while (hi > 0) {
        const int n = (int)(hi < TS ? hi : TS);
        const int64_t lo = hi - n;
        const int64_t nhi = lo;
        if (nhi > 0) {
            const int nn = (int)(nhi < TS ? nhi : TS);
            issue_tile(buf ^ 1, nhi - nn, nn);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const double *tw = tile[buf];
        const double *tm = tw + TS * R;
        const double *tP = tm + TS * D;
        for (int j = n - 1; j >= 0; j--) {
            const double *rec = tw + j * R;                // [G | mp | Pp]
            const double *Gm = rec, *mp = rec + DD, *Pp = rec + DD + D;
            // X = Ps - Pp ; dm = ms - mp
            CGP_UNROLL for (int q = 0; q < EPL; q++) { const int e = lane + 32 * q; if (e < DD) Xs[e] = Pcur[q] - Pp[e]; }
            if (lane < D) dm_s[lane] = mcur - mp[lane];
            __syncwarp();
            // T1 = G X ; ms = mf + G dm
            CGP_UNROLL for (int q = 0; q < EPL; q++) {
                const int e = lane + 32 * q;
                if (e < DD) {
                    const int r = e / D, c = e % D;
                    double s = Gm[r * D] * Xs[c];
                    CGP_UNROLL for (int k = 1; k < D; k++) s = fma(Gm[r * D + k], Xs[k * D + c], s);
                    T1[e] = s;
                }
            }
            if (lane < D) {
                double s = Gm[lane * D] * dm_s[0];
                CGP_UNROLL for (int k = 1; k < D; k++) s = fma(Gm[lane * D + k], dm_s[k], s);
                mcur = tm[j * D + lane] + s;
                outb[j * OUT + lane] = mcur;
            }
            __syncwarp();
            // Ps = Pf + T1 G^T
            CGP_UNROLL for (int q = 0; q < EPL; q++) {
                const int e = lane + 32 * q;
                if (e < DD) {
                    const int r = e / D, c = e % D;
                    double s = T1[r * D] * Gm[c * D];
                    CGP_UNROLL for (int k = 1; k < D; k++) s = fma(T1[r * D + k], Gm[c * D + k], s);
                    Pcur[q] = tP[j * DD + e] + s;
                    outb[j * OUT + D + e] = Pcur[q];
                }
            }
            __syncwarp();
        }
        // coalesced write-back of the n finished steps
        for (int i = lane; i < n * D; i += 32) mss[lo * D + i] = outb[(i / D) * OUT + (i % D)];
        for (int i = lane; i < n * DD; i += 32) Pss[lo * DD + i] = outb[(i / DD) * OUT + D + (i % DD)];
        __syncwarp();
        hi = lo;
        buf ^= 1;
    }
}

}  // namespace cgp
