// cgp_multi.cuh -- EXPERIMENT (profiles/r2_multi_chirp.txt): Gauss-Hermite sigma-point filter with NCH chirps per warp,
// interleaved in ONE instruction stream.  At 1000 chirps the warp-per-chirp kernels are bound by the dependency chain of a
// single warp while half of the issue slots idle, and 1000 chain warps on 592 SM sub-partitions leave 408 of them with two
// chains.  Here every lane carries the replicated state of NCH chirps and each phase of the step (Cholesky + sigma points +
// model + partial sums | shared-memory reduction | totals + measurement update) is written for all NCH chirps at once, so
// that the independent chains sit in the same basic block and the compiler / scoreboard can overlap them.  nll-only.
#pragma once
#include "cgp_fast.cuh"

namespace cgp {

template <int NCH, bool H_E1>
__global__ void __launch_bounds__(32) gh_warp_multi_nll_kernel(const CgpProblem p, const double *__restrict__ ys, double *__restrict__ nell_last) {
    using Pred = GhPredictLCD<1, 3>;
    using Model = Pred::Model;
    constexpr int D = Pred::D, V = Pred::V, NS = Pred::NS, NA = Pred::NA, P = 3;
    __shared__ double red[NCH][NA][33];
    __shared__ __align__(16) double res[NCH][(NA + 1) & ~1];
    __shared__ double nl[NCH][32];
    const int lane = threadIdx.x;
    const int64_t T = p.T;
    Pred pred[NCH];
    double m[NCH][D], Pc[NCH][NS], H[D], carry[NCH], Sk[NCH], rk[NCH], yv[NCH];
    const double *__restrict__ y[NCH];
    int64_t bs[NCH];
    CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
    CGP_UNROLL for (int c = 0; c < NCH; c++) {
        int64_t b = (int64_t)blockIdx.x * NCH + c;
        if (b >= p.B) b = p.B - 1;                          // surplus slot of the last warp: recomputes the last chirp
        bs[c] = b;
        pred[c].load(p, b, lane);
        load_vec<D>(p.m0 + b * p.m0_stride, m[c]);
        load_sym<D>(p.P0 + b * p.P0_stride, Pc[c]);
        y[c] = ys + (b / p.ys_repeat) * T;
        yv[c] = (lane < T) ? __ldg(y[c] + lane) : 0.;
        carry[c] = 0.; Sk[c] = 1.; rk[c] = 0.;
    }
    for (int64_t t = 0; t < T; t++) {
        const int slot = (int)(t & 31);
        double yt[NCH];
        CGP_UNROLL for (int c = 0; c < NCH; c++) {
            yt[c] = __shfl_sync(0xffffffffu, yv[c], slot);
            if (slot == 31 && t + 1 < T) yv[c] = (t + 1 + lane < T) ? __ldg(y[c] + t + 1 + lane) : 0.;
        }
        // ---- phase A: per-lane partial sums of every chirp (independent chains, one basic block)
        CGP_UNROLL for (int c = 0; c < NCH; c++) {
            const Model &mdl = pred[c].mdl;
            const GhLane<D, P> &tab = pred[c].tab;
            double L[NS];
            chol_lower_sym_rsqrt<D>(Pc[c], L);
            double chi[D], slast;
            tab.points(m[c], L, chi, slast);
            const typename Model::Trig trig = mdl.template prep_v<true>(chi[V]);
            double ev[D], S0 = 0., S1 = 0., q00 = 0., q10 = 0., q11 = 0.;
            CGP_UNROLL for (int k = 0; k < P; k++) {
                chi[D - 1] = m[c][D - 1] + fma(L[sidx(D - 1, D - 1)], tab.xlast[k], slast);
                if (k == 0) mdl.mean_with(trig, chi, ev); else mdl.mean_tail(chi, ev);
                const double w = tab.wl[k];
                S0 = fma(w, ev[V], S0);
                S1 = fma(w, ev[V + 1], S1);
                q00 = fma(w, ev[V] * ev[V] + mdl.sig(V, V), q00);
                q10 = fma(w, ev[V + 1] * ev[V] + mdl.sig(V + 1, V), q10);
                q11 = fma(w, ev[V + 1] * ev[V + 1] + mdl.sig(V + 1, V + 1), q11);
            }
            double a[NA];
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
            CGP_UNROLL for (int k = 0; k < NA; k++) red[c][k][lane] = a[k];
        }
        __syncwarp();
        // ---- phase B: the 14 sums of every chirp (two lanes per sum, as warp_sum_smem)
        {
            const int k = lane % 16, h = lane / 16;
            CGP_UNROLL for (int c = 0; c < NCH; c++) {
                const bool ok = k < NA;
                double v[16];
                CGP_UNROLL for (int j = 0; j < 16; j++) v[j] = red[c][ok ? k : 0][h * 16 + j];
                CGP_UNROLL for (int w2 = 1; w2 < 16; w2 <<= 1)
                    CGP_UNROLL for (int j = 0; j + w2 < 16; j += 2 * w2) v[j] += v[j + w2];
                double sacc = v[0];
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 16);
                if (ok && h == 0) res[c][k] = sacc;
            }
        }
        __syncwarp();
        // ---- phase C: totals, measurement update
        CGP_UNROLL for (int c = 0; c < NCH; c++) {
            double tot[NA], mp[D], Pp[NS];
            CGP_UNROLL for (int k2 = 0; k2 < NA; k2 += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(&res[c][k2]);
                tot[k2] = v.x; tot[k2 + 1] = v.y;
            }
            CGP_UNROLL for (int r = 0; r < D; r++) mp[r] = tot[r];
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int q = 0; q <= r; q++)
                Pp[sidx(r, q)] = fma(-mp[r], mp[q], tot[D + sidx(r, q)]);
            double S, resid;
            linear_update_fast<D, H_E1>(mp, Pp, H, p.Xi, yt[c], m[c], Pc[c], S, resid);
            if (lane == slot) { Sk[c] = S; rk[c] = resid; }
        }
        if (slot == 31 || t == T - 1) {
            const int n = slot + 1;
            CGP_UNROLL for (int c = 0; c < NCH; c++) nl[c][lane] = lane < n ? nll_increment(Sk[c], rk[c]) : 0.;
            __syncwarp();
            if (lane < NCH) {
                double cc = 0.;
                CGP_UNROLL for (int c = 0; c < NCH; c++) if (lane == c) cc = carry[c];
                for (int j = 0; j < n; j++) { cc = cc + nl[lane][j]; nl[lane][j] = cc; }
            }
            __syncwarp();
            CGP_UNROLL for (int c = 0; c < NCH; c++) carry[c] = nl[c][n - 1];
            __syncwarp();
        }
    }
    if (lane == 0) {
        CGP_UNROLL for (int c = 0; c < NCH; c++)
            if ((int64_t)blockIdx.x * NCH + c < p.B) nell_last[bs[c]] = carry[c];
    }
}

}  // namespace cgp

namespace cgp {

// ---- EXPERIMENT 2: HALF a warp per chirp, two base indices per lane.  The replicated part of a step (Cholesky, moments ->
// (mp, Pp), measurement update) is the same instruction stream for the two chirps of a warp, and 1000 chirps are 500 chain warps:
// at most one per SM sub-partition.  Lane g of a half owns base indices g and g + 16 (< 27) of the 3^3 index prefixes; their
// partial sums go to separate slots and are added in the order of warp_sum_smem: results are bit-identical to the plain kernel.
template <int NH> struct GhHalf {
    using Model = ModelLCD<NH>;
    static constexpr int D = Model::D, V = Model::V, NS = NSym<D>::value, NA = D + NS, P = 3, RP = 34;
    // partial sums a[NA] of one base index (GhPredictLCD::predict_impl) from chi[0..D-2], slast and the e-scaled rotation
    static CGP_DEV void partials(const Model &mdl, const GhLane<D, P> &tab, const typename Model::Trig &trig, const double (&m)[D],
                                 const double (&L)[NS], double (&chi)[D], double slast, double (&a)[NA], double (&ev)[D]) {
        double S0 = 0., S1 = 0., q00 = 0., q10 = 0., q11 = 0.;
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
    }
    // One prediction for the chirp of this half-warp.  red: [NA][RP] of this half, res: [16] of this half.  xop (optional):
    // [V][33] of this half, receives ev[0..V-1] of every base index.
    template <bool EXPORT>
    static CGP_DEV void predict(const Model &mdl, const GhLane<D, P> (&tab)[2], double (*red)[RP], double *res, double (*xop)[33],
                                int lane, const double (&m)[D], const double (&Pc)[NS], double (&mp)[D], double (&Pp)[NS]) {
        const int g = lane & 15;
        double L[NS];
        chol_lower_sym_rsqrt<D>(Pc, L);
        double chi[2][D], slast[2];
        CGP_UNROLL for (int e = 0; e < 2; e++) tab[e].points(m, L, chi[e], slast[e]);
        // softplus branch: as fast_softplus_warp of the plain kernel, decided by all 32 slots (27 base indices + 5 idle ones
        // sitting at chi = m) of THIS chirp
        const bool okl = chi[0][V] >= 3. && chi[0][V] <= 700. && chi[1][V] >= 3. && chi[1][V] <= 700.;
        const unsigned bal = __ballot_sync(0xffffffffu, okl);
        const bool series = ((bal >> (lane & 16)) & 0xffffu) == 0xffffu;
        double gv[2];
        if (series) { CGP_UNROLL for (int e = 0; e < 2; e++) gv[e] = softplus_series(chi[e][V]); }
        else { CGP_UNROLL for (int e = 0; e < 2; e++) gv[e] = softplus_general(chi[e][V]); }
        __syncwarp();
        CGP_UNROLL for (int e = 0; e < 2; e++) {
            const typename Model::Trig trig = mdl.prep_g(gv[e]);
            double a[NA], ev[D];
            partials(mdl, tab[e], trig, m, L, chi[e], slast[e], a, ev);
            CGP_UNROLL for (int k = 0; k < NA; k++) red[k][g + 16 * e] = a[k];
            if constexpr (EXPORT) { CGP_UNROLL for (int q = 0; q < V; q++) xop[q][g + 16 * e] = ev[q]; }
        }
        __syncwarp();
        {
            const bool ok = g < NA;
            const double *row = &red[ok ? g : 0][0];
            double v[32];
            CGP_UNROLL for (int j = 0; j < 32; j += 2) {
                const double2 x = *reinterpret_cast<const double2 *>(row + j);
                v[j] = x.x; v[j + 1] = x.y;
            }
            CGP_UNROLL for (int hh = 0; hh < 2; hh++)
                CGP_UNROLL for (int w2 = 1; w2 < 16; w2 <<= 1)
                    CGP_UNROLL for (int j = 0; j + w2 < 16; j += 2 * w2) v[16 * hh + j] += v[16 * hh + j + w2];
            if (ok) res[g] = v[0] + v[16];
        }
        __syncwarp();
        double tot[NA];
        CGP_UNROLL for (int k2 = 0; k2 < NA; k2 += 2) {
            const double2 x = *reinterpret_cast<const double2 *>(&res[k2]);
            tot[k2] = x.x; tot[k2 + 1] = x.y;
        }
        CGP_UNROLL for (int r = 0; r < D; r++) mp[r] = tot[r];
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int q = 0; q <= r; q++)
            Pp[sidx(r, q)] = fma(-mp[r], mp[q], tot[D + sidx(r, q)]);
    }
};

template <bool H_E1>
__global__ void __launch_bounds__(32) gh_half_nll_kernel(const CgpProblem p, const double *__restrict__ ys, double *__restrict__ nell_last) {
    using GH = GhHalf<1>;
    using Model = GH::Model;
    constexpr int D = GH::D, NS = GH::NS, NA = GH::NA;
    static_assert(NA % 2 == 0, "16-byte reads of the totals");
    __shared__ __align__(16) double red[2][NA][GH::RP];
    __shared__ __align__(16) double res[2][16];
    __shared__ double nl[2][32];
    const int lane = threadIdx.x, h = lane >> 4, g = lane & 15;
    const int64_t T = p.T;
    const bool active = (int64_t)blockIdx.x * 2 + h < p.B;
    const int64_t b = active ? (int64_t)blockIdx.x * 2 + h : p.B - 1;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    GhLane<D, 3> tab[2];
    CGP_UNROLL for (int e = 0; e < 2; e++) tab[e].load(p, g + 16 * e);
    double m[D], Pc[NS], H[D];
    load_vec<D>(p.m0 + b * p.m0_stride, m);
    load_sym<D>(p.P0 + b * p.P0_stride, Pc);
    CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
    const double *__restrict__ y = ys + (b / p.ys_repeat) * T;
    double yv[2], Sk[2] = {1., 1.}, rk[2] = {0., 0.}, carry = 0.;
    CGP_UNROLL for (int e = 0; e < 2; e++) yv[e] = (g + 16 * e < T) ? __ldg(y + g + 16 * e) : 0.;
    for (int64_t t = 0; t < T; t++) {
        const int slot = (int)(t & 31);
        const double yt = __shfl_sync(0xffffffffu, slot < 16 ? yv[0] : yv[1], (lane & 16) + (slot & 15));
        if (slot == 31 && t + 1 < T) {
            CGP_UNROLL for (int e = 0; e < 2; e++) yv[e] = (t + 1 + g + 16 * e < T) ? __ldg(y + t + 1 + g + 16 * e) : 0.;
        }
        double mp[D], Pp[NS];
        GH::predict<false>(mdl, tab, red[h], res[h], nullptr, lane, m, Pc, mp, Pp);
        double S, resid;
        linear_update_fast<D, H_E1>(mp, Pp, H, p.Xi, yt, m, Pc, S, resid);
        if (g == (slot & 15)) {
            if (slot < 16) { Sk[0] = S; rk[0] = resid; } else { Sk[1] = S; rk[1] = resid; }
        }
        if (slot == 31 || t == T - 1) {
            const int n = slot + 1;
            CGP_UNROLL for (int e = 0; e < 2; e++) nl[h][g + 16 * e] = (g + 16 * e < n) ? nll_increment(Sk[e], rk[e]) : 0.;
            __syncwarp();
            if (g == 0) {
                double cc = carry;
                for (int j = 0; j < n; j++) { cc = cc + nl[h][j]; nl[h][j] = cc; }
            }
            __syncwarp();
            carry = nl[h][n - 1];
            __syncwarp();
        }
    }
    if (g == 0 && active) nell_last[b] = carry;
}

}  // namespace cgp
