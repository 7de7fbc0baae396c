// cgp_oct.cuh -- sgp_filter (+ the smoother records) for LARGE batches of the headline configuration (chirp LCD model, d = 4,
// Gauss-Hermite order 3): EIGHT lanes per chirp, four chirps per warp, four base indices per lane.
//
// With a few thousand chirps or more the sigma-point filter is no longer bound by the dependency chain of one warp but by the
// FP64 pipe (2 cycles per warp instruction whatever the number of active lanes, profiles/microbench/fp64_lanes.txt), i.e. by the
// number of FP64 warp instructions issued per chirp and step.  The warp-per-chirp kernels (cgp_fast.cuh, cgp_duo.cuh) replicate
// the Cholesky factorisation, the moments -> (mp, Pp) step and the measurement update in all 32 lanes of ONE chirp and run the
// smoother's sums in a second warp: ~275 + ~60 FP64 instructions per chirp and step.  Here the replicated part serves four
// chirps per instruction, every lane walks the base indices g, g + 8, g + 16, g + 24 (< 27) of its chirp one after the other
// (independent evaluations: the scheduler overlaps them), accumulates its 14 (+ 6 cross-covariance) partial sums in registers and
// only those go through shared memory (8 partials per sum); the smoother record of a step is evaluated every 8 steps with lane
// (chirp, j) working on step j, as the consumer warp of cgp_duo.cuh does every 32 steps.  ~215 FP64 instructions per chirp and
// step including the gains.
//
// Summation order: 4 base indices in the lane, then a tree over the 8 lanes -- not the order of the warp-per-chirp kernels, so
// results agree with them to rounding (~1e-16 per step), not bit for bit.  Everything else (softplus branch per chirp, model,
// update, nll accumulation in the reference's order, record arithmetic = gain_record) is the same code.
#pragma once
#include "cgp_fast.cuh"

namespace cgp {

struct OctCfg {
    static constexpr int D = 4, V = 2, NS = 10, NA = 14, NE = 6, DD = 16, BLK = 8, CH = 4, NEV = 4;
    static constexpr int KS = 10;                       // doubles per (sum) row of the transposition scratch: 8 partials + 2
    static constexpr int REC = ws_record<4>();
};

template <bool GAINS> struct OctSmem {
    using C = OctCfg;
    static constexpr int NSUM = GAINS ? C::NA + C::NE : C::NA;
    static constexpr int CS = 24 * C::KS + 8;           // chirp stride = 8 mod 16 doubles: the four chirps of an STS hit all banks
    static constexpr int TABP = 10;                     // 5 x 16 bytes per row: the 8 rows a quarter-warp reads hit 8 bank groups
    double tab[32][TABP];                               // per base index: xb[0..2] | W | w[0..2] | -
    static constexpr int REDN = C::CH * CS, RECN = GAINS ? C::CH * C::BLK * (((C::REC / 2) % 2 == 1) ? C::REC : C::REC + 2) : 0;
    double red[REDN > RECN ? REDN : RECN];              // [chirp][sum][8] during the steps; the block's smoother records after them
    // row pitches with an ODD number of 16-byte units: in the per-block phase lane (chirp, j) works on row j of its chirp
    static constexpr int RESP = ((NSUM / 2) % 2 == 1) ? NSUM : NSUM + 2;
    static constexpr int RINGP = 18, RECP = ((C::REC / 2) % 2 == 1) ? C::REC : C::REC + 2;
    // chirp strides: 8 mod 16 doubles (res: the 8-byte stores of two chirps fill all banks) / an odd number of 16-byte units
    // (ring: the four g = 0 lanes store 16 bytes each)
    static constexpr int RESC = C::BLK * RESP + 8, RINGC = C::BLK * RINGP + 2;
    static_assert(RESC % 16 == 8 && (RINGC / 2) % 2 == 1, "bank layout");
    double res[C::CH * RESC];                           // [chirp][step of the block][RESP]: totals of every step
    double ring[C::CH * RINGC];                         // [chirp][step of the block][RINGP]: m | P packed | S | r after the update
    double prev[C::CH][18];                             // (m, P) of the last step of the previous block
    double nl[C::CH][C::BLK];
};

// One warp = 4 chirps.  255 registers: 8 warps per SM.
template <bool H_E1, bool GAINS>
__global__ void __launch_bounds__(32, 12) gh_oct_filter_kernel(const CgpProblem p, const FilterIO io) {
    using C = OctCfg;
    using Model = ModelLCD<1>;
    using S = OctSmem<GAINS>;
    constexpr int D = C::D, V = C::V, NS = C::NS, NA = C::NA, NE = C::NE, DD = C::DD, BLK = C::BLK, NEV = C::NEV, KS = C::KS, REC = C::REC;
    constexpr int NSUM = S::NSUM, CS = S::CS, P3 = 3;
    __shared__ __align__(16) S sm;
    const int lane = threadIdx.x, c = lane >> 3, g = lane & 7;
    const int64_t T = p.T;
    const int64_t gid = (int64_t)blockIdx.x * C::CH + c;
    const bool active = gid < p.B;
    const int64_t b = active ? gid : p.B - 1;           // an idle octet shadows the last chirp (and stores nothing)
    // ---- tables: row l < 27 of the base-index table, as GhLane (quadratures.py:157-196, dimension 0 fastest)
    {
        const bool has = lane < 27;
        double w0 = 0., w1 = 0., w2 = 0.;
        if (has) { w0 = p.sig_w[lane]; w1 = p.sig_w[lane + 27]; w2 = p.sig_w[lane + 54]; }
        sm.tab[lane][0] = has ? p.sig_xi[lane * D + 0] : 0.;
        sm.tab[lane][1] = has ? p.sig_xi[lane * D + 1] : 0.;
        sm.tab[lane][2] = has ? p.sig_xi[lane * D + 2] : 0.;
        sm.tab[lane][3] = (w0 + w1) + w2;               // GhLane::load: Wl += wl[c] in order
        sm.tab[lane][4] = w0; sm.tab[lane][5] = w1; sm.tab[lane][6] = w2;
    }
    double xlast[P3];
    CGP_UNROLL for (int k = 0; k < P3; k++) xlast[k] = p.sig_xi[(k * 27) * D + (D - 1)];
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    double m[D], Pc[NS], H[D];
    load_vec<D>(p.m0 + b * p.m0_stride, m);
    load_sym<D>(p.P0 + b * p.P0_stride, Pc);
    CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
    const double Xi = p.Xi;
    const double *__restrict__ y = io.ys + (b / p.ys_repeat) * T;
    const bool store_state = io.mfs != nullptr && active;
    const bool store_nell = io.nell != nullptr && active;
    const bool gains = GAINS && io.ws != nullptr;
    double carry = 0.;
    if (g < 2) { CGP_UNROLL for (int i = 0; i < 8; i++) sm.prev[c][8 * g + i] = 0.; }
    double *redc = &sm.red[c * CS];
    double yv = (g < T) ? __ldg(y + g) : 0.;            // 8 measurements per octet and block, the next 8 in flight
    __syncwarp();

    for (int64_t t0 = 0; t0 < T; t0 += BLK) {
        const int n = (T - t0 < BLK) ? (int)(T - t0) : BLK;
        const double ynext = (t0 + BLK + g < T) ? __ldg(y + t0 + BLK + g) : 0.;
        for (int slot = 0; slot < n; slot++) {
            const double yt = __shfl_sync(0xffffffffu, yv, (lane & 24) + slot);
            // ---- prediction (filters_smoothers.py:88-121)
            double L[NS];
            chol_lower_sym_rsqrt<D>(Pc, L);
            double chi[NEV][D], slast[NEV];
            bool okl = true;
            CGP_UNROLL for (int e = 0; e < NEV; e++) {
                const double *tb = &sm.tab[g + 8 * e][0];
                const double2 x01 = *reinterpret_cast<const double2 *>(tb);
                const double x2 = tb[2];
                chi[e][0] = m[0] + L[sidx(0, 0)] * x01.x;
                chi[e][1] = m[1] + fma(L[sidx(1, 1)], x01.y, L[sidx(1, 0)] * x01.x);
                chi[e][2] = m[2] + fma(L[sidx(2, 2)], x2, fma(L[sidx(2, 1)], x01.y, L[sidx(2, 0)] * x01.x));
                slast[e] = fma(L[sidx(3, 2)], x2, fma(L[sidx(3, 1)], x01.y, L[sidx(3, 0)] * x01.x));
                okl = okl && chi[e][V] >= 3. && chi[e][V] <= 700.;
            }
            // softplus branch: per chirp, decided by all 32 slots of the chirp (the 5 idle ones sit at chi = m), as in the
            // warp-per-chirp kernels
            const unsigned bal = __ballot_sync(0xffffffffu, okl);
            const bool series = ((bal >> (lane & 24)) & 0xffu) == 0xffu;
            double gv[NEV];
            if (series) { CGP_UNROLL for (int e = 0; e < NEV; e++) gv[e] = softplus_series(chi[e][V]); }
            else { CGP_UNROLL for (int e = 0; e < NEV; e++) gv[e] = softplus_general(chi[e][V]); }
            __syncwarp();
            double acc[NSUM];
            CGP_UNROLL for (int e = 0; e < NEV; e++) {
                const double *tb = &sm.tab[g + 8 * e][0];
                const double2 x01 = *reinterpret_cast<const double2 *>(tb);
                const double2 x2W = *reinterpret_cast<const double2 *>(tb + 2);
                const double2 w01 = *reinterpret_cast<const double2 *>(tb + 4);
                const double wl[P3] = {w01.x, w01.y, tb[6]};
                const double Wl = x2W.y;
                const typename Model::Trig trig = mdl.prep_g(gv[e]);
                double ev[D], S0 = 0., S1 = 0., q00 = 0., q10 = 0., q11 = 0.;
                CGP_UNROLL for (int k = 0; k < P3; k++) {
                    chi[e][D - 1] = m[D - 1] + fma(L[sidx(D - 1, D - 1)], xlast[k], slast[e]);
                    if (k == 0) mdl.mean_with(trig, chi[e], ev); else mdl.mean_tail(chi[e], ev);
                    const double w = wl[k];
                    S0 = fma(w, ev[V], S0);
                    S1 = fma(w, ev[V + 1], S1);
                    q00 = fma(w, ev[V] * ev[V] + mdl.sig(V, V), q00);
                    q10 = fma(w, ev[V + 1] * ev[V] + mdl.sig(V + 1, V), q10);
                    q11 = fma(w, ev[V + 1] * ev[V + 1] + mdl.sig(V + 1, V + 1), q11);
                }
                double a[NSUM];
                CGP_UNROLL for (int r = 0; r < V; r++) a[r] = Wl * ev[r];
                a[V] = S0; a[V + 1] = S1;
                CGP_UNROLL for (int r = 0; r < V; r++) CGP_UNROLL for (int q = 0; q <= r; q++) {
                    double v = ev[r] * ev[q];
                    if (Model::has_sig(r, q)) v += mdl.sig(r, q);
                    a[D + sidx(r, q)] = Wl * v;
                }
                CGP_UNROLL for (int q = 0; q < V; q++) {
                    a[D + sidx(V, q)] = ev[q] * S0;
                    a[D + sidx(V + 1, q)] = ev[q] * S1;
                }
                a[D + sidx(V, V)] = q00; a[D + sidx(V + 1, V)] = q10; a[D + sidx(V + 1, V + 1)] = q11;
                if constexpr (GAINS) {                  // GhPredictLCD::cross_partials
                    const double xb[3] = {x01.x, x01.y, x2W.x};
                    CGP_UNROLL for (int cc = 0; cc < D - 1; cc++)
                        CGP_UNROLL for (int q = 0; q < V; q++) a[NA + cc * V + q] = xb[cc] * (Wl * ev[q]);
                }
                CGP_UNROLL for (int k = 0; k < NSUM; k++) acc[k] = (e == 0) ? a[k] : acc[k] + a[k];
            }
            // ---- the NSUM sums over the 8 lanes of the chirp: transposition through shared memory, fixed tree
            CGP_UNROLL for (int k = 0; k < NSUM; k++) redc[k * KS + g] = acc[k];
            __syncwarp();
            double *resc = &sm.res[c * S::RESC + slot * S::RESP];
            CGP_UNROLL for (int k0 = 0; k0 < NSUM; k0 += 8) {
                const int k = k0 + g;
                if (k0 + 8 <= NSUM || k < NSUM) {
                    const double *row = redc + k * KS;
                    double v[8];
                    CGP_UNROLL for (int j = 0; j < 8; j += 2) {
                        const double2 x = *reinterpret_cast<const double2 *>(row + j);
                        v[j] = x.x; v[j + 1] = x.y;
                    }
                    resc[k] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
                }
            }
            __syncwarp();
            double mp[D], Pp[NS];
            {
                double tot[NA];
                CGP_UNROLL for (int k2 = 0; k2 < NA; k2 += 2) {
                    const double2 x = *reinterpret_cast<const double2 *>(resc + k2);
                    tot[k2] = x.x; tot[k2 + 1] = x.y;
                }
                CGP_UNROLL for (int r = 0; r < D; r++) mp[r] = tot[r];
                CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int q = 0; q <= r; q++)
                    Pp[sidx(r, q)] = fma(-mp[r], mp[q], tot[D + sidx(r, q)]);
            }
            // ---- measurement update (filters_smoothers.py:55-68)
            double Sv, resid;
            linear_update_fast<D, H_E1>(mp, Pp, H, Xi, yt, m, Pc, Sv, resid);
            if (g == 0) {
                double *o = &sm.ring[c * S::RINGC + slot * S::RINGP];
                store_vec<D>(o, m);
                store_vec<NS>(o + D, Pc);
                *reinterpret_cast<double2 *>(o + D + NS) = make_double2(Sv, resid);
            }
        }
        yv = ynext;
        __syncwarp();
        // ---- per block: nll increments in SIMD (lane (chirp, j) = step t0 + j), accumulation in the reference's order, stores
        const double *ringc = &sm.ring[c * S::RINGC];
        sm.nl[c][g] = g < n ? nll_increment(ringc[g * S::RINGP + D + NS], ringc[g * S::RINGP + D + NS + 1]) : 0.;
        __syncwarp();
        if (g == 0) {
            double cc = carry;
            for (int j = 0; j < n; j++) { cc = cc + sm.nl[c][j]; sm.nl[c][j] = cc; }
        }
        __syncwarp();
        carry = sm.nl[c][n - 1];
        if (store_nell && !io.nell_last_only && g < n) io.nell[b * T + t0 + g] = sm.nl[c][g];
        if (store_state) {
            double2 *dm = reinterpret_cast<double2 *>(io.mfs + (b * T + t0) * D);
            for (int i = g; i < n * (D / 2); i += 8)
                dm[i] = *reinterpret_cast<const double2 *>(&ringc[(i / (D / 2)) * S::RINGP + 2 * (i % (D / 2))]);
            double2 *dP = reinterpret_cast<double2 *>(io.Pfs + (b * T + t0) * DD);
            for (int i = g; i < n * (DD / 2); i += 8) {
                const int j = i / (DD / 2), q = i % (DD / 2), r = q / (D / 2), cq = 2 * (q % (D / 2));
                dP[i] = make_double2(ringc[j * S::RINGP + D + sidx(r, cq)], ringc[j * S::RINGP + D + sidx(r, cq + 1)]);
            }
        }
        if constexpr (GAINS) {
            if (gains) {
                // lane (chirp, j): [E | tot] of iteration t0 + j and the state of step t0 + j - 1 -> workspace record t0 + j - 1
                if (g < n) {
                    const double f[4] = {mdl.f00, mdl.f01, mdl.f10, mdl.f11};
                    const double *rs = &sm.res[c * S::RESC + g * S::RESP];
                    gain_record<1>(rs + NA, rs, (g == 0) ? &sm.prev[c][0] : &ringc[(g - 1) * S::RINGP], f,
                                   &sm.red[(c * BLK + g) * S::RECP]);
                }
                __syncwarp();
                const int j0 = (t0 == 0) ? 1 : 0;       // iteration 0 predicts from (m0, P0): no smoother record
                if (active) {
                    double2 *dw = reinterpret_cast<double2 *>(io.ws + (b * T + t0 - 1 + j0) * REC);
                    for (int i = g; i < (n - j0) * (REC / 2); i += 8)
                        dw[i] = *reinterpret_cast<const double2 *>(&sm.red[(c * BLK + j0 + i / (REC / 2)) * S::RECP + 2 * (i % (REC / 2))]);
                }
            }
        }
        if (g < 2) {
            CGP_UNROLL for (int i = 0; i < 8; i += 2)
                *reinterpret_cast<double2 *>(&sm.prev[c][8 * g + i]) = *reinterpret_cast<const double2 *>(&ringc[(n - 1) * S::RINGP + 8 * g + i]);
        }
        __syncwarp();
    }
    if (store_nell && io.nell_last_only && g == 0) io.nell[b] = carry;
}

}  // namespace cgp
