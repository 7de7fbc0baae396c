// cgp_cubduo.cuh -- warp-specialised sgp_filter (+ the smoother gains) for the harmonic chirp models with the spherical
// cubature rule (quadratures.py:139-150: 2 d points m +- sqrt(d) L e_j, equal weights): BASELINE config 4 (d = 8, 16 points).
//
// Same split as cgp_duo.cuh (chirp model, Gauss-Hermite): the time loop of a sigma-point filter is one dependency chain, every
// instruction issued by the warp that walks it costs ~3-4 cycles per step, and the smoother's time-parallel half
// (filters_smoothers.py:520-527) recomputes exactly the prediction the filter has just made.  So one CTA = two warps:
//
//   producer  walks the chain for TWO chirps at once (one per half-warp, lane = sigma point; the replicated parts -- Cholesky,
//             moments, update -- are the same instructions for both): chol(P) -> the lane's column of L (the point m +- s l_j
//             needs nothing else: chi = m + L xi with xi = +- s e_j is exactly m +- s l_j) -> model mean -> d + d(d+1)/2
//             products per lane, summed over the 16 lanes through shared memory in a fixed tree -> (mp, Pp) -> measurement
//             update with the non-zero entries of H only (H = [0 1 0 1 0 1 0 0] for three harmonics);
//   consumer  does everything that is not on the chain, for both chirps: the cross-covariance (D = L E with
//             E_j = s w (mu_j+ - mu_j-): the m mp^T terms cancel identically, as in cubature_gain_kernel), the nll increments,
//             the coalesced stores of mfs / Pfs / nell, and every 8 steps the smoother records [G | c | C] (cgp_kernels.cuh),
//             a pair of lanes per (chirp, step) record: two d x d Cholesky factorisations (both lanes), d pairs of triangular
//             solves (rows dealt to the two), in place in shared memory.
//
// Hand-over exactly as in cgp_duo.cuh: NBUF buffers, FULL = one named barrier per buffer (producer bar.arrive, consumer
// bar.sync), EMPTY = a progress word the producer reads one step ahead.  255 registers x 64 threads x 4 CTAs per SM: 592 CTAs
// are resident at once, so 1000 chirps (500 CTAs) run in one wave with at most one producer per SM sub-partition.
#pragma once
#include "cgp_duo.cuh"

namespace cgp {

template <int NH> struct CubDuoCfg {
    static constexpr int D = 2 * NH + 2, V = D - 2, NS = NSym<D>::value, NA = D + NS, DD = D * D;
    static constexpr int NBUF = 5, BLK = 8;              // the consumer's 8-step flush lasts ~3 producer steps
    static constexpr int RPITCH = 18;                                   // 9 x 16 bytes: the 16-byte reads of 8 lanes hit 8 x 4 different banks
    static constexpr int NAP = (NA + 1) & ~1;
    static constexpr int SROW = (D + NS + 2 + 1) & ~1;                  // m | P packed | S | r
    static constexpr int XROW = D + 2;                                  // ev of one lane (+ 16 bytes: conflict-free 16-byte stores)
    static constexpr int REC = ws_record<D>();
    static constexpr int RROW = ((REC / 2) % 2 == 1) ? REC : REC + 2;   // odd number of 16-byte units per row
    static_assert(2 * D <= 16, "one sigma point per lane of a half-warp");
    static_assert(DD + NA <= REC, "[E | tot] is overwritten in place by [G | c | C]");
};

template <int NH> struct CubDuoSmem {
    using C = CubDuoCfg<NH>;
    double red[2][C::NA][C::RPITCH];            // producer: transposition scratch of the moment sums (per half-warp)
    double res[C::NBUF][2][C::NAP];             // totals of the step
    double stp[C::NBUF][2][C::SROW];            // m | P packed | S | r after the measurement update
    double xop[C::NBUF][2][16][C::XROW];        // model mean at every sigma point
    // consumer only
    double ring[2][C::BLK + 1][C::SROW];        // row 0: last step of the previous block; row j + 1: step t0 + j
    double rec[2][C::BLK][C::RROW];             // [E | tot] of step t0 + j -> record [G | c | C] of step t0 + j - 1
    double nl[2][C::BLK];
    int producer_warp;
    int consumed;
};

// (mp, Pp) from the unweighted totals of a cubature rule with equal weights w:  mp = w sum mu,
// Pp = w sum mu mu^T + Sigma - mp mp^T (filters_smoothers.py:119-120; sum_i w_i Sigma = Sigma).  Used by the producer and,
// for the smoother gain, by the consumer: bit-identical.
template <int NH>
CGP_DEV void cub_moments(const ModelLCD<NH> &mdl, double w, const double (&tot)[CubDuoCfg<NH>::NA], double (&mp)[2 * NH + 2],
                         double (&Pp)[NSym<2 * NH + 2>::value]) {
    constexpr int D = 2 * NH + 2;
    CGP_UNROLL for (int r = 0; r < D; r++) mp[r] = w * tot[r];
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++) {
        const double v = ModelLCD<NH>::has_sig(r, c) ? fma(w, tot[D + sidx(r, c)], mdl.sig(r, c)) : w * tot[D + sidx(r, c)];
        Pp[sidx(r, c)] = fma(-mp[r], mp[c], v);
    }
}

// Smoother record of one step, in place: row = [E (d x d) | tot (d + NSym)] -> [G | c | C]; mPq = [m | P packed] the prediction
// started from.  D = L E (L = chol(Pq)) overwrites E from the last row up (row r needs E rows <= r); the rows of G = D Pp^{-1}
// then overwrite D from the last row up as well, because C_rq = P_rq - G_r . D_q needs the rows q <= r of D.
// A PAIR of lanes works on one record (the block holds 8 steps per chirp, a half-warp has 16 lanes): both factorise (the
// Cholesky factors live in registers), member 0 takes the odd rows of D and of G, member 1 the even ones -- the rows are
// independent given the factors.  The in-place order above then needs the pair in step: every row is stored behind a
// __syncwarp() that follows the other member's reads of it.  Called by all 32 lanes; `valid` = the pair's step exists.
template <int NH>
CGP_DEV void cub_gain_record(const ModelLCD<NH> &mdl, double w, double *row, const double *mPq, int member, bool valid) {
    using C = CubDuoCfg<NH>;
    constexpr int D = C::D, NS = C::NS, NA = C::NA, DD = C::DD;
    static_assert(D % 2 == 0, "rows are dealt to the two members in pairs");
    const bool m1 = member != 0;
    {
        double Pq[NS], L[NS];
        load_vec<NS>(mPq + D, Pq);
        chol_lower_sym_rsqrt<D>(Pq, L);
        CGP_UNROLL for (int ra = D - 1; ra >= 1; ra -= 2) {          // member 0: row ra, member 1: row ra - 1 (its last term is 0 e)
            double d[D];
            CGP_UNROLL for (int k = 0; k <= ra; k++) {
                double e[D];
                load_vec<D>(row + k * D, e);
                const double coef = (k == ra) ? selp(m1, 0., L[sidx(ra, ra)]) : selp(m1, L[sidx(ra - 1, k)], L[sidx(ra, k)]);
                CGP_UNROLL for (int c = 0; c < D; c++) d[c] = (k == 0) ? coef * e[c] : fma(coef, e[c], d[c]);
            }
            __syncwarp();
            if (valid) store_vec<D>(row + (ra - member) * D, d);
        }
    }
    double mp[D], Lq[NS], rinv[D];
    {
        double tot[NA], Pp[NS];
        load_vec<NA>(row + DD, tot);
        cub_moments<NH>(mdl, w, tot, mp, Pp);
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
    __syncwarp();                                                // both members hold tot; every row of D is in place
    // a real loop: nothing in it indexes registers by r, and unrolled the rows are ~1200 instructions that the consumer
    // warp walks once per block, i.e. always cold in the instruction cache (ncu: stall_no_instruction 1.4 per issue)
    #pragma unroll 1
    for (int r = D - 1 - member; r >= 0; r -= 2) {
        double z[D];
        load_vec<D>(row + r * D, z);
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
        double cacc = mPq[r];                                   // c_r = m_r - G_r . mp
        CGP_UNROLL for (int k = 0; k < D; k++) cacc = fma(-z[k], mp[k], cacc);
        if (valid) row[DD + r] = cacc;
        #pragma unroll 1
        for (int q = 0; q <= r; q++) {                           // C_rq = P_rq - G_r . D_q   (rows q <= r of D are still in place)
            double dq[D];
            load_vec<D>(row + q * D, dq);
            double acc = mPq[D + sidx(r, q)];
            CGP_UNROLL for (int k = 0; k < D; k++) acc = fma(-z[k], dq[k], acc);
            if (valid) row[DD + D + sidx(r, q)] = acc;
        }
        __syncwarp();
        if (valid) store_vec<D>(row + r * D, z);
    }
}

template <int NH, bool H_HARM>
__global__ void __launch_bounds__(64, 4) cub_duo_filter_kernel(const CgpProblem p, const FilterIO io) {
    using Model = ModelLCD<NH>;
    using C = CubDuoCfg<NH>;
    constexpr int D = C::D, V = C::V, NS = C::NS, NA = C::NA, DD = C::DD, NBUF = C::NBUF, BLK = C::BLK, SROW = C::SROW, REC = C::REC;
    extern __shared__ __align__(16) unsigned char cubduo_smem_raw[];          // 53 KB at d = 8: dynamic (4 CTAs per SM)
    CubDuoSmem<NH> &sm = *reinterpret_cast<CubDuoSmem<NH> *>(cubduo_smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l = lane & 15, h = lane >> 4;
    const int64_t gid = (int64_t)blockIdx.x * 2 + h;
    const bool active = gid < p.B;
    const int64_t b = active ? gid : p.B - 1;              // an idle half-warp shadows the last chirp (and stores nothing)
    const int64_t T = p.T;
    if (threadIdx.x == 0) {
        // one producer per SM sub-partition: with 4 resident CTAs of two (adjacent) warp slots each, take the even slot in the
        // first two CTAs and the odd one in the others (placement heuristic only, as in cgp_duo.cuh)
        unsigned wid;
        asm("mov.u32 %0, %%warpid;" : "=r"(wid));
        sm.producer_warp = (int)(((wid >> 2) ^ wid) & 1u);
        sm.consumed = 0;
    }
    static_assert(NBUF <= 5, "named_bar_*5 cover ids 0..4");
    named_bar_sync5(0);                                     // (first use of barrier 0; completes before the loops start)
    const double w = __ldg(p.sig_w);                        // equal weights 1 / (2 d)
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);

    if (warp == sm.producer_warp) {
        // ------------------------------------------------------------------------------------------ the chain
        const bool has = l < 2 * D;
        double xi[D];                                       // the lane's row of the table: +- sqrt(d) e_j (quadratures.py:139-150)
        CGP_UNROLL for (int c = 0; c < D; c++) xi[c] = has ? __ldg(p.sig_xi + l * D + c) : 0.;
        double m[D], Pc[NS], H[D];
        load_vec<D>(p.m0 + b * p.m0_stride, m);
        load_sym<D>(p.P0 + b * p.P0_stride, Pc);
        CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
        const double Xi = p.Xi;
        const double *__restrict__ y = io.ys + (b / p.ys_repeat) * T;
        double yv = (l < T) ? __ldg(y + l) : 0., ynext = 0.;      // 16 measurements per half-warp, the next 16 in flight
        double (*red)[C::RPITCH] = sm.red[h];
        int cons = 0, buf = 0;
        for (int64_t t = 0; t < T; t++, buf = (buf + 1 == NBUF) ? 0 : buf + 1) {
            const int slot = (int)(t & 15);
            const double yt = __shfl_sync(0xffffffffu, yv, (lane & 16) + slot);
            if (slot == 0) ynext = (t + 16 + l < T) ? __ldg(y + t + 16 + l) : 0.;
            if (slot == 15) yv = ynext;
            // buffer `buf` is free once the consumer has finished step t - NBUF (value read during the previous step)
            while (cons < (int)t - NBUF + 1) cons = ld_volatile_shared(&sm.consumed);
            const int cons_next = ld_volatile_shared(&sm.consumed);
            // ---- prediction (filters_smoothers.py:88-121)
            double ev[D];
            {
                double L[NS], chi[D];
                chol_lower_sym_rsqrt<D>(Pc, L);
                CGP_UNROLL for (int r = 0; r < D; r++) {            // chi = m + L xi (quadratures.py gen_sigma_points): the zeros of
                    double v = L[sidx(r, 0)] * xi[0];               // xi add nothing, the lane's column of L is picked without a select
                    CGP_UNROLL for (int c = 1; c <= r; c++) v = fma(L[sidx(r, c)], xi[c], v);
                    chi[r] = m[r] + v;
                }
                mdl.mean_with(mdl.template prep_v<true>(chi[V]), chi, ev);
                if (!has) { CGP_UNROLL for (int r = 0; r < D; r++) ev[r] = 0.; }
            }
            CGP_UNROLL for (int r = 0; r < D; r++) red[r][l] = ev[r];
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++) red[D + sidx(r, c)][l] = ev[r] * ev[c];
            CGP_UNROLL for (int q = 0; q < D; q += 2)
                *reinterpret_cast<double2 *>(&sm.xop[buf][h][l][q]) = make_double2(ev[q], ev[q + 1]);
            __syncwarp();
            CGP_UNROLL for (int e0 = 0; e0 < NA; e0 += 16) {       // lane l adds the 16 partials of sums l, l + 16, ...
                const int e = e0 + l;
                const bool ok = e < NA;
                double v[16];
                CGP_UNROLL for (int q = 0; q < 16; q += 2) {
                    const double2 x = *reinterpret_cast<const double2 *>(&red[ok ? e : 0][q]);
                    v[q] = x.x; v[q + 1] = x.y;
                }
                CGP_UNROLL for (int w2 = 1; w2 < 16; w2 <<= 1)
                    CGP_UNROLL for (int q = 0; q + w2 < 16; q += 2 * w2) v[q] += v[q + w2];
                if (ok) sm.res[buf][h][e] = v[0];
            }
            __syncwarp();
            double mp[D], Pp[NS];
            {
                double tot[NA];
                load_vec<NA>(&sm.res[buf][h][0], tot);
                cub_moments<NH>(mdl, w, tot, mp, Pp);
            }
            // ---- measurement update (filters_smoothers.py:55-68).  H_HARM: H = sum_k e_(2k+1), the measurement row of the harmonic
            // chirp models (models.py:257): the products with 0 and 1 are exact, so sums of the selected entries give the same result
            double PH[D], S, pred;
            if constexpr (H_HARM) {
                CGP_UNROLL for (int i = 0; i < D; i++) {
                    PH[i] = Pp[sidx(i, 1)];
                    CGP_UNROLL for (int k = 1; k < NH; k++) PH[i] += Pp[sidx(i, 2 * k + 1)];
                }
                S = PH[1]; pred = mp[1];
                CGP_UNROLL for (int k = 1; k < NH; k++) { S += PH[2 * k + 1]; pred += mp[2 * k + 1]; }
            } else {
                CGP_UNROLL for (int i = 0; i < D; i++) {
                    PH[i] = Pp[sidx(i, 0)] * H[0];
                    CGP_UNROLL for (int q = 1; q < D; q++) PH[i] = fma(Pp[sidx(i, q)], H[q], PH[i]);
                }
                S = PH[0] * H[0]; pred = H[0] * mp[0];
                CGP_UNROLL for (int q = 1; q < D; q++) { S = fma(PH[q], H[q], S); pred = fma(H[q], mp[q], pred); }
            }
            S += Xi;
            const double rS = fast_rcp(S), resid = yt - pred;
            double K[D];
            CGP_UNROLL for (int i = 0; i < D; i++) K[i] = PH[i] * rS;
            CGP_UNROLL for (int i = 0; i < D; i++) m[i] = fma(K[i], resid, mp[i]);
            CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int q = 0; q <= i; q++)
                Pc[sidx(i, q)] = fma(-K[i], PH[q], Pp[sidx(i, q)]);          // as linear_update_fast
            if (l == 0) {
                double *o = &sm.stp[buf][h][0];
                store_vec<D>(o, m);
                store_vec<NS>(o + D, Pc);
                o[D + NS] = S;
                o[D + NS + 1] = resid;
            }
            named_bar_arrive5(buf);
            cons = cons_next;
        }
        return;
    }

    // ---------------------------------------------------------------------------------------------- everything else
    const bool store_state = io.mfs != nullptr && active;
    const bool store_nell = io.nell != nullptr && active;
    const bool gains = io.ws != nullptr;
    const double sw = __ldg(p.sig_xi) * w;                  // s w with s = +sqrt(d) (point 0, coordinate 0)
    double carry = 0.;                                      // cumulative nll up to the last flushed step
    for (int i = l; i < SROW; i += 16) sm.ring[h][0][i] = 0.;
    int buf = 0;
    for (int64_t t = 0; t < T; t++, buf = (buf + 1 == NBUF) ? 0 : buf + 1) {
        const int slot = (int)(t % BLK);
        named_bar_sync5(buf);
        for (int i = l; i < SROW / 2; i += 16)
            *reinterpret_cast<double2 *>(&sm.ring[h][slot + 1][2 * i]) = *reinterpret_cast<const double2 *>(&sm.stp[buf][h][2 * i]);
        if (gains) {
            for (int i = l; i < NA; i += 16) sm.rec[h][slot][DD + i] = sm.res[buf][h][i];
            for (int e = l; e < DD; e += 16) {               // E_jc = s w (mu_j+ - mu_j-)_c
                const int jj = e / D, c = e % D;
                sm.rec[h][slot][e] = sw * (sm.xop[buf][h][jj][c] - sm.xop[buf][h][jj + D][c]);
            }
        }
        __syncwarp();                                        // every lane is done with the hand-over buffers of step t
        if (lane == 0) st_volatile_shared(&sm.consumed, (int)t + 1);
        if (slot != BLK - 1 && t != T - 1) continue;
        // ---- every BLK steps (and at the end): nll increments in SIMD, sequential accumulation, coalesced stores, smoother records
        const int n = slot + 1;
        const int64_t t0 = t - slot;
        if (l < BLK) sm.nl[h][l] = l < n ? nll_increment(sm.ring[h][l + 1][D + NS], sm.ring[h][l + 1][D + NS + 1]) : 0.;
        __syncwarp();
        if (l == 0) {
            double c = carry;
            for (int q = 0; q < n; q++) { c = c + sm.nl[h][q]; sm.nl[h][q] = c; }      // reference order: n_ell = n_ell + inc
        }
        __syncwarp();
        carry = sm.nl[h][n - 1];
        if (store_nell && !io.nell_last_only && l < n) io.nell[b * T + t0 + l] = sm.nl[h][l];
        if (store_state) {
            double2 *dm = reinterpret_cast<double2 *>(io.mfs + (b * T + t0) * D);
            for (int i = l; i < n * (D / 2); i += 16)
                dm[i] = *reinterpret_cast<const double2 *>(&sm.ring[h][1 + i / (D / 2)][2 * (i % (D / 2))]);
            double2 *dP = reinterpret_cast<double2 *>(io.Pfs + (b * T + t0) * DD);
            for (int i = l; i < n * (DD / 2); i += 16) {
                const int q = i % (DD / 2), r = q / (D / 2), c = 2 * (q % (D / 2));
                const double *src = &sm.ring[h][1 + i / (DD / 2)][D];
                dP[i] = make_double2(src[sidx(r, c)], src[sidx(r, c + 1)]);
            }
        }
        if (gains) {
            // lane j: [E | tot] of iteration t0 + j and the state of step t0 + j - 1 (ring row j) -> workspace record t0 + j - 1
            // (all 32 lanes call it: the pair (l >> 1) of each half-warp works on step t0 + (l >> 1))
            static_assert(BLK == 8, "two lanes per record, 16 lanes per chirp");
            const int jr = l >> 1;
            const bool valid = jr < n;
            cub_gain_record<NH>(mdl, w, &sm.rec[h][valid ? jr : 0][0], &sm.ring[h][valid ? jr : 0][0], l & 1, valid);
            __syncwarp();
            const int j0 = (t0 == 0) ? 1 : 0;              // iteration 0 predicts from (m0, P0): no smoother record
            if (active) {
                double2 *dw = reinterpret_cast<double2 *>(io.ws + (b * T + t0 - 1 + j0) * REC);
                for (int i = l; i < (n - j0) * (REC / 2); i += 16)
                    dw[i] = *reinterpret_cast<const double2 *>(&sm.rec[h][j0 + i / (REC / 2)][2 * (i % (REC / 2))]);
            }
        }
        for (int i = l; i < SROW; i += 16) sm.ring[h][0][i] = sm.ring[h][n][i];
        __syncwarp();
    }
    if (store_nell && io.nell_last_only && l == 0) io.nell[b] = carry;
}

}  // namespace cgp
