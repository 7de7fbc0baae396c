// cgp_duo.cuh -- warp-specialised sgp_filter that also produces the smoother gains, for the headline path (chirp LCD model,
// Gauss-Hermite order 3): cgp_sgp_filter_gains_f64.
//
// At the headline batch size the sigma-point filter is bound by the dependency chain of ONE warp walking the time loop
// (profiles/: ~4 cycles per issued instruction, FP64 pipe < 45 % busy): every instruction added to that warp costs ~4
// cycles per step, while more than half of the issue slots of the SM sub-partition stay idle.  So the CTA has TWO warps:
//
//   producer  runs nothing but the chain of filters_smoothers.py:480-487 -- Cholesky, sigma points, model, the 14 moment
//             sums, measurement update -- and hands each step over through a small ring of shared-memory buffers:
//             (m, P, S, r), the 14 moment totals and, for the smoother gains, the two per-lane model values the six
//             cross-covariance partial sums are made of (GhPredictLCD::cross_partials);
//   consumer  does everything that is not on the chain: the nll increments (32 at a time, in SIMD, accumulated in the
//             reference's order), the coalesced 16-byte stores of mfs / Pfs / nell, and the smoother workspace
//             [G | c | C] that sgp_smoother's time-parallel half would otherwise recompute from (mf, Pf)
//             (filters_smoothers.py:520-527): the cross sums reduced over the lanes, then gain_record.
//
// Hand-over: NBUF buffers.  FULL: one mbarrier per buffer (an elected producer lane arrives -- it never waits --, the consumer
// spins on try_wait; a named barrier needs an immediate id, and dispatching on it cost ~45-60 cycles at the end of every step).
// EMPTY: a bar.sync on the producer side would put the barrier's ~100-cycle latency on the chain at every step, so the
// consumer publishes the number of steps it has finished in a shared-memory word instead; the producer reads it one step
// ahead of the use (latency hidden) and only spins if the consumer has fallen NBUF steps behind, which does not happen
// in steady state: the consumer needs ~1/4 of the producer's time per step.  NBUF = 4 (measured 2 / 4 / 8 buffers: 3.76 / 3.64 /
// 3.66 ms at 1000 chirps: two are too few to ride out the consumer's 32-step flush).
#pragma once
#include "cgp_fast.cuh"

namespace cgp {

// Named barriers with IMMEDIATE ids (a register id makes ptxas reserve all 16 hardware barriers of the CTA, and barriers
// are an SM resource: 16 per CTA would cap the SM at 4 CTAs).  64 = both warps of the CTA.  One predicated instruction per
// candidate id instead of a branch tree.  Used by cgp_cubduo.cuh; gh_duo_filter_kernel hands over through mbarriers (below):
// even predicated, the WARPSYNC + BAR.ARV pairs cost ~60 cycles at the end of every step of the chain.
CGP_DEV void named_bar_arrive5(int id) {
    asm volatile("{\n .reg .pred q;\n"
                 " setp.eq.s32 q, %0, 0;\n @q bar.arrive 0, 64;\n setp.eq.s32 q, %0, 1;\n @q bar.arrive 1, 64;\n"
                 " setp.eq.s32 q, %0, 2;\n @q bar.arrive 2, 64;\n setp.eq.s32 q, %0, 3;\n @q bar.arrive 3, 64;\n"
                 " setp.eq.s32 q, %0, 4;\n @q bar.arrive 4, 64;\n}" ::"r"(id) : "memory");
}
CGP_DEV void named_bar_sync5(int id) {
    asm volatile("{\n .reg .pred q;\n"
                 " setp.eq.s32 q, %0, 0;\n @q bar.sync 0, 64;\n setp.eq.s32 q, %0, 1;\n @q bar.sync 1, 64;\n"
                 " setp.eq.s32 q, %0, 2;\n @q bar.sync 2, 64;\n setp.eq.s32 q, %0, 3;\n @q bar.sync 3, 64;\n"
                 " setp.eq.s32 q, %0, 4;\n @q bar.sync 4, 64;\n}" ::"r"(id) : "memory");
}
// mbarrier (shared-memory barrier object, address in a register): the FULL side of the hand-over in gh_duo_filter_kernel.
// One elected lane arrives (release), the consumer spins on try_wait (acquire); no convergence barrier, no id dispatch.
CGP_DEV void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
CGP_DEV void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
CGP_DEV void mbar_wait(unsigned long long *bar, int parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
CGP_DEV int ld_volatile_shared(const int *q) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(q)) : "memory");
    return v;
}
CGP_DEV void st_volatile_shared(int *q, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(q)), "r"(v) : "memory");
}

struct DuoSmem4 {                       // D = 4 (chirp model)
    static constexpr int D = 4, V = 2, NS = 10, DD = 16, NA = 14, NE = 6, OTOT = 8, NBUF = 4;
    static constexpr int SROW = 18;     // consumer state ring row: m (4) | P packed (10) | S | r | pad 2  (9 x 16 bytes: odd)
    static constexpr int WROW = 38;     // consumer gain ring row: E (6) | pad | tot (14) -> [G | c | C] (30)   (19 x 16 bytes)
    // producer <-> consumer hand-over, NBUF deep
    double red[RedRows<NA>::value][kRedPitch];  // producer: transposition scratch of the moment sums
    double res[NBUF][16];               // moment totals of the step (producer reads them back, consumer keeps them)
    double stp[NBUF][16];               // m | P packed | S | r   after the measurement update
    double xop[NBUF][V][33];            // per-lane cross-covariance operands ev[0..V-1]
    // consumer only
    double red2[NE][kSmallSumPitch];
    double ring[32][SROW];
    double ring2[32][WROW];
    double nl[32];
    double prev[D + NS + 2];            // filtering mean and covariance of the last step of the previous 32-block
    double ybuf[2][32];                 // measurements of two 32-step blocks: written by the consumer, read by the producer
    unsigned long long full[NBUF];      // mbarriers: buffer b holds step t (t % NBUF == b, phase parity (t / NBUF) & 1)
    int producer_warp;
    int consumed;                       // steps the consumer has finished reading (EMPTY side of the hand-over)
};

// One CTA = one chirp = 2 warps.  128 registers: 7 CTAs = 14 warps per SM need 4 warps on one sub-partition (16 K registers).
template <bool H_E1>
__global__ void __maxnreg__(128) gh_duo_filter_kernel(const CgpProblem p, const FilterIO io) {
    using Pred = GhPredictLCD<1, 3>;
    using S = DuoSmem4;
    constexpr int D = S::D, V = S::V, NS = S::NS, DD = S::DD, NA = S::NA, NE = S::NE, OTOT = S::OTOT, NBUF = S::NBUF;
    constexpr int WREC = ws_record<D>();
    static_assert(Pred::D == D && Pred::NA == NA && Pred::NE == NE && Pred::V == V, "layout");
    __shared__ __align__(16) S sm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b = blockIdx.x;
    const int64_t T = p.T;
    if (threadIdx.x == 0) {
        // Spread the producers over the four sub-partitions of the SM: hardware warp slots are handed out in order and slot w
        // runs on sub-partition w % 4, so take the even slot of the CTA's (adjacent) pair in every other pair.  Only a
        // placement heuristic (measured: 3.3 vs 5.2 ms when warp 0 is always the producer); any choice is correct.
        unsigned wid;
        asm("mov.u32 %0, %%warpid;" : "=r"(wid));
        sm.producer_warp = (int)(((wid >> 2) ^ wid) & 1u);
        sm.consumed = 0;
        CGP_UNROLL for (int i = 0; i < NBUF; i++) mbar_init(&sm.full[i], 1);
    }
    // Measurements: the CONSUMER fetches them, 32 per coalesced load, two blocks ahead of the producer, into a double buffer in
    // shared memory -- no global load (and no register for a block in flight) on the chain, and the load latency stays hidden
    // even when `ys` is pinned HOST memory read over PCIe (zero-copy input of host callers).  Blocks 0 and 1 before the loops:
    const double *__restrict__ y = io.ys + (b / p.ys_repeat) * T;
    sm.ybuf[warp][lane] = (32 * warp + lane < T) ? __ldg(y + 32 * warp + lane) : 0.;
    __syncthreads();                                        // mbarriers initialised, measurement blocks 0 and 1 in place
    const bool producer = warp == sm.producer_warp;

    if (producer) {
        // ------------------------------------------------------------------------------------------ the chain
        Pred pred;
        pred.load(p, b, lane);
        double m[D], Pc[NS], H[D];
        load_vec<D>(p.m0 + b * p.m0_stride, m);
        load_sym<D>(p.P0 + b * p.P0_stride, Pc);
        CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
        double yv = sm.ybuf[0][lane];                       // one block in a register, broadcast by shuffle
        int cons = 0;
        // blocks of 32 steps (= 8 rounds of the NBUF buffers): the inner loop is the chain and nothing else
        for (int t0 = 0; t0 < (int)T; t0 += 32) {
            const int n = ((int)T - t0 < 32) ? (int)T - t0 : 32;
            for (int slot = 0; slot < n; slot++) {
                const int t = t0 + slot, buf = slot % NBUF;
                const double yt = __shfl_sync(0xffffffffu, yv, slot);
                // buffer `buf` is free once the consumer has finished step t - NBUF (value read during the previous step)
                while (cons < t - NBUF + 1) cons = ld_volatile_shared(&sm.consumed);
                const int cons_next = ld_volatile_shared(&sm.consumed);
                double mp[D], Pp[NS];
                pred.template predict_impl<true>(sm.red, &sm.res[buf][0], sm.xop[buf], lane, m, Pc, mp, Pp);
                double Sv, resid;
                linear_update_fast<D, H_E1>(mp, Pp, H, p.Xi, yt, m, Pc, Sv, resid);
                if (lane == 0) {
                    store_vec<D>(&sm.stp[buf][0], m);
                    store_vec<NS>(&sm.stp[buf][D], Pc);
                    *reinterpret_cast<double2 *>(&sm.stp[buf][D + NS]) = make_double2(Sv, resid);
                }
                __syncwarp();                               // the lanes' xop / res / stp stores are ordered before the arrive
                if (lane == 0) mbar_arrive(&sm.full[buf]);
                cons = cons_next;
            }
            // block t0 / 32 + 1 was stored by the consumer while it flushed block t0 / 32 - 1, i.e. before it published
            // step t0 - 29; the producer is never more than NBUF steps ahead of the published count
            yv = sm.ybuf[((t0 >> 5) + 1) & 1][lane];
        }
        return;
    }

    // ---------------------------------------------------------------------------------------------- everything else
    GhLane<D, 3> tab;
    Pred::Model mdl;
    tab.load(p, lane);
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    const bool store_nell = io.nell != nullptr;
    double carry = 0.;                 // cumulative nll up to the last flushed step
    if (lane < D + NS + 2) sm.prev[lane] = 0.;
    for (int64_t t = 0; t < T; t++) {
        const int slot = (int)(t & 31), buf = (int)(t % NBUF);
        mbar_wait(&sm.full[buf], (int)((t / NBUF) & 1));
        if (lane < 16) sm.ring[slot][lane] = sm.stp[buf][lane];
        {
            double ev[V], ec[NE];
            CGP_UNROLL for (int q = 0; q < V; q++) ev[q] = sm.xop[buf][q][lane];
            Pred::cross_partials(tab, ev, ec);
            CGP_UNROLL for (int k = 0; k < NE; k++) sm.red2[k][lane] = ec[k];
            if (lane < NA) sm.ring2[slot][OTOT + lane] = sm.res[buf][lane];
            __syncwarp();
            small_sums_tail<NE>(sm.red2, &sm.ring2[slot][0], lane);
        }
        __syncwarp();                                       // every lane is done with the hand-over buffers of step t
        if (lane == 0) st_volatile_shared(&sm.consumed, (int)t + 1);
        if (slot != 31 && t != T - 1) continue;
        // ---- every 32 steps (and at the end): nll increments in SIMD, sequential accumulation, coalesced stores
        const int n = slot + 1;
        const int64_t t0 = t - slot;
        // measurements of the block after the next one (the producer took this buffer's previous content 32 steps ago)
        const double ynew = (slot == 31 && t + 33 + lane < T) ? __ldg(y + t + 33 + lane) : 0.;
        sm.nl[lane] = lane < n ? nll_increment(sm.ring[lane][D + NS], sm.ring[lane][D + NS + 1]) : 0.;
        __syncwarp();
        if (lane == 0) {
            double c = carry;
            for (int j = 0; j < n; j++) { c = c + sm.nl[j]; sm.nl[j] = c; }     // reference order: n_ell = n_ell + inc
        }
        __syncwarp();
        carry = sm.nl[n - 1];
        if (store_nell && !io.nell_last_only && lane < n) io.nell[b * T + t0 + lane] = sm.nl[lane];
        double2 *dm = reinterpret_cast<double2 *>(io.mfs + (b * T + t0) * D);
        for (int i = lane; i < n * (D / 2); i += 32)
            dm[i] = *reinterpret_cast<const double2 *>(&sm.ring[i / (D / 2)][2 * (i % (D / 2))]);
        double2 *dP = reinterpret_cast<double2 *>(io.Pfs + (b * T + t0) * DD);
        for (int i = lane; i < n * (DD / 2); i += 32) {
            const int j = i / (DD / 2), q = i % (DD / 2), r = q / (D / 2), c = 2 * (q % (D / 2));
            dP[i] = make_double2(sm.ring[j][D + sidx(r, c)], sm.ring[j][D + sidx(r, c + 1)]);
        }
        // lane j: record of iteration t0 + j = workspace record t0 + j - 1 (its prediction started from step t0 + j - 1)
        if (lane < n) {
            const double f[4] = {mdl.f00, mdl.f01, mdl.f10, mdl.f11};
            gain_record<1>(&sm.ring2[lane][0], &sm.ring2[lane][OTOT], (lane == 0) ? &sm.prev[0] : &sm.ring[lane - 1][0], f,
                           &sm.ring2[lane][0]);
        }
        __syncwarp();
        const int j0 = (t0 == 0) ? 1 : 0;              // iteration 0 predicts from (m0, P0): no smoother record
        double2 *dw = reinterpret_cast<double2 *>(io.ws + (b * T + t0 - 1 + j0) * WREC);
        for (int i = lane; i < (n - j0) * (WREC / 2); i += 32)
            dw[i] = *reinterpret_cast<const double2 *>(&sm.ring2[j0 + i / (WREC / 2)][2 * (i % (WREC / 2))]);
        if (lane < D + NS) sm.prev[lane] = sm.ring[31][lane];
        if (slot == 31) sm.ybuf[(t >> 5) & 1][lane] = ynew;
        __syncwarp();
    }
    if (store_nell && io.nell_last_only && lane == 0) io.nell[b] = carry;
}

}  // namespace cgp
