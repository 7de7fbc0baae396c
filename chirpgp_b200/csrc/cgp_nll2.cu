// cgp_nll2.cu -- MLE path, second generation: EKF negative log-likelihood and its reverse-mode adjoint as PERSISTENT,
// ticket-scheduled kernels (what jax.grad(lambda th: ekf(...)[-1][-1]) computes in demos/ekfs_mle.py:42-49 by
// differentiating through lax.scan; BASELINE config 5: 10 000 chirps x 16 candidates x T = 1e5).
//
// Decomposition.  32 problems (chirp x candidate) form a CHAIN that one warp advances, one thread per problem, everything in
// registers.  Time is cut into SEGMENTS of `ckpt_every` steps; a UNIT of work is (chain, segment).  Warps are persistent:
// each takes the next unit from a global ticket counter (segment-major order), waits until the previous unit of that chain is
// published (per-chain progress word, release / acquire), runs the segment from the chain's checkpoint and publishes the next
// one.  A chain therefore migrates between warps / SMs at every segment boundary, and the work stays balanced whatever the
// number of chains: 625 chains (config 5 on 8 GPUs: 20 000 problems per GPU) keep 592 warps -- one per SM sub-partition -- busy
// all the time, where a static thread-per-problem launch leaves 33 sub-partitions with twice the work of the others; 5 000
// chains (one GPU) run without the partly filled last wave.  Deadlock-free: a unit's predecessor always holds a smaller
// ticket, so it was taken by a warp that is running (or done); waits are bounded anyway (abort word) so that a bug cannot hang
// the GPU.
//
// Arithmetic.  The covariance is carried as its packed lower triangle (10 instead of 16 entries for d = 4): in exact arithmetic
// the EKF covariance is symmetric, the reference's full-matrix products (filters_smoothers.py:257, :66) differ from the packed
// ones by rounding only (nll agrees to ~1e-13 relative, tests/test_gpu_mle.py).  The Jacobian's structural zeros are skipped
// at compile time.  Forward:  ~215 FP64 instructions per step for d = 4 with H = e_1 (the old full-matrix kernel: ~490).
//
// Adjoint.  One unit = recompute the segment forward from its checkpoint, keeping per step the inputs (m, P) and the
// transcendental results (e cos, e sin, sigmoid) in a scratch slot PRIVATE TO THE WARP (ckpt_every x 17 x 256 bytes, reused
// for every unit the warp runs, so it lives in L2 instead of streaming through HBM), then sweep the segment in reverse.
// With W = Pb + Pb^T (the only combination of the covariance cotangent the recursion reads; packed symmetric) and
// c = Pp h, S = h.c + Xi, v = y - h.mp  (one step: m' = mp + c v/S, P' = Pp - c c^T/S, l += (log 2 pi S + v^2/S)/2):
//     Sb  = lw (1/S - v^2/S^2)/2 - (mb.c) v/S^2 + (c^T W c)/(2 S^2)        vb = (lw v + mb.c)/S
//     cb  = mb v/S - W c/S + Sb h          mpb = mb - vb h                  Xib += Sb
//     Wp  = W + cb h^T + h cb^T            Sigmab += Wp/2                   Jb = Wp (J P)      (P symmetric)
//     W'  = J^T Wp J                       mb' = J^T mpb + sum_ij Jb_ij dJ_ij/dm
// The antisymmetric part A = (Pb - Pb^T)/2 of JAX's unsymmetrised covariance cotangent feeds nothing but itself
// (A' = J^T (A + (cb h^T - h cb^T)/2) J); it is tracked only when the caller asks for the raw P0 cotangent (RAW).
#include <stdlib.h>
#include "cgp_dispatch.cuh"

namespace cgp {

namespace nll2 {

constexpr int kLanes = 32;
constexpr unsigned kSpinLimit = 1u << 24;      // x >= 100 ns: a wait that long means a bug, not load imbalance

template <int NH> struct Shape {
    using Model = ModelLCD<NH>;
    static constexpr int D = Model::D, V = Model::V, NS = NSym<D>::value, NA = D * (D - 1) / 2;
    static constexpr int CK = D + NS + 1;              // checkpoint record: m, P (packed lower), running nll
    static constexpr int REC = D + NS + 2 * NH + 1;    // scratch record: m, P, (e cos, e sin) per harmonic, sigmoid
    static constexpr int NACC = 10;                    // eb, fb[4], qb, sb00, sb01, sb11, xib
    static constexpr int CARRY_SYM = D + NS + NACC, CARRY_RAW = D + NS + NA + NACC;
};

CGP_DEV constexpr int aidx(int i, int j) { return i * (i - 1) / 2 + j; }     // strictly lower (i > j)

CGP_DEV unsigned ld_acquire(const unsigned *q) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(q) : "memory");
    return v;
}
CGP_DEV void st_release(unsigned *q, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(q), "r"(v) : "memory"); }

// scheduling words: [0,1] 64-bit ticket, [2] abort, [3] unused, [4 + c] progress of chain c
struct Sched {
    unsigned *w;
    CGP_DEV unsigned long long ticket(int lane) const {
        unsigned long long n = 0;
        if (lane == 0) n = atomicAdd(reinterpret_cast<unsigned long long *>(w), 1ull);
        return __shfl_sync(0xffffffffu, n, 0);
    }
    // true when chain c has published `need` units; false = aborted
    CGP_DEV bool wait(int c, unsigned need, int lane) const {
        int ok = 1;
        if (lane == 0) {
            unsigned spins = 0;
            while (ld_acquire(w + 4 + c) < need) {
                __nanosleep(100);
                if (++spins > kSpinLimit || ld_acquire(w + 2) != 0u) { atomicExch(w + 2, 1u); ok = 0; break; }
            }
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        __threadfence();
        return ok != 0;
    }
    CGP_DEV void publish(int c, unsigned done, int lane) const {
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release(w + 4 + c, done);
    }
};

#define CGP_SSUM(nzexpr, aexpr, bexpr)                                                            \
    double sacc = 0.; bool first = true;                                                          \
    CGP_UNROLL for (int k = 0; k < D; k++)                                                        \
        if (nzexpr) { sacc = first ? (aexpr) * (bexpr) : fma((aexpr), (bexpr), sacc); first = false; }

template <int NH> struct Lin { double ce[NH], se[NH], sg; };      // e cos(theta_k), e sin(theta_k), sigmoid(u_V)

// the transcendental part of one step (models.py:296-298, :370-372), same operation order as ModelLCD::mean_jac
template <int NH> CGP_DEV void trig(const ModelLCD<NH> &mdl, double uv, Lin<NH> &tr) {
    double gv;
    softplus_and_sigmoid(uv, gv, tr.sg);
    const double w = (kTwoPi * gv) * mdl.fs;
    CGP_UNROLL for (int k = 0; k < NH; k++) {
        double sn, cs;
        fast_sincos((mdl.dt * (double)(k + 1)) * w, &sn, &cs);
        tr.ce[k] = cs * mdl.e; tr.se[k] = sn * mdl.e;
    }
}
// mean and Jacobian (structural zeros are never read) from the state and its transcendental results
template <int NH>
CGP_DEV void mean_jac_from(const ModelLCD<NH> &mdl, const Lin<NH> &tr, const double (&u)[2 * NH + 2], double (&mp)[2 * NH + 2],
                           double (&J)[2 * NH + 2][2 * NH + 2]) {
    constexpr int D = 2 * NH + 2, V = D - 2;
    const double dw = (kTwoPi * tr.sg) * mdl.fs;
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) J[r][c] = 0.;
    CGP_UNROLL for (int k = 0; k < NH; k++) {
        const double dth = (mdl.dt * (double)(k + 1)) * dw, ce = tr.ce[k], se = tr.se[k];
        const double u0 = u[2 * k], u1 = u[2 * k + 1];
        mp[2 * k] = fma(-se, u1, ce * u0);
        mp[2 * k + 1] = fma(ce, u1, se * u0);
        J[2 * k][2 * k] = ce;     J[2 * k][2 * k + 1] = -se;
        J[2 * k + 1][2 * k] = se; J[2 * k + 1][2 * k + 1] = ce;
        J[2 * k][V] = -mp[2 * k + 1] * dth;
        J[2 * k + 1][V] = mp[2 * k] * dth;
    }
    mp[V] = fma(mdl.f01, u[V + 1], mdl.f00 * u[V]);
    mp[V + 1] = fma(mdl.f11, u[V + 1], mdl.f10 * u[V]);
    J[V][V] = mdl.f00; J[V][V + 1] = mdl.f01; J[V + 1][V] = mdl.f10; J[V + 1][V + 1] = mdl.f11;
}
// JP = J P, P packed symmetric
template <class Model, int D> CGP_DEV void jmul_sym(const double (&J)[D][D], const double (&P)[NSym<D>::value], double (&JP)[D][D]) {
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j < D; j++) { CGP_SSUM(Model::jnz(i, k), J[i][k], P[sidx(k, j)]) JP[i][j] = sacc; }
}
// Pp = JP J^T + Sigma, lower triangle only
template <class Model, int D>
CGP_DEV void jpjt_sym(const Model &mdl, const double (&JP)[D][D], const double (&J)[D][D], double (&Pp)[NSym<D>::value]) {
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j <= i; j++) {
        CGP_SSUM(Model::jnz(j, k), JP[i][k], J[j][k])
        Pp[sidx(i, j)] = Model::has_sig(i, j) ? sacc + mdl.sig(i, j) : sacc;
    }
}
// c = Pp h without forming Pp: JP (J^T h) + Sigma h
template <class Model, int D, bool H_E1>
CGP_DEV void gain_column(const Model &mdl, const double (&JP)[D][D], const double (&J)[D][D], const double (&H)[D], double (&c)[D]) {
    if constexpr (H_E1) {
        CGP_UNROLL for (int i = 0; i < D; i++) {
            CGP_SSUM(Model::jnz(1, k), JP[i][k], J[1][k])
            c[i] = Model::has_sig(i, 1) ? sacc + mdl.sig(i, 1) : sacc;
        }
    } else {
        double g[D];
        CGP_UNROLL for (int k2 = 0; k2 < D; k2++) { CGP_SSUM(Model::jnz(k, k2), J[k][k2], H[k]) g[k2] = sacc; }
        CGP_UNROLL for (int i = 0; i < D; i++) {
            double s = JP[i][0] * g[0];
            CGP_UNROLL for (int k = 1; k < D; k++) s = fma(JP[i][k], g[k], s);
            CGP_UNROLL for (int j = 0; j < D; j++) if (Model::has_sig(i, j)) s = fma(mdl.sig(i, j), H[j], s);
            c[i] = s;
        }
    }
}

// measurement update on the packed covariance (filters_smoothers.py:55-68); returns S, r and 1/S
template <int D, bool H_E1>
CGP_DEV void update_sym(const double (&mp)[D], const double (&Pp)[NSym<D>::value], const double (&H)[D], double Xi, double y,
                        double (&mf)[D], double (&Pf)[NSym<D>::value], double &S, double &r, double &rS) {
    double PH[D], pred;
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
    rS = fast_rcp(S);
    r = y - pred;
    double K[D];
    CGP_UNROLL for (int i = 0; i < D; i++) K[i] = PH[i] * rS;
    CGP_UNROLL for (int i = 0; i < D; i++) mf[i] = fma(K[i], r, mp[i]);
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j <= i; j++) Pf[sidx(i, j)] = fma(-K[i], PH[j], Pp[sidx(i, j)]);
}

// one forward step from (m, P) with its transcendental results given; returns the nll increment if WANT_NLL
template <int NH, bool H_E1, bool WANT_NLL>
CGP_DEV double step_from(const ModelLCD<NH> &mdl, const Lin<NH> &tr, const double (&H)[2 * NH + 2], double Xi, double y,
                         double (&m)[2 * NH + 2], double (&P)[NSym<2 * NH + 2>::value]) {
    using Model = ModelLCD<NH>;
    constexpr int D = Model::D;
    double mp[D], J[D][D], JP[D][D], Pp[NSym<D>::value];
    mean_jac_from<NH>(mdl, tr, m, mp, J);
    jmul_sym<Model, D>(J, P, JP);
    jpjt_sym<Model, D>(mdl, JP, J, Pp);
    double S, r, rS;
    update_sym<D, H_E1>(mp, Pp, H, Xi, y, m, P, S, r, rS);
    if constexpr (WANT_NLL) return fma(r * r, rS, fast_log_pos(kTwoPi * S)) * 0.5;    // (log(2 pi S) + r^2 / S) / 2  (:44-45)
    return 0.;
}

// ------------------------------------------------------------------------------------------------ forward
// geometry of one launch: chains of `lanes` problems (32, or 16 / 8 when problems are scarce: more, narrower warps hide each
// other's latency; lanes >= `lanes` of a warp idle), `nseg` segments of `every` steps, units of `spu` segments
struct Geo {
    int every, nseg, nchains, lanes, spu, nunits;
};

template <int NH, bool H_E1>
__global__ void __launch_bounds__(128) fwd_kernel(const CgpProblem p, const double *__restrict__ ys, double *__restrict__ nll,
                                                  double *__restrict__ ckpt, const Geo g, unsigned *__restrict__ sched_words) {
    using Sh = Shape<NH>;
    using Model = ModelLCD<NH>;
    constexpr int D = Sh::D, NS = Sh::NS, CK = Sh::CK;
    const int lane = threadIdx.x & 31;
    const Sched sched{sched_words};
    const unsigned long long total = (unsigned long long)g.nunits * (unsigned long long)g.nchains;
    const unsigned long long nworkers = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long next_static = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double H[D];
    CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
    for (;;) {
        unsigned long long n;
        if (sched_words) n = sched.ticket(lane);
        else { n = next_static; next_static += nworkers; }            // no workspace: one unit per chain, dealt statically
        if (n >= total) break;
        const int u = (int)(n / (unsigned)g.nchains), c = (int)(n % (unsigned)g.nchains);
        if (sched_words && u > 0 && !sched.wait(c, (unsigned)u, lane)) break;
        const int64_t b = (int64_t)c * g.lanes + lane;
        const bool live = lane < g.lanes && b < p.B;
        const int64_t bb = live ? b : ((int64_t)c * g.lanes < p.B ? (int64_t)c * g.lanes : p.B - 1);
        Model mdl;
        mdl.load(p.consts + bb * p.consts_stride, p.dt);
        const int s0 = u * g.spu, s1 = (s0 + g.spu < g.nseg) ? s0 + g.spu : g.nseg;
        double m[D], P[NS], acc;
        if (s0 == 0) {
            load_vec<D>(p.m0 + bb * p.m0_stride, m);
            load_sym<D>(p.P0 + bb * p.P0_stride, P);
            acc = 0.;
        } else {
            const double *q = ckpt + ((int64_t)s0 * g.nchains + c) * (CK * kLanes) + lane;
            CGP_UNROLL for (int i = 0; i < D; i++) m[i] = __ldcg(q + i * kLanes);
            CGP_UNROLL for (int i = 0; i < NS; i++) P[i] = __ldcg(q + (D + i) * kLanes);
            acc = __ldcg(q + (D + NS) * kLanes);
        }
        const double *__restrict__ y = ys + (bb / p.ys_repeat) * p.T;
        int64_t t = (int64_t)s0 * g.every;
        double ynext = __ldg(y + t);
        for (int s = s0; s < s1; s++) {
            const int64_t t1 = (t + g.every < p.T) ? t + g.every : p.T;
            for (; t < t1; t++) {
                const double yt = ynext;
                if (t + 1 < p.T) ynext = __ldg(y + t + 1);
                Lin<NH> tr;
                trig<NH>(mdl, m[Sh::V], tr);
                acc = acc + step_from<NH, H_E1, true>(mdl, tr, H, p.Xi, yt, m, P);
            }
            if (s + 1 < g.nseg && ckpt) {
                double *q = ckpt + ((int64_t)(s + 1) * g.nchains + c) * (CK * kLanes) + lane;
                CGP_UNROLL for (int i = 0; i < D; i++) __stcg(q + i * kLanes, m[i]);
                CGP_UNROLL for (int i = 0; i < NS; i++) __stcg(q + (D + i) * kLanes, P[i]);
                __stcg(q + (D + NS) * kLanes, acc);
            }
        }
        if (s1 < g.nseg) sched.publish(c, (unsigned)(u + 1), lane);
        else if (live && nll) nll[b] = acc;
    }
}

// ------------------------------------------------------------------------------------------------ adjoint
template <int NH, bool RAW> struct Carry {
    using Sh = Shape<NH>;
    double mb[Sh::D], W[Sh::NS], A[RAW ? Sh::NA : 1];
    double eb, fb[4], qb, sb00, sb01, sb11, xib;
    CGP_DEV void zero() {
        CGP_UNROLL for (int i = 0; i < Sh::D; i++) mb[i] = 0.;
        CGP_UNROLL for (int i = 0; i < Sh::NS; i++) W[i] = 0.;
        CGP_UNROLL for (int i = 0; i < (RAW ? Sh::NA : 1); i++) A[i] = 0.;
        eb = qb = sb00 = sb01 = sb11 = xib = 0.;
        CGP_UNROLL for (int i = 0; i < 4; i++) fb[i] = 0.;
    }
    template <class F> CGP_DEV void each(F &&f) {
        int k = 0;
        CGP_UNROLL for (int i = 0; i < Sh::D; i++) f(k++, mb[i]);
        CGP_UNROLL for (int i = 0; i < Sh::NS; i++) f(k++, W[i]);
        if constexpr (RAW) { CGP_UNROLL for (int i = 0; i < Sh::NA; i++) f(k++, A[i]); }
        f(k++, eb); f(k++, fb[0]); f(k++, fb[1]); f(k++, fb[2]); f(k++, fb[3]);
        f(k++, qb); f(k++, sb00); f(k++, sb01); f(k++, sb11); f(k++, xib);
    }
};

// reverse of one step; (m, P, tr) are the step's inputs, y its measurement, lw the weight of its nll increment
template <int NH, bool H_E1, bool RAW>
CGP_DEV void reverse_step(const ModelLCD<NH> &mdl, const double (&H)[2 * NH + 2], double Xi, double y, double lw,
                          const double (&m)[2 * NH + 2], const double (&P)[NSym<2 * NH + 2>::value], const Lin<NH> &tr,
                          Carry<NH, RAW> &cy) {
    using Model = ModelLCD<NH>;
    using Sh = Shape<NH>;
    constexpr int D = Sh::D, V = Sh::V, NS = Sh::NS;
    double mp[D], J[D][D], JP[D][D], c[D];
    mean_jac_from<NH>(mdl, tr, m, mp, J);
    jmul_sym<Model, D>(J, P, JP);
    gain_column<Model, D, H_E1>(mdl, JP, J, H, c);
    double S, pred;
    if constexpr (H_E1) { S = c[1] + Xi; pred = mp[1]; }
    else {
        S = c[0] * H[0]; pred = H[0] * mp[0];
        CGP_UNROLL for (int i = 1; i < D; i++) { S = fma(c[i], H[i], S); pred = fma(H[i], mp[i], pred); }
        S += Xi;
    }
    const double v = y - pred, iS = fast_rcp(S), iS2 = iS * iS;
    // ---- reverse of the measurement update
    double Wc[D], mc = 0., cWc = 0.;
    CGP_UNROLL for (int i = 0; i < D; i++) mc = fma(cy.mb[i], c[i], mc);
    CGP_UNROLL for (int i = 0; i < D; i++) {
        double s = cy.W[sidx(i, 0)] * c[0];
        CGP_UNROLL for (int k = 1; k < D; k++) s = fma(cy.W[sidx(i, k)], c[k], s);
        Wc[i] = s;
        cWc = fma(c[i], s, cWc);
    }
    const double Sb = fma(lw * 0.5, fma(-v * v, iS2, iS), fma(0.5 * cWc, iS2, -(mc * v) * iS2));
    const double vb = fma(lw, v, mc) * iS;
    cy.xib += Sb;
    double cb[D], mpb[D], Wp[NS];
    CGP_UNROLL for (int i = 0; i < D; i++) {
        const double t = fma(cy.mb[i], v, -Wc[i]) * iS;
        if constexpr (H_E1) { cb[i] = (i == 1) ? t + Sb : t; mpb[i] = (i == 1) ? cy.mb[i] - vb : cy.mb[i]; }
        else { cb[i] = fma(Sb, H[i], t); mpb[i] = fma(-vb, H[i], cy.mb[i]); }
    }
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j <= i; j++) {
        if constexpr (H_E1) {
            double w = cy.W[sidx(i, j)];
            if (j == 1) w += cb[i];
            if (i == 1) w += cb[j];
            Wp[sidx(i, j)] = w;
        } else {
            Wp[sidx(i, j)] = fma(cb[i], H[j], fma(cb[j], H[i], cy.W[sidx(i, j)]));
        }
    }
    // ---- reverse of the prediction  Pp = J P J^T + Sigma,  mp = f(m)
    CGP_UNROLL for (int r = 0; r < V; r++) cy.qb += Wp[sidx(r, r)];
    cy.sb00 += Wp[sidx(V, V)]; cy.sb01 += Wp[sidx(V + 1, V)]; cy.sb11 += Wp[sidx(V + 1, V + 1)];
    double Jb[D][D], T1[D][D];
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int cc = 0; cc < D; cc++) {
        if (Model::jnz(r, cc)) {                                     // Jb is needed only where J varies
            double s = Wp[sidx(r, 0)] * JP[0][cc];
            CGP_UNROLL for (int k = 1; k < D; k++) s = fma(Wp[sidx(r, k)], JP[k][cc], s);
            Jb[r][cc] = s;
        } else {
            Jb[r][cc] = 0.;
        }
    }
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int cc = 0; cc < D; cc++) { CGP_SSUM(Model::jnz(k, cc), Wp[sidx(i, k)], J[k][cc]) T1[i][cc] = sacc; }
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j <= i; j++) { CGP_SSUM(Model::jnz(k, i), J[k][i], T1[k][j]) cy.W[sidx(i, j)] = sacc; }
    if constexpr (RAW) {
        double Ap[D][D], TA[D][D];
        CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j < D; j++) {
            if (i == j) Ap[i][j] = 0.;
            else {
                const double a = (i > j) ? cy.A[aidx(i, j)] : -cy.A[aidx(j, i)];
                if constexpr (H_E1) Ap[i][j] = a + ((j == 1) ? 0.5 * cb[i] : 0.) - ((i == 1) ? 0.5 * cb[j] : 0.);
                else Ap[i][j] = fma(0.5 * cb[i], H[j], fma(-0.5 * cb[j], H[i], a));
            }
        }
        CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int cc = 0; cc < D; cc++) { CGP_SSUM(Model::jnz(k, cc), Ap[i][k], J[k][cc]) TA[i][cc] = sacc; }
        CGP_UNROLL for (int i = 1; i < D; i++) CGP_UNROLL for (int j = 0; j < i; j++) { CGP_SSUM(Model::jnz(k, i), J[k][i], TA[k][j]) cy.A[aidx(i, j)] = sacc; }
    }
    CGP_UNROLL for (int i = 0; i < D; i++) { CGP_SSUM(Model::jnz(k, i), J[k][i], mpb[k]) cy.mb[i] = sacc; }
    // model-specific part: second derivatives of the mean and cotangents of the constants (see cgp_nll.cu)
    const double sg = tr.sg;
    const double w1 = (kTwoPi * sg) * mdl.fs, w2 = (kTwoPi * (sg * (1. - sg))) * mdl.fs;
    double esum = 0.;
    CGP_UNROLL for (int k = 0; k < NH; k++) {
        const int a = 2 * k, c2 = 2 * k + 1;
        const double dtk = mdl.dt * (double)(k + 1);
        const double th1 = dtk * w1, th2 = dtk * w2;
        const double ce = J[a][a], se = J[c2][a];
        cy.mb[a] += Jb[a][V] * (-se * th1) + Jb[c2][V] * (ce * th1);
        cy.mb[c2] += Jb[a][V] * (-ce * th1) + Jb[c2][V] * (-se * th1);
        cy.mb[V] += (Jb[a][a] * (-se) + Jb[a][c2] * (-ce) + Jb[c2][a] * ce + Jb[c2][c2] * (-se)) * th1
                    + Jb[a][V] * (-mp[a] * th1 * th1 - mp[c2] * th2) + Jb[c2][V] * (-mp[c2] * th1 * th1 + mp[a] * th2);
        esum += mpb[a] * mp[a] + mpb[c2] * mp[c2] + Jb[a][a] * J[a][a] + Jb[a][c2] * J[a][c2] + Jb[c2][a] * J[c2][a]
                + Jb[c2][c2] * J[c2][c2] + Jb[a][V] * J[a][V] + Jb[c2][V] * J[c2][V];
    }
    cy.eb += esum;                                                   // divided by e once, at the very end
    cy.fb[0] += mpb[V] * m[V] + Jb[V][V];
    cy.fb[1] += mpb[V] * m[V + 1] + Jb[V][V + 1];
    cy.fb[2] += mpb[V + 1] * m[V] + Jb[V + 1][V];
    cy.fb[3] += mpb[V + 1] * m[V + 1] + Jb[V + 1][V + 1];
}

// Named barriers of the PAIR variant (ids are immediates: a register id would reserve all 16 barriers of the CTA).
#define CGP_NB(op, id) asm volatile(op " " #id ", 64;" ::: "memory")
CGP_DEV void pair_full_arrive(int slot) { if (slot) CGP_NB("bar.arrive", 2); else CGP_NB("bar.arrive", 1); }
CGP_DEV void pair_full_sync(int slot) { if (slot) CGP_NB("bar.sync", 2); else CGP_NB("bar.sync", 1); }
CGP_DEV void pair_empty_arrive(int slot) { if (slot) CGP_NB("bar.arrive", 4); else CGP_NB("bar.arrive", 3); }
CGP_DEV void pair_empty_sync(int slot) { if (slot) CGP_NB("bar.sync", 4); else CGP_NB("bar.sync", 3); }
CGP_DEV void pair_ticket_sync() { CGP_NB("bar.sync", 5); }
#undef CGP_NB

// PAIR (few chains: every SM sub-partition would run a single warp, which cannot hide its own dependency stalls): the CTA is
// a PAIR of warps working on the same unit.  Warp 0 recomputes segment after segment into two alternating scratch slots, warp
// 1 sweeps them in reverse -- the two halves of the adjoint are independent dependency chains, and on separate warps the
// hardware interleaves them (writing both into one loop body of one warp did not: profiles/r2_nll_sweeps.txt).  Hand-over:
// one FULL and one EMPTY named barrier per slot (arrive by one warp, sync by the other, strictly alternating).
template <int NH, bool H_E1, bool RAW, bool PAIR>
__global__ void __launch_bounds__(PAIR ? 64 : 128, (NH == 1 ? (PAIR ? 6 : 3) : 1))
bwd_kernel(const CgpProblem p, const double *__restrict__ ys, const double *__restrict__ nll_bar, const double *__restrict__ ckpt,
           double *__restrict__ scratch, double *__restrict__ carry, const Geo g, unsigned *__restrict__ sched_words,
           double *__restrict__ consts_bar, double *__restrict__ m0_bar, double *__restrict__ P0_bar, double *__restrict__ Xi_bar) {
    using Sh = Shape<NH>;
    using Model = ModelLCD<NH>;
    constexpr int D = Sh::D, V = Sh::V, NS = Sh::NS, CK = Sh::CK, REC = Sh::REC, NCARRY = RAW ? Sh::CARRY_RAW : Sh::CARRY_SYM;
    const int lane = threadIdx.x & 31;
    const Sched sched{sched_words};
    const unsigned long long total = (unsigned long long)g.nunits * (unsigned long long)g.nchains;
    const int64_t slot_doubles = (int64_t)g.every * REC * kLanes;
    // scratch slots: one per warp; the PAIR variant uses the two slots of its CTA's two warps alternately
    const int64_t worker = PAIR ? (int64_t)blockIdx.x * 2 : ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double *__restrict__ slot0 = scratch + worker * slot_doubles + lane;
    __shared__ unsigned long long pair_ticket;
    __shared__ int pair_rec_warp;
    int role = 0;                                                  // PAIR: 0 = recompute, 1 = reverse
    if constexpr (PAIR) {
        // Hardware warp slot w runs on sub-partition w % 4 and the two warps of a CTA take adjacent slots: with warp 0 always
        // the recompute warp, all recompute warps of an SM would sit on sub-partitions 0 and 2 and all (FP64-pipe hungry)
        // reverse warps on 1 and 3 (measured: reverse warps alone 3.9 ms, twice their time when spread).  Alternate the
        // roles between neighbouring CTAs so that every sub-partition gets one warp of each kind.
        if (threadIdx.x == 0) {
            unsigned wid;
            asm("mov.u32 %0, %%warpid;" : "=r"(wid));
            pair_rec_warp = (int)(((wid >> 2) ^ wid) & 1u);
        }
        __syncthreads();
        role = ((int)(threadIdx.x >> 5) == pair_rec_warp) ? 0 : 1;
        if (role == 1) { pair_empty_arrive(0); pair_empty_arrive(1); }    // both slots start empty
    }
    double H[D];
    CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
    for (;;) {
        unsigned long long n;
        if constexpr (PAIR) {
            if (role == 0 && lane == 0) pair_ticket = atomicAdd(reinterpret_cast<unsigned long long *>(sched_words), 1ull);
            pair_ticket_sync();
            n = pair_ticket;
            pair_ticket_sync();                                    // both warps have read it before it is overwritten
        } else {
            n = sched.ticket(lane);
        }
        if (n >= total) break;
        const int round = (int)(n / (unsigned)g.nchains), c = (int)(n % (unsigned)g.nchains);
        const int s_hi = g.nseg - 1 - round * g.spu, s_lo = (s_hi - g.spu + 1 > 0) ? s_hi - g.spu + 1 : 0;
        const int64_t b = (int64_t)c * g.lanes + lane;
        const bool live = lane < g.lanes && b < p.B;
        const int64_t bb = live ? b : ((int64_t)c * g.lanes < p.B ? (int64_t)c * g.lanes : p.B - 1);
        Model mdl;
        mdl.load(p.consts + bb * p.consts_stride, p.dt);
        const double lw = nll_bar ? nll_bar[bb] : 1.;
        const double *__restrict__ yrow = ys + (bb / p.ys_repeat) * p.T;
        const double Xi = p.Xi;

        auto seg_len = [&](int s) { const int64_t t0 = (int64_t)s * g.every; return (int)((p.T - t0 < g.every) ? (p.T - t0) : g.every); };
        auto rec_load = [&](int s, double (&m)[D], double (&P)[NS]) {
            if (s == 0) {
                load_vec<D>(p.m0 + bb * p.m0_stride, m);
                load_sym<D>(p.P0 + bb * p.P0_stride, P);
            } else {
                const double *q = ckpt + ((int64_t)s * g.nchains + c) * (CK * kLanes) + lane;
                CGP_UNROLL for (int i = 0; i < D; i++) m[i] = __ldcg(q + i * kLanes);
                CGP_UNROLL for (int i = 0; i < NS; i++) P[i] = __ldcg(q + (D + i) * kLanes);
            }
        };
        // forward recomputation: record j = the inputs (m, P) of step j and its transcendental results, then advance
        auto rec_step = [&](const bool advance, int j, const double *__restrict__ y, double (&m)[D], double (&P)[NS], double *__restrict__ slot) {
            Lin<NH> tr;
            trig<NH>(mdl, m[V], tr);
            double *q = slot + (int64_t)j * (REC * kLanes);
            CGP_UNROLL for (int i = 0; i < D; i++) q[i * kLanes] = m[i];
            CGP_UNROLL for (int i = 0; i < NS; i++) q[(D + i) * kLanes] = P[i];
            CGP_UNROLL for (int k = 0; k < NH; k++) { q[(D + NS + 2 * k) * kLanes] = tr.ce[k]; q[(D + NS + 2 * k + 1) * kLanes] = tr.se[k]; }
            q[(D + NS + 2 * NH) * kLanes] = tr.sg;
            if (advance) step_from<NH, H_E1, false>(mdl, tr, H, Xi, __ldg(y + j), m, P);
        };
        auto rev_step = [&](int j, const double *__restrict__ y, const double *__restrict__ slot, Carry<NH, RAW> &cy) {
            const double *q = slot + (int64_t)j * (REC * kLanes);
            double m[D], P[NS];
            Lin<NH> tr;
            CGP_UNROLL for (int i = 0; i < D; i++) m[i] = q[i * kLanes];
            CGP_UNROLL for (int i = 0; i < NS; i++) P[i] = q[(D + i) * kLanes];
            CGP_UNROLL for (int k = 0; k < NH; k++) { tr.ce[k] = q[(D + NS + 2 * k) * kLanes]; tr.se[k] = q[(D + NS + 2 * k + 1) * kLanes]; }
            tr.sg = q[(D + NS + 2 * NH) * kLanes];
            reverse_step<NH, H_E1, RAW>(mdl, H, Xi, __ldg(y + j), lw, m, P, tr, cy);
        };
        auto recompute = [&](int s, double *__restrict__ slot) {
            double m[D], P[NS];
            rec_load(s, m, P);
            const int nst = seg_len(s);
            const double *__restrict__ y = yrow + (int64_t)s * g.every;
            for (int j = 0; j + 1 < nst; j++) rec_step(true, j, y, m, P, slot);
            rec_step(false, nst - 1, y, m, P, slot);
        };
        auto load_carry = [&](Carry<NH, RAW> &cy) -> bool {
            if (round == 0) { cy.zero(); return true; }
            if (!sched.wait(c, (unsigned)round, lane)) return false;
            const double *q = carry + (int64_t)c * (NCARRY * kLanes) + lane;
            cy.each([&](int k, double &x) { x = __ldcg(q + k * kLanes); });
            return true;
        };

        Carry<NH, RAW> cy;
        if constexpr (PAIR) {
            if (role == 0) {
                for (int s = s_hi; s >= s_lo; s--) {
                    const int sl = (s_hi - s) & 1;
                    pair_empty_sync(sl);
                    recompute(s, slot0 + sl * slot_doubles);
                    __threadfence_block();
                    pair_full_arrive(sl);
                }
                continue;                                          // the reverse warp finishes the unit
            }
            const bool ok = load_carry(cy);                        // an aborted wait still drains the hand-over below
            for (int s = s_hi; s >= s_lo; s--) {
                const int sl = (s_hi - s) & 1;
                pair_full_sync(sl);
                const double *__restrict__ y = yrow + (int64_t)s * g.every;
                for (int j = seg_len(s) - 1; j >= 0; j--) rev_step(j, y, slot0 + sl * slot_doubles, cy);
                pair_empty_arrive(sl);
            }
            if (!ok) continue;
        } else {
            for (int s = s_hi; s >= s_lo; s--) {
                recompute(s, slot0);                       // independent of the chain's reverse state: runs before the wait
                if (s == s_hi && !load_carry(cy)) return;
                const double *__restrict__ y = yrow + (int64_t)s * g.every;
                for (int j = seg_len(s) - 1; j >= 0; j--) rev_step(j, y, slot0, cy);
            }
        }
        if (s_lo > 0) {
            double *q = carry + (int64_t)c * (NCARRY * kLanes) + lane;
            cy.each([&](int k, double &x) { __stcg(q + k * kLanes, x); });
            sched.publish(c, (unsigned)(round + 1), lane);
        } else if (live) {
            double *cb = consts_bar + b * CGP_NC_LCD;
            cb[0] = cy.eb / mdl.e; cb[1] = cy.fb[0]; cb[2] = cy.fb[1]; cb[3] = cy.fb[2]; cb[4] = cy.fb[3];
            cb[5] = 0.5 * cy.qb; cb[6] = 0.5 * cy.sb00; cb[7] = cy.sb01; cb[8] = 0.5 * cy.sb11; cb[9] = 0.;
            CGP_UNROLL for (int i = 0; i < D; i++) m0_bar[b * D + i] = cy.mb[i];
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int cc = 0; cc < D; cc++) {
                double x = 0.5 * cy.W[sidx(r, cc)];
                if constexpr (RAW) { if (r > cc) x += cy.A[aidx(r, cc)]; else if (r < cc) x -= cy.A[aidx(cc, r)]; }
                P0_bar[b * D * D + r * D + cc] = x;
            }
            if (Xi_bar) Xi_bar[b] = cy.xib;
        }
    }
}
#undef CGP_SSUM

// ------------------------------------------------------------------------------------------------ host side
struct Plan {
    int ck, rec, ncarry_max, sms;
    int64_t every, nseg, lanes, nchains, workers_max;
    size_t off_sched_f, off_sched_b, off_ckpt, off_carry, off_scratch, total;
};

static int device_sms() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) {
        cudaGetLastError();
        sms = 148;                                     // B200; also the answer on a box without a GPU (size queries only)
    }
    return sms;
}
constexpr int kMaxBlocksPerSM = 5;                     // upper bound on persistent 4-warp blocks per SM over all kernel variants

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static int env_int(const char *name, int dflt) {      // tuning knobs for profiles/scripts (not part of the ABI)
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// lanes per chain: full warps unless they would leave SM sub-partitions without any warp; then half / quarter warps put the
// problems on more sub-partitions.  (Narrow chains never pay once every sub-partition has a warp: the FP64 pipe takes 2 cycles
// per warp instruction whatever the number of active lanes -- profiles/microbench/fp64_lanes.cu -- and 20 000 problems ran
// 7.0 / 9.8 / 18.0 ms as chains of 32 / 16 / 8, profiles/r2_nll_sweeps.txt.)
static int64_t pick_lanes(int64_t B, int sms) {
    const int forced = env_int("CGP_NLL_LANES", 0);
    if (forced == 8 || forced == 16 || forced == 32) return forced;
    const int64_t smsp = (int64_t)sms * 4;
    if ((B + 31) / 32 >= smsp) return 32;
    if ((B + 15) / 16 >= smsp) return 16;
    return (B + 15) / 16 * 2 > smsp ? 16 : 8;
}

static Plan make_plan(const CgpProblem &p, int64_t every) {
    Plan pl{};
    const int d = p.d, ns = d * (d + 1) / 2;
    pl.sms = device_sms();
    pl.ck = d + ns + 1;
    pl.rec = d + ns + 2 * p.num_harmonics + 1;
    pl.ncarry_max = d + ns + d * (d - 1) / 2 + 10;
    pl.every = every;
    pl.nseg = (p.T + every - 1) / every;
    pl.lanes = pick_lanes(p.B, pl.sms);
    pl.nchains = (p.B + pl.lanes - 1) / pl.lanes;
    const int64_t cap = (int64_t)pl.sms * kMaxBlocksPerSM * 4;
    const int64_t chains4 = (pl.nchains + 3) / 4 * 4;          // blocks hold 4 warps
    pl.workers_max = chains4 < cap ? chains4 : cap;            // scratch slots = persistent warps ...
    const int64_t pairs2 = 2 * (pl.nchains < (int64_t)pl.sms * 4 ? pl.nchains : (int64_t)pl.sms * 4);
    if (pl.workers_max < pairs2) pl.workers_max = pairs2;      // ... or two per recompute / reverse pair
    const size_t sched_bytes = align_up((size_t)(4 + pl.nchains) * sizeof(unsigned), 256);
    size_t off = 0;
    pl.off_sched_f = off; off += sched_bytes;
    pl.off_sched_b = off; off += sched_bytes;
    pl.off_ckpt = off;    off += align_up((size_t)pl.nseg * pl.nchains * pl.ck * kLanes * sizeof(double), 256);
    pl.off_carry = off;   off += align_up((size_t)pl.nchains * pl.ncarry_max * kLanes * sizeof(double), 256);
    pl.off_scratch = off; off += align_up((size_t)pl.workers_max * every * pl.rec * kLanes * sizeof(double), 256);
    pl.total = off;
    return pl;
}

// persistent grid: whole multiples of the SM count while there are at least that many chains, so that every SM
// sub-partition runs the same number of warps (blocks of 4 warps, one per sub-partition)
static int grid_blocks(const Plan &pl, int max_blocks_per_sm) {
    const int64_t blocks_needed = (pl.nchains + 3) / 4;
    if (blocks_needed <= pl.sms) return (int)blocks_needed;
    int64_t k = pl.nchains / ((int64_t)pl.sms * 4);
    if (k < 1) k = 1;
    if (k > max_blocks_per_sm) k = max_blocks_per_sm;
    return (int)(k * pl.sms);
}
// segments per unit: ~`steps` time steps between two visits to the ticket counter, but at least ~8 units per chain
static Geo make_geo(const Plan &pl, int steps) {
    Geo g{};
    g.every = (int)pl.every; g.nseg = (int)pl.nseg; g.nchains = (int)pl.nchains; g.lanes = (int)pl.lanes;
    int64_t spu = steps / pl.every;
    if (spu > pl.nseg / 8) spu = pl.nseg / 8;
    if (spu < 1) spu = 1;
    g.spu = (int)spu;
    g.nunits = (int)((pl.nseg + spu - 1) / spu);
    return g;
}

}  // namespace nll2

using namespace nll2;

size_t nll2_workspace_bytes(const CgpProblem &p, int64_t every) { return make_plan(p, every).total; }

bool nll2_supported(const CgpProblem &p) {
    return p.model == CGP_MODEL_LCD && p.num_harmonics >= 1 && p.num_harmonics <= 3 && p.d == 2 * p.num_harmonics + 2;
}

template <int NH, bool H_E1>
static int fwd_launch(const CgpProblem &p, const double *ys, double *nll, char *ws, const Plan &pl, cudaStream_t s) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fwd_kernel<NH, H_E1>, 128, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;                            // measured: 3 / 4 / 5 blocks per SM -> 10.6 / 10.1 / 11.2 ms
    per_sm = env_int("CGP_NLL_KF", per_sm);
    if (per_sm > kMaxBlocksPerSM) per_sm = kMaxBlocksPerSM;
    const int grid = grid_blocks(pl, per_sm);
    if (!ws) {
        Geo g{(int)p.T, 1, (int)pl.nchains, (int)pl.lanes, 1, 1};
        fwd_kernel<NH, H_E1><<<grid, 128, 0, s>>>(p, ys, nll, nullptr, g, nullptr);
        return check_launch();
    }
    unsigned *sched = reinterpret_cast<unsigned *>(ws + pl.off_sched_f);
    cudaError_t e = cudaMemsetAsync(sched, 0, (size_t)(4 + pl.nchains) * sizeof(unsigned), s);
    if (e != cudaSuccess) return (int)e;
    fwd_kernel<NH, H_E1><<<grid, 128, 0, s>>>(p, ys, nll, reinterpret_cast<double *>(ws + pl.off_ckpt),
                                              make_geo(pl, env_int("CGP_NLL_UNIT_F", 256)), sched);
    return check_launch();
}

int nll2_fwd(const CgpProblem &p, const double *ys, double *nll, void *workspace, int64_t every, cudaStream_t s) {
    const Plan pl = make_plan(p, workspace ? every : p.T);
    if (pl.nseg > 0x7fffffff || pl.nchains > 0x3fffffff || pl.every > 0x7fffffff) return CGP_ERR_UNSUPPORTED;
    // unfinished chains (abort) must not look like results
    cudaError_t e = cudaMemsetAsync(nll, 0xff, (size_t)p.B * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
    const bool e1 = p.h_unit_index == 1;
    char *ws = static_cast<char *>(workspace);
    switch (p.num_harmonics) {
        case 1: return e1 ? fwd_launch<1, true>(p, ys, nll, ws, pl, s) : fwd_launch<1, false>(p, ys, nll, ws, pl, s);
        case 2: return fwd_launch<2, false>(p, ys, nll, ws, pl, s);
        case 3: return fwd_launch<3, false>(p, ys, nll, ws, pl, s);
        default: return CGP_ERR_UNSUPPORTED;
    }
}

template <int NH, bool H_E1, bool RAW, bool PAIR>
static int bwd_launch_v(const CgpProblem &p, const double *ys, const double *nll_bar, char *ws, const Plan &pl, double *consts_bar,
                        double *m0_bar, double *P0_bar, double *Xi_bar, cudaStream_t s) {
    constexpr int kBlockThreads = PAIR ? 64 : 128;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bwd_kernel<NH, H_E1, RAW, PAIR>, kBlockThreads, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    int grid;
    if (PAIR) {
        // one pair per chain, at most one pair per SM sub-partition
        const int64_t cap = (int64_t)pl.sms * 4;
        grid = (int)(pl.nchains < cap ? pl.nchains : cap);
        if ((int64_t)grid * 2 > pl.workers_max) return CGP_ERR_WORKSPACE;
    } else {
        if (per_sm > kMaxBlocksPerSM) per_sm = kMaxBlocksPerSM;
        per_sm = env_int("CGP_NLL_KB", per_sm);
        if (per_sm > kMaxBlocksPerSM) per_sm = kMaxBlocksPerSM;
        grid = grid_blocks(pl, per_sm);
        if ((int64_t)grid * 4 > pl.workers_max) return CGP_ERR_WORKSPACE;  // one scratch slot per persistent warp
    }
    unsigned *sched = reinterpret_cast<unsigned *>(ws + pl.off_sched_b);
    cudaError_t e = cudaMemsetAsync(sched, 0, (size_t)(4 + pl.nchains) * sizeof(unsigned), s);
    if (e != cudaSuccess) return (int)e;
    bwd_kernel<NH, H_E1, RAW, PAIR><<<grid, kBlockThreads, 0, s>>>(
        p, ys, nll_bar, reinterpret_cast<const double *>(ws + pl.off_ckpt), reinterpret_cast<double *>(ws + pl.off_scratch),
        reinterpret_cast<double *>(ws + pl.off_carry), make_geo(pl, env_int("CGP_NLL_UNIT_B", PAIR ? 512 : 128)), sched, consts_bar,
        m0_bar, P0_bar, Xi_bar);
    return check_launch();
}
template <int NH, bool H_E1, bool RAW>
static int bwd_launch(const CgpProblem &p, const double *ys, const double *nll_bar, char *ws, const Plan &pl, double *consts_bar,
                      double *m0_bar, double *P0_bar, double *Xi_bar, cudaStream_t s) {
    if constexpr (NH == 1) {
        // fewer than 1.5 chains per SM sub-partition: recompute / reverse warp pairs
        if (env_int("CGP_NLL_PAIR", 2 * pl.nchains < 3 * (int64_t)pl.sms * 4 ? 1 : 0) != 0)
            return bwd_launch_v<NH, H_E1, RAW, true>(p, ys, nll_bar, ws, pl, consts_bar, m0_bar, P0_bar, Xi_bar, s);
    }
    return bwd_launch_v<NH, H_E1, RAW, false>(p, ys, nll_bar, ws, pl, consts_bar, m0_bar, P0_bar, Xi_bar, s);
}

int nll2_bwd(const CgpProblem &p, const double *ys, const double *nll_bar, void *workspace, int64_t every, bool raw_p0_bar,
             double *consts_bar, double *m0_bar, double *P0_bar, double *Xi_bar, cudaStream_t s) {
    const Plan pl = make_plan(p, every);
    if (pl.nseg > 0x7fffffff || pl.nchains > 0x3fffffff || pl.every > 0x7fffffff) return CGP_ERR_UNSUPPORTED;
    cudaError_t e = cudaMemsetAsync(consts_bar, 0xff, (size_t)p.B * CGP_NC_LCD * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
    const bool e1 = p.h_unit_index == 1;
    char *ws = static_cast<char *>(workspace);
#define CGP_BWD(NH, E1)                                                                                                   \
    (raw_p0_bar ? bwd_launch<NH, E1, true>(p, ys, nll_bar, ws, pl, consts_bar, m0_bar, P0_bar, Xi_bar, s)                \
                : bwd_launch<NH, E1, false>(p, ys, nll_bar, ws, pl, consts_bar, m0_bar, P0_bar, Xi_bar, s))
    switch (p.num_harmonics) {
        case 1: return e1 ? CGP_BWD(1, true) : CGP_BWD(1, false);
        case 2: return CGP_BWD(2, false);
        case 3: return CGP_BWD(3, false);
        default: return CGP_ERR_UNSUPPORTED;
    }
#undef CGP_BWD
}

}  // namespace cgp
