// cgp_nll.cu -- MLE path: EKF negative log-likelihood without per-step outputs and its hand-written reverse-mode
// adjoint (what jax.grad(lambda th: ekf(...)[-1][-1]) computes in demos/ekfs_mle.py:42-49 by differentiating
// through lax.scan).
//
// One thread owns one problem (chirp x hyper-parameter candidate).  The forward sweep stores a checkpoint of
// (m, P) every `ckpt_every` steps; the adjoint walks the segments backwards: it re-runs a segment forward from its
// checkpoint, keeping the inputs (m, P) of every step in a per-thread scratch slice, then sweeps the segment in
// reverse.  Checkpoints and scratch are laid out [..][20][B] (problem index fastest) so that a warp's accesses
// coalesce.  With ckpt_every ~ sqrt(T) the memory is O(sqrt(T)) per problem (T = 1e5: ~100 KB).
//
// Cotangents are returned w.r.t. the kernel's inputs -- the derived model constants (e, F, q, S), m0, P0 (a general,
// unsymmetrised matrix cotangent: JAX's convention) and Xi; the map theta -> constants stays in host autodiff
// (chirpgp_b200/mle.py), which also reproduces the lam == 0 branch of models.py:302-308.
//
// One EKF step with c = Pp h, S = h^T Pp h + Xi, v = y - h^T mp:
//     m' = mp + c v / S,   P' = Pp - c c^T / S,   l += (log(2 pi S) + v^2 / S) / 2        (filters_smoothers.py:55-68)
// reverse (incoming mb = dL/dm', Pb = dL/dP', lw = dL/dl):
//     Sb  = lw (1/S - v^2/S^2)/2 - (mb.c) v/S^2 + (c^T Pb c)/S^2
//     vb  = lw v/S + (mb.c)/S
//     cb  = mb v/S - (Pb + Pb^T) c/S + Sb h
//     mpb = mb - vb h,   Ppb = Pb + cb h^T,   Xib += Sb
//     Pb' = J^T Ppb J,   Sigmab += Ppb,   Jb = Ppb J P^T + Ppb^T J P
//     mb' = J^T mpb + sum_ij Jb_ij dJ_ij/dm          (second derivatives of the model mean)
#include "cgp_dispatch.cuh"

namespace cgp {

template <int D> struct NllLayout {
    static constexpr int REC = D + D * D;
};

template <int D>
CGP_DEV void save_state(double *__restrict__ base, int64_t B, int64_t b, const double (&m)[D], const double (&P)[D][D]) {
    CGP_UNROLL for (int i = 0; i < D; i++) base[(int64_t)i * B + b] = m[i];
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) base[(int64_t)(D + r * D + c) * B + b] = P[r][c];
}
template <int D>
CGP_DEV void load_state(const double *__restrict__ base, int64_t B, int64_t b, double (&m)[D], double (&P)[D][D]) {
    CGP_UNROLL for (int i = 0; i < D; i++) m[i] = base[(int64_t)i * B + b];
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) P[r][c] = base[(int64_t)(D + r * D + c) * B + b];
}

// forward EKF step, identical arithmetic to ekf_thread_kernel
template <class Model>
CGP_DEV double ekf_step(const Model &mdl, const double (&H)[Model::D], double Xi, double y, double (&m)[Model::D],
                        double (&P)[Model::D][Model::D]) {
    constexpr int D = Model::D;
    double mp[D], J[D][D], JP[D][D], Pp[D][D];
    mdl.mean_jac(m, mp, J);
    jmul<Model, D>(J, P, JP);
    mul_jt<Model, D>(JP, J, Pp);
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++)
        if (Model::has_sig(r, c)) Pp[r][c] += mdl.sig(r, c);
    return linear_update<D>(mp, Pp, H, Xi, y, m, P);
}

template <int NH>
__global__ void __launch_bounds__(128) ekf_nll_fwd_kernel(const CgpProblem p, const double *__restrict__ ys, double *__restrict__ nll,
                                                          double *__restrict__ ckpt, int64_t ckpt_every,
                                                          double *__restrict__ nell_path /* [B, T] cumulative, or NULL */) {
    using Model = ModelLCD<NH>;
    constexpr int D = Model::D, REC = NllLayout<D>::REC;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    double m[D], P[D][D], H[D];
    load_vec<D>(p.m0 + b * p.m0_stride, m);
    load_mat<D>(p.P0 + b * p.P0_stride, P);
    CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
    const double *__restrict__ y = ys + (b / p.ys_repeat) * p.T;
    double acc = 0.;
    int64_t seg = 0, left = 0;
    NellRowWriter pathw;
    if (nell_path) pathw.init(nell_path, b, p.T);
    for (int64_t t = 0; t < p.T; t++) {
        if (left == 0) {
            if (ckpt) save_state<D>(ckpt + seg * REC * p.B, p.B, b, m, P);
            seg++;
            left = ckpt_every;
        }
        left--;
        acc = acc + ekf_step<Model>(mdl, H, p.Xi, __ldg(y + t), m, P);
        if (nell_path) pathw.put(t, p.T, acc);
    }
    if (nll) nll[b] = acc;
}

template <int NH>
__global__ void __launch_bounds__(128, NH == 1 ? 3 : 1) ekf_nll_bwd_kernel(const CgpProblem p, const double *__restrict__ ys,
                                                          const double *__restrict__ nll_bar,
                                                          const double *__restrict__ step_w /* [B, T] weight of every increment, or NULL */,
                                                          const double *__restrict__ ckpt,
                                                          double *__restrict__ scratch, int64_t ckpt_every,
                                                          double *__restrict__ consts_bar, double *__restrict__ m0_bar,
                                                          double *__restrict__ P0_bar, double *__restrict__ Xi_bar) {
    using Model = ModelLCD<NH>;
    constexpr int D = Model::D, V = Model::V, REC = NllLayout<D>::REC;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const int64_t B = p.B, T = p.T;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    double H[D];
    CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
    const double *__restrict__ y = ys + (b / p.ys_repeat) * T;
    const double lw_all = nll_bar ? nll_bar[b] : (step_w ? 0. : 1.);
    const double Xi = p.Xi;
    double mb[D], Pb[D][D];
    CGP_UNROLL for (int i = 0; i < D; i++) mb[i] = 0.;
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) Pb[r][c] = 0.;
    double eb = 0., fb[4] = {0., 0., 0., 0.}, qb = 0., sb00 = 0., sb01 = 0., sb11 = 0., xib = 0.;
    const int64_t nseg = (T + ckpt_every - 1) / ckpt_every;
    for (int64_t seg = nseg - 1; seg >= 0; seg--) {
        const int64_t t0 = seg * ckpt_every;
        const int n = (int)((T - t0 < ckpt_every) ? (T - t0) : ckpt_every);
        {   // forward recomputation of the segment: scratch[j] = inputs (m, P) of step t0 + j
            double m[D], P[D][D];
            load_state<D>(ckpt + seg * REC * B, B, b, m, P);
            for (int j = 0; j < n; j++) {
                save_state<D>(scratch + (int64_t)j * REC * B, B, b, m, P);
                if (j + 1 < n) ekf_step<Model>(mdl, H, Xi, __ldg(y + t0 + j), m, P);
            }
        }
        for (int j = n - 1; j >= 0; j--) {
            double m[D], P[D][D];
            load_state<D>(scratch + (int64_t)j * REC * B, B, b, m, P);
            const double yt = __ldg(y + t0 + j);
            const double lw = step_w ? lw_all + __ldg(step_w + b * T + t0 + j) : lw_all;     // dL / d(increment of this step)
            // ---- recompute the forward quantities of this step
            double mp[D], J[D][D], JP[D][D], Pp[D][D];
            mdl.mean_jac(m, mp, J);
            jmul<Model, D>(J, P, JP);
            mul_jt<Model, D>(JP, J, Pp);
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++)
                if (Model::has_sig(r, c)) Pp[r][c] += mdl.sig(r, c);
            double cv[D], S = 0., pred = 0.;
            CGP_UNROLL for (int i = 0; i < D; i++) {
                double s = Pp[i][0] * H[0];
                CGP_UNROLL for (int k = 1; k < D; k++) s = fma(Pp[i][k], H[k], s);
                cv[i] = s;
            }
            CGP_UNROLL for (int jj = 0; jj < D; jj++) {
                double hp = H[0] * Pp[0][jj];
                CGP_UNROLL for (int i = 1; i < D; i++) hp = fma(H[i], Pp[i][jj], hp);
                S = fma(hp, H[jj], S);
            }
            S += Xi;
            CGP_UNROLL for (int i = 0; i < D; i++) pred = fma(H[i], mp[i], pred);
            const double v = yt - pred, iS = 1. / S, iS2 = iS * iS;
            // ---- reverse of the measurement update
            double mc = 0., cPc = 0., PPc[D];
            CGP_UNROLL for (int i = 0; i < D; i++) mc = fma(mb[i], cv[i], mc);
            CGP_UNROLL for (int i = 0; i < D; i++) {
                double s = 0., s2 = 0.;
                CGP_UNROLL for (int k = 0; k < D; k++) { s = fma(Pb[i][k] + Pb[k][i], cv[k], s); s2 = fma(Pb[i][k], cv[k], s2); }
                PPc[i] = s;
                cPc = fma(cv[i], s2, cPc);
            }
            const double Sb = lw * 0.5 * (iS - v * v * iS2) - mc * v * iS2 + cPc * iS2;
            const double vb = lw * v * iS + mc * iS;
            xib += Sb;
            double mpb[D], Ppb[D][D];
            CGP_UNROLL for (int i = 0; i < D; i++) {
                const double cb = mb[i] * v * iS - PPc[i] * iS + Sb * H[i];
                mpb[i] = mb[i] - vb * H[i];
                CGP_UNROLL for (int k = 0; k < D; k++) Ppb[i][k] = fma(cb, H[k], Pb[i][k]);
            }
            // ---- reverse of the prediction  Pp = J P J^T + Sigma,  mp = f(m)
            CGP_UNROLL for (int r = 0; r < V; r++) qb += Ppb[r][r];
            sb00 += Ppb[V][V]; sb01 += Ppb[V][V + 1] + Ppb[V + 1][V]; sb11 += Ppb[V + 1][V + 1];
            double JPt[D][D], Jb[D][D], T1[D][D];
            jmul_nt<Model, D>(J, P, JPt);                              // J P^T
            CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) {
                double s = 0.;
                CGP_UNROLL for (int k = 0; k < D; k++) s = fma(Ppb[r][k], JPt[k][c], fma(Ppb[k][r], JP[k][c], s));
                Jb[r][c] = s;                                          // Ppb J P^T + Ppb^T J P
            }
            jtmul<Model, D>(J, Ppb, T1);                               // T1 = J^T Ppb
            mul_j<Model, D>(T1, J, Pb);                                // Pb <- J^T Ppb J
            jtvec<Model, D>(J, mpb, mb);                               // mb <- J^T mpb (+ second-order terms below)
            // model-specific part: second derivatives of the mean and cotangents of the constants
            double gv, sg;
            softplus_and_sigmoid(m[V], gv, sg);
            const double w1 = (kTwoPi * sg) * mdl.fs, w2 = (kTwoPi * (sg * (1. - sg))) * mdl.fs;
            double esum = 0.;
            CGP_UNROLL for (int k = 0; k < NH; k++) {
                const int a = 2 * k, c2 = 2 * k + 1;
                const double dtk = mdl.dt * (double)(k + 1);
                const double th1 = dtk * w1, th2 = dtk * w2;
                const double ce = J[a][a], se = J[c2][a];
                mb[a] += Jb[a][V] * (-se * th1) + Jb[c2][V] * (ce * th1);
                mb[c2] += Jb[a][V] * (-ce * th1) + Jb[c2][V] * (-se * th1);
                mb[V] += (Jb[a][a] * (-se) + Jb[a][c2] * (-ce) + Jb[c2][a] * ce + Jb[c2][c2] * (-se)) * th1
                         + Jb[a][V] * (-mp[a] * th1 * th1 - mp[c2] * th2) + Jb[c2][V] * (-mp[c2] * th1 * th1 + mp[a] * th2);
                esum += mpb[a] * mp[a] + mpb[c2] * mp[c2] + Jb[a][a] * J[a][a] + Jb[a][c2] * J[a][c2] + Jb[c2][a] * J[c2][a]
                        + Jb[c2][c2] * J[c2][c2] + Jb[a][V] * J[a][V] + Jb[c2][V] * J[c2][V];
            }
            eb += esum / mdl.e;
            fb[0] += mpb[V] * m[V] + Jb[V][V];
            fb[1] += mpb[V] * m[V + 1] + Jb[V][V + 1];
            fb[2] += mpb[V + 1] * m[V] + Jb[V + 1][V];
            fb[3] += mpb[V + 1] * m[V + 1] + Jb[V + 1][V + 1];
        }
    }
    double *cb = consts_bar + b * CGP_NC_LCD;
    cb[0] = eb; cb[1] = fb[0]; cb[2] = fb[1]; cb[3] = fb[2]; cb[4] = fb[3]; cb[5] = qb; cb[6] = sb00; cb[7] = sb01; cb[8] = sb11;
    cb[9] = 0.;
    CGP_UNROLL for (int i = 0; i < D; i++) m0_bar[b * D + i] = mb[i];
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) P0_bar[b * D * D + r * D + c] = Pb[r][c];
    if (Xi_bar) Xi_bar[b] = xib;
}

static size_t ckpt_doubles(const CgpProblem &p, int64_t every) {
    const int64_t nseg = (p.T + every - 1) / every;
    return (size_t)nseg * (size_t)(p.d + p.d * p.d) * (size_t)p.B;
}
static size_t scratch_doubles(const CgpProblem &p, int64_t every) {
    return (size_t)every * (size_t)(p.d + p.d * p.d) * (size_t)p.B;
}

// second-generation kernels (cgp_nll2.cu): persistent, ticket-scheduled, packed-symmetric covariance
size_t nll2_workspace_bytes(const CgpProblem &p, int64_t every);
bool nll2_supported(const CgpProblem &p);
int nll2_fwd(const CgpProblem &p, const double *ys, double *nll, void *workspace, int64_t every, cudaStream_t s);
int nll2_bwd(const CgpProblem &p, const double *ys, const double *nll_bar, void *workspace, int64_t every, bool raw_p0_bar,
             double *consts_bar, double *m0_bar, double *P0_bar, double *Xi_bar, cudaStream_t s);

}  // namespace cgp

using namespace cgp;

extern "C" {

int64_t cgp_ekf_nll_default_ckpt(int64_t T) {
    // segments of 16 steps keep the adjoint's per-warp scratch slots (16 x 17 x 256 bytes for d = 4, 1776 warps: 124 MB)
    // about the size of the L2; measured on 160 000 problems: 8 / 16 / 24 / 32 steps -> 19.7 / 18.5 / 16.7 / 16.1 G steps/s
    // (profiles/r2_nll_sweeps.txt).  Short series get ~sqrt(T).
    int64_t c = 1;
    while (c * c < T) c++;
    return c > 16 ? 16 : (c < 1 ? 1 : c);
}

size_t cgp_ekf_nll_workspace_bytes(const CgpProblem *p, int64_t ckpt_every) {
    if (!p || ckpt_every < 1 || !nll2_supported(*p)) return 0;
    return nll2_workspace_bytes(*p, ckpt_every);                      // persistent kernels (scalar nll)
}
size_t cgp_ekf_nll_path_workspace_bytes(const CgpProblem *p, int64_t ckpt_every) {
    if (!p || ckpt_every < 1) return 0;
    return (ckpt_doubles(*p, ckpt_every) + scratch_doubles(*p, ckpt_every)) * sizeof(double);   // thread-per-problem kernels
}

static int nll_fwd(const CgpProblem *p, const double *ys, double *nll, double *nell_path, void *workspace, size_t ws_bytes,
                   int64_t ckpt_every, void *stream) {
    if (!p || !ys || (!nll && !nell_path) || p->B < 1 || p->T < 1 || !p->consts || !p->m0 || !p->P0 || !p->H || p->ys_repeat < 1)
        return CGP_ERR_BAD_ARG;
    if (p->model != CGP_MODEL_LCD || p->d != 2 * p->num_harmonics + 2) return CGP_ERR_UNSUPPORTED;
    double *ckpt = nullptr;
    if (workspace) {
        if (ckpt_every < 1) return CGP_ERR_BAD_ARG;
        if (ws_bytes < cgp_ekf_nll_path_workspace_bytes(p, ckpt_every)) return CGP_ERR_WORKSPACE;
        ckpt = (double *)workspace;
    } else {
        ckpt_every = p->T;
    }
    const int block = 128;
    const unsigned grid = (unsigned)ceil_div(p->B, block);
    cudaStream_t s = (cudaStream_t)stream;
    switch (p->num_harmonics) {
        case 1: ekf_nll_fwd_kernel<1><<<grid, block, 0, s>>>(*p, ys, nll, ckpt, ckpt_every, nell_path); break;
        case 2: ekf_nll_fwd_kernel<2><<<grid, block, 0, s>>>(*p, ys, nll, ckpt, ckpt_every, nell_path); break;
        case 3: ekf_nll_fwd_kernel<3><<<grid, block, 0, s>>>(*p, ys, nll, ckpt, ckpt_every, nell_path); break;
        default: return CGP_ERR_UNSUPPORTED;
    }
    return check_launch();
}

int cgp_ekf_nll_fwd_f64(const CgpProblem *p, const double *ys, double *nll, void *workspace, size_t ws_bytes,
                        int64_t ckpt_every, void *stream) {
    if (!p || !ys || !nll || p->B < 1 || p->T < 1 || !p->consts || !p->m0 || !p->P0 || !p->H || p->ys_repeat < 1)
        return CGP_ERR_BAD_ARG;
    if (!nll2_supported(*p)) return CGP_ERR_UNSUPPORTED;
    if (workspace) {
        if (ckpt_every < 1) return CGP_ERR_BAD_ARG;
        if (ws_bytes < cgp_ekf_nll_workspace_bytes(p, ckpt_every)) return CGP_ERR_WORKSPACE;
    }
    return nll2_fwd(*p, ys, nll, workspace, ckpt_every, (cudaStream_t)stream);
}
int cgp_ekf_nll_path_fwd_f64(const CgpProblem *p, const double *ys, double *nell, void *workspace, size_t ws_bytes,
                             int64_t ckpt_every, void *stream) {
    return nll_fwd(p, ys, nullptr, nell, workspace, ws_bytes, ckpt_every, stream);
}

static int nll_bwd(const CgpProblem *p, const double *ys, const double *nll_bar, const double *step_w, void *workspace,
                   size_t ws_bytes, int64_t ckpt_every, double *consts_bar, double *m0_bar, double *P0_bar, double *Xi_bar,
                   void *stream) {
    if (!p || !ys || !workspace || !consts_bar || !m0_bar || !P0_bar || p->B < 1 || p->T < 1 || ckpt_every < 1 ||
        !p->consts || !p->H || p->ys_repeat < 1)
        return CGP_ERR_BAD_ARG;
    if (p->model != CGP_MODEL_LCD || p->d != 2 * p->num_harmonics + 2) return CGP_ERR_UNSUPPORTED;
    if (ws_bytes < cgp_ekf_nll_path_workspace_bytes(p, ckpt_every)) return CGP_ERR_WORKSPACE;
    double *ckpt = (double *)workspace;
    double *scratch = ckpt + ckpt_doubles(*p, ckpt_every);
    const int block = 128;
    const unsigned grid = (unsigned)ceil_div(p->B, block);
    cudaStream_t s = (cudaStream_t)stream;
    switch (p->num_harmonics) {
        case 1: ekf_nll_bwd_kernel<1><<<grid, block, 0, s>>>(*p, ys, nll_bar, step_w, ckpt, scratch, ckpt_every, consts_bar, m0_bar, P0_bar, Xi_bar); break;
        case 2: ekf_nll_bwd_kernel<2><<<grid, block, 0, s>>>(*p, ys, nll_bar, step_w, ckpt, scratch, ckpt_every, consts_bar, m0_bar, P0_bar, Xi_bar); break;
        case 3: ekf_nll_bwd_kernel<3><<<grid, block, 0, s>>>(*p, ys, nll_bar, step_w, ckpt, scratch, ckpt_every, consts_bar, m0_bar, P0_bar, Xi_bar); break;
        default: return CGP_ERR_UNSUPPORTED;
    }
    return check_launch();
}
static int nll_bwd2(const CgpProblem *p, const double *ys, const double *nll_bar, void *workspace, size_t ws_bytes,
                    int64_t ckpt_every, bool raw, double *consts_bar, double *m0_bar, double *P0_bar, double *Xi_bar, void *stream) {
    if (!p || !ys || !workspace || !consts_bar || !m0_bar || !P0_bar || p->B < 1 || p->T < 1 || ckpt_every < 1 ||
        !p->consts || !p->m0 || !p->P0 || !p->H || p->ys_repeat < 1)
        return CGP_ERR_BAD_ARG;
    if (!nll2_supported(*p)) return CGP_ERR_UNSUPPORTED;
    if (ws_bytes < cgp_ekf_nll_workspace_bytes(p, ckpt_every)) return CGP_ERR_WORKSPACE;
    return nll2_bwd(*p, ys, nll_bar, workspace, ckpt_every, raw, consts_bar, m0_bar, P0_bar, Xi_bar, (cudaStream_t)stream);
}
int cgp_ekf_nll_bwd_f64(const CgpProblem *p, const double *ys, const double *nll_bar, void *workspace, size_t ws_bytes,
                        int64_t ckpt_every, double *consts_bar, double *m0_bar, double *P0_bar, double *Xi_bar, void *stream) {
    return nll_bwd2(p, ys, nll_bar, workspace, ws_bytes, ckpt_every, true, consts_bar, m0_bar, P0_bar, Xi_bar, stream);
}
int cgp_ekf_nll_bwd_sym_f64(const CgpProblem *p, const double *ys, const double *nll_bar, void *workspace, size_t ws_bytes,
                            int64_t ckpt_every, double *consts_bar, double *m0_bar, double *P0_bar, double *Xi_bar, void *stream) {
    return nll_bwd2(p, ys, nll_bar, workspace, ws_bytes, ckpt_every, false, consts_bar, m0_bar, P0_bar, Xi_bar, stream);
}
int cgp_ekf_nll_path_bwd_f64(const CgpProblem *p, const double *ys, const double *step_weights, void *workspace,
                             size_t ws_bytes, int64_t ckpt_every, double *consts_bar, double *m0_bar, double *P0_bar,
                             double *Xi_bar, void *stream) {
    if (!step_weights) return CGP_ERR_BAD_ARG;
    return nll_bwd(p, ys, nullptr, step_weights, workspace, ws_bytes, ckpt_every, consts_bar, m0_bar, P0_bar, Xi_bar, stream);
}

}  // extern "C"
