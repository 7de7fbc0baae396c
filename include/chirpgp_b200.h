/*
 * chirpgp_b200.h -- C ABI of the B200-native chirpgp filtering/smoothing hot path.
 *
 * Drop-in boundary: each entry point replaces one function of the reference's
 * chirpgp/filters_smoothers.py (cited per function, paths relative to /root/reference/chirpgp/) for a whole
 * batch of independent chirps (what the reference obtains with jax.vmap, tetralith/jobs/crlb_ekf.py:68-72).
 * Plain pointers and sizes only: no torch / JAX types.  Bound by
 *   - chirpgp_b200/_native.py (ctypes; this image),
 *   - an XLA-FFI shim (chirpgp_b200/csrc/xla_ffi_shim.cc; needs jaxlib headers).
 *
 * Conventions
 *   - float64 everywhere, row-major, batch-major: ys [B,T], mfs [B,T,d], Pfs [B,T,d,d], nell [B,T].
 *   - All data pointers are DEVICE pointers unless the function name ends in _host.
 *   - The caller owns every buffer.  The library allocates nothing persistent; smoothers that need scratch
 *     take a caller-provided workspace whose size cgp_workspace_bytes() reports.
 *   - Asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, no mutable global state.
 *   - Return value: 0 on success, a negative CGP_ERR_* for argument errors, a positive cudaError_t for CUDA
 *     launch errors.  Never throws, never exits.  Numerical failure (non-PD covariance) is reported in-band
 *     as NaNs, matching JAX semantics.
 *   - T >= 1.  A smoother with T == 1 copies the filter result (filters_smoothers.py:140-142, :218).
 */
#ifndef CHIRPGP_B200_H
#define CHIRPGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGP_ABI_VERSION 3

/* model ids (chirpgp_b200/models.py uses the same numbers) */
enum {
    CGP_MODEL_LINEAR_DISC = 0, /* (u,dt)->(F u, Sigma): consts [F d*d | Sigma d*d]           (kf/rts, test lambdas) */
    CGP_MODEL_LCD         = 1, /* chirp/harmonic/La Scala LCD, models.py:295-309,:369-384,:423-432:
                                  consts [e, F00,F01,F10,F11, q, S00,S01,S11, freq_scale]  (CGP_NC_LCD)          */
    CGP_MODEL_LINEAR_SDE  = 2, /* u -> A u: consts [A d*d]                                                        */
    CGP_MODEL_SDE         = 3, /* chirp/harmonic drift, models.py:104-110,:164-168: consts
                                  [lam, gamma^2, 2 gamma, freq_scale]                       (CGP_NC_SDE)          */
    CGP_MODEL_KPT         = 4  /* KPT model, models.py:522-580: linear prediction consts [F d*d | Sigma d*d], d = num_harmonics + 2,
                                  measurement h(x) = sum_k x_k sin(k g(x_0 + x_{d-1}))      (cgp_ekf_for_kpt_f64 only)  */
};
#define CGP_NC_LCD 10
#define CGP_NC_SDE 4

/* sigma-point table layout hints (results do not depend on them; they select a kernel specialisation) */
enum {
    CGP_SIGMA_GENERIC = 0,
    CGP_SIGMA_GAUSS_HERMITE = 1, /* table is bit-identical to quadratures.py:157-196 with `gh_order` nodes/dim */
    CGP_SIGMA_CUBATURE = 2       /* table is bit-identical to quadratures.py:139-150 (2d points +-sqrt(d) e_j)       */
};

#define CGP_H_HARMONIC (-2)    /* CgpProblem::h_unit_index: H = [0 1 0 1 ... 0 0], models.py:257 */

enum {
    CGP_ERR_BAD_ARG = -1,      /* null pointer, B/T < 1, stride mismatch                       */
    CGP_ERR_UNSUPPORTED = -2,  /* (model, d, n_sigma) combination has no compiled kernel       */
    CGP_ERR_WORKSPACE = -3     /* workspace missing or too small                               */
};

/* One batch of independent filtering problems.  Problem p (0 <= p < B) reads
 *   ys     + (p / ys_repeat) * T          (ys_repeat >= 1: consecutive problems sharing one measurement row,
 *                                          used for hyper-parameter candidate grids)
 *   consts + p * consts_stride            (stride 0 = shared by all problems), likewise m0, P0.            */
typedef struct CgpProblem {
    int64_t B;                 /* number of problems (chirps x candidates)                     */
    int64_t T;                 /* samples per chirp                                            */
    int32_t model;             /* CGP_MODEL_*                                                  */
    int32_t d;                 /* state dimension                                              */
    int32_t num_harmonics;     /* LCD / SDE models: d == 2 * num_harmonics + 2                 */
    int32_t n_sigma;           /* sigma points (0 for kf/ekf/cd_ekf and their smoothers)       */
    int32_t sigma_kind;        /* CGP_SIGMA_*                                                  */
    int32_t gh_order;          /* nodes per dimension when sigma_kind == GAUSS_HERMITE         */
    int64_t ys_repeat;         /* >= 1                                                         */
    int32_t h_unit_index;      /* hint: j if H is exactly the unit vector e_j; CGP_H_HARMONIC if H = sum_k e_(2k+1),
                                  k < num_harmonics (the measurement row of the harmonic chirp models); else -1.
                                  Results are identical with and without the hint                            */
    int32_t in_flight;         /* hint: how many launches of this shape the caller keeps in flight on other streams (0 or 1:
                                  this one has the GPU to itself).  Only steers the choice between kernels that agree to
                                  rounding: with several batches in flight the GPU is throughput- rather than latency-bound   */
    const double *consts;  int64_t consts_stride;   /* model constants, see model ids          */
    const double *m0;      int64_t m0_stride;       /* [B|1, d]                                */
    const double *P0;      int64_t P0_stride;       /* [B|1, d, d] (symmetric)                 */
    const double *H;           /* [d] measurement row (1-D measurement, filters_smoothers.py:57-58) */
    const double *Qc;      int64_t Qc_stride;       /* [B|1, d, d] = b b^T, CD variants only   */
    const double *sig_w;       /* [n_sigma]                                                    */
    const double *sig_xi;      /* [n_sigma, d]                                                 */
    double Xi;                 /* measurement variance                                         */
    double dt;                 /* sampling interval (positive; smoothers negate internally)    */
} CgpProblem;

int cgp_abi_version(void);

/* ---- filters: ys [.,T] -> mfs [B,T,d], Pfs [B,T,d,d], nell [B,T] (cumulative -log lik., :180-184).
 * mfs/Pfs may both be NULL (nll-only: nothing but nell is stored);  nell may be NULL.
 * If nell_last_only != 0, nell is [B] and receives only the final cumulative value (the MLE objective,
 * demos/ekfs_mle.py:42-45).                                                                          */
int cgp_kf_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell,
               int nell_last_only, void *stream);                    /* filters_smoothers.py:145-184 */
int cgp_ekf_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell,
                int nell_last_only, void *stream);                   /* :222-264 */
int cgp_sgp_filter_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell,
                       int nell_last_only, void *stream);            /* :446-490 (+ :88-121) */
int cgp_ekf_for_kpt_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell,
                        int nell_last_only, void *stream);           /* :267-314, model CGP_MODEL_KPT (H, n_sigma unused);
                                                                        smooth its result with cgp_rts_f64 (same F, Sigma) */
int cgp_cd_ekf_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell,
                   int nell_last_only, void *stream);                /* :352-397 (+ quadratures.py:34-54) */
int cgp_cd_sgp_filter_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell,
                          int nell_last_only, void *stream);         /* :534-582 (+ :124-137) */

/* ---- smoothers: mfs [B,T,d], Pfs [B,T,d,d] -> mss, Pss (same shapes).  In-place (mss == mfs) is NOT allowed. */
size_t cgp_workspace_bytes(const char *function_name, const CgpProblem *p);
int cgp_rts_f64(const CgpProblem *p, const double *mfs, const double *Pfs, double *mss, double *Pss,
                void *workspace, size_t workspace_bytes, void *stream);          /* :187-219 */
int cgp_eks_f64(const CgpProblem *p, const double *mfs, const double *Pfs, double *mss, double *Pss,
                void *workspace, size_t workspace_bytes, void *stream);          /* :317-349 */
int cgp_sgp_smoother_f64(const CgpProblem *p, const double *mfs, const double *Pfs, double *mss, double *Pss,
                         void *workspace, size_t workspace_bytes, void *stream); /* :493-531 */
int cgp_cd_eks_f64(const CgpProblem *p, const double *mfs, const double *Pfs, double *mss, double *Pss,
                   void *workspace, size_t workspace_bytes, void *stream);       /* :400-443 */
int cgp_cd_sgp_smoother_f64(const CgpProblem *p, const double *mfs, const double *Pfs, double *mss, double *Pss,
                            void *workspace, size_t workspace_bytes, void *stream); /* :585-632 */

/* ---- fused filter + smoother gains.  sgp_smoother's reverse scan (:520-528) recomputes, from (mf_k, Pf_k), the sigma-point
 * prediction that sgp_filter's step k+1 already made (:88-121, called from :483 and :523); only the last two lines of
 * _gaussian_smoother_common (:83-84) depend on the smoothed state.  cgp_sgp_filter_gains_f64 is cgp_sgp_filter_f64 that
 * additionally fills the smoother workspace ([G | c | C packed] per (chirp, step) with c = mf - G mp, C = Pf - G D: d^2 + d + d (d + 1) / 2 doubles,
 * cgp_workspace_bytes("sgp_filter_gains")): inside the filter kernel where one exists (cgp_sgp_filter_gains_fused() == 1: chirp
 * LCD model with Gauss-Hermite order 3; harmonic chirp models d = 6, 8 with the cubature rule),
 * otherwise by running the time-parallel gain kernel after the filter.  cgp_smoother_sweep_f64 then finishes any of
 * rts / eks / sgp_smoother from a filled workspace (:83-84 for k = T-2 .. 0, stacking :140-142); it reads B, T, d of the
 * problem only.  Filter + sweep give the results of cgp_sgp_filter_f64 + cgp_sgp_smoother_f64 up to rounding. */
int cgp_sgp_filter_gains_fused(const CgpProblem *p);
int cgp_sgp_filter_gains_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell,
                             int nell_last_only, void *workspace, size_t workspace_bytes, void *stream);
int cgp_smoother_sweep_f64(const CgpProblem *p, const double *mfs, const double *Pfs, double *mss, double *Pss,
                           void *workspace, size_t workspace_bytes, void *stream);

/* ---- MLE path: EKF negative log-likelihood without per-step outputs, and its reverse-mode adjoint
 * (jax.grad of `ekf(...)[-1][-1]`, demos/ekfs_mle.py:42-49, tetralith/jobs/ekfs_mle.py:41-48).  LCD models only.
 * Persistent, ticket-scheduled kernels (csrc/cgp_nll2.cu): 32 problems form a chain, time is cut into segments of
 * `ckpt_every` steps, and warps pick (chain, segment) units from a global counter -- any number of problems keeps every SM
 * sub-partition equally busy.  The covariance is carried as its packed lower triangle (exactly symmetric arithmetic).
 *   fwd: nll [B] = final cumulative negative log-likelihood.  With a workspace, (m, P) checkpoints are stored every
 *        `ckpt_every` steps for the adjoint; workspace == NULL gives a pure objective evaluation (one unit per chain).
 *   bwd: given nll_bar [B] (NULL = ones) and the workspace filled by fwd (same problem, same ckpt_every), returns the
 *        cotangents of the kernel inputs: consts_bar [B, CGP_NC_LCD], m0_bar [B, d], P0_bar [B, d, d] (general matrix,
 *        JAX's unsymmetrised convention), Xi_bar [B] (may be NULL).  Shared inputs (stride 0) get per-problem
 *        cotangents that the caller sums.
 *   bwd_sym: the same with P0_bar symmetrised, (X + X^T)/2 of the above -- identical for every use in which P0 is a
 *        symmetric-matrix-valued function of the parameters (all callers of the reference), and ~10 % cheaper: the
 *        antisymmetric part of the covariance cotangent feeds nothing but itself and is not tracked.
 * The workspace is read and written by both calls (checkpoints, scheduling words, per-warp scratch); one workspace serves one
 * fwd / bwd pair at a time.  Its size depends on the current CUDA device (number of SMs). */
int64_t cgp_ekf_nll_default_ckpt(int64_t T);                       /* min(16, ~sqrt(T)) */
size_t cgp_ekf_nll_workspace_bytes(const CgpProblem *p, int64_t ckpt_every);
int cgp_ekf_nll_fwd_f64(const CgpProblem *p, const double *ys, double *nll, void *workspace, size_t workspace_bytes,
                        int64_t ckpt_every, void *stream);
int cgp_ekf_nll_bwd_f64(const CgpProblem *p, const double *ys, const double *nll_bar, void *workspace,
                        size_t workspace_bytes, int64_t ckpt_every, double *consts_bar, double *m0_bar, double *P0_bar,
                        double *Xi_bar, void *stream);
int cgp_ekf_nll_bwd_sym_f64(const CgpProblem *p, const double *ys, const double *nll_bar, void *workspace,
                            size_t workspace_bytes, int64_t ckpt_every, double *consts_bar, double *m0_bar, double *P0_bar,
                            double *Xi_bar, void *stream);

/* ---- forward-mode derivative of the nll of ANY filter on the path (csrc/cgp_tangent.cu): what jax.grad of
 * `sgp_filter(...)[-1][-1]`, `cd_ekf(...)[-1][-1]`, `cd_sgp_filter(...)[-1][-1]` (and `ekf`) computes in demos/ghfs_mle.py:54-61,
 * demos/cd_ekfs_mle.py, demos/cd_ghfs_mle.py.  `filter` is one of "ekf", "sgp_filter" (model CGP_MODEL_LCD), "cd_ekf",
 * "cd_sgp_filter" (model CGP_MODEL_SDE, p->Qc set); sigma tables in p for the sigma-point filters.  For each of the n_dir
 * parameter directions k the caller passes the tangents of the kernel inputs (any pointer may be NULL = zero tangent):
 *   consts_dot [B|1, n_dir, NC], m0_dot [B|1, n_dir, d], P0_dot [B|1, n_dir, d, d] (symmetrised on load),
 *   Qc_dot [B|1, n_dir, d, d], Xi_dot [n_dir];  the *_stride arguments are per-problem strides in doubles (0 = shared).
 * Outputs: nll [B] (may be NULL) and nll_dot [B, n_dir] = d nll / d direction_k.  One group of lanes per (problem, direction);
 * no workspace. */
int cgp_filter_nll_tangent_f64(const char *filter, const CgpProblem *p, const double *ys, int n_dir,
                               const double *consts_dot, int64_t consts_dot_stride, const double *m0_dot, int64_t m0_dot_stride,
                               const double *P0_dot, int64_t P0_dot_stride, const double *Qc_dot, int64_t Qc_dot_stride,
                               const double *Xi_dot, double *nll, double *nll_dot, void *stream);

/* ---- post-processing right after the smoothers: chirpgp.quadratures.gaussian_expectation (quadratures.py:234-274) for its
 * default integrand g (softplus, models.py:50) and d = 1 -- the frequency estimate E[g(V_k)], V_k ~ N(ms_k, chol_k^2), every
 * demo forms from the smoothing result (demos/ghfs_mle.py:87-89).  ms / sd are DEVICE pointers read with element strides
 * (so they may point into mss / Pss: ms = mss + 2, stride d; sd = Pss + 2 d + 2, stride d*d, sd_is_variance = 1 applies the
 * sqrt the callers apply); w_host / xi_host [order] are HOST pointers to the d = 1 Gauss-Hermite table of
 * quadratures.py:157-196 (order <= 64); out [n] is a device pointer. */
int cgp_gaussian_expectation_softplus_f64(int64_t n, const double *ms, int64_t ms_stride, const double *sd,
                                          int64_t sd_stride, int sd_is_variance, const double *w_host,
                                          const double *xi_host, int order, double *out, void *stream);

/* ---- Monte-Carlo input side (tetralith/jobs/crlb_ekf.py:41-56, test/test_crlb.py:41-55, tools.py:81-170): B trajectories of
 * a discretised model (CGP_MODEL_LINEAR_DISC or CGP_MODEL_LCD) and their measurements,
 *   x_0 = m0 + chol(P0) eps,  x_k = mean(x_{k-1}) + chol(Sigma) eps_k,  y_k = H x_k + sqrt(Xi) eps'_k,   k = 1 .. T,
 * with normals from the counter-based Philox4x32-10 generator (key = seed, counter = (draw, k, first_trajectory + b)) and
 * Box-Muller in float64: trajectory i depends on (seed, i) only, whatever the batch split.  Outputs (device pointers, each may
 * be NULL): x0 [B, d], xs [B, T, d], ys [B, T].  No bit parity with jax.random (threefry); oracle/sim_oracle.py restates the
 * generator in NumPy. */
int cgp_simulate_f64(const CgpProblem *p, uint64_t seed, uint64_t first_trajectory, double *x0, double *xs, double *ys,
                     void *stream);
/* test hooks: one Philox4x32-10 block (out_dev: 4 uint32 on the device); 2 n normals of a fixed counter pattern */
int cgp_test_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out_dev,
                    void *stream);
int cgp_test_normals(uint64_t seed, int64_t n, double *out_dev, void *stream);

/* The same pair for a cotangent on the WHOLE n_ell output (jax.vjp of ekf(...)[-1] with an arbitrary (T,) cotangent ct):
 *   path_fwd: nell [B, T] = cumulative negative log-likelihood at every step (:180-184), checkpoints as above;
 *   path_bwd: step_weights [B, T], the weight of every step's increment = sum_{j >= k} ct_j (a reversed cumulative sum the
 *             caller forms), otherwise as cgp_ekf_nll_bwd_f64. */
size_t cgp_ekf_nll_path_workspace_bytes(const CgpProblem *p, int64_t ckpt_every);
int cgp_ekf_nll_path_fwd_f64(const CgpProblem *p, const double *ys, double *nell, void *workspace, size_t workspace_bytes,
                             int64_t ckpt_every, void *stream);
int cgp_ekf_nll_path_bwd_f64(const CgpProblem *p, const double *ys, const double *step_weights, void *workspace,
                             size_t workspace_bytes, int64_t ckpt_every, double *consts_bar, double *m0_bar, double *P0_bar,
                             double *Xi_bar, void *stream);

/* ---- measurement utility: DFMA-only kernel (8 independent chains / thread) for the FP64 roofline denominator.
 * `out` holds blocks * 256 doubles.  Returns the flops issued (caller times the stream), < 0 on error. */
double cgp_bench_dfma(double *out, int blocks, int iters, void *stream);

/* ---- test hook: element-wise evaluation of the library's internal FP64 elementary functions (csrc/cgp_math.cuh)
 * kind: 0 exp, 1 softplus log(exp(x)+1), 2 sin, 3 cos, 4 rsqrt, 5 1/x, 6 sigmoid, 7 softplus (paired variant).
 * x, out: device pointers [n]. */
int cgp_test_math(int kind, int64_t n, const double *x, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CHIRPGP_B200_H */
