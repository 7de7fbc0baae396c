/*
 * chirpgp_oracle.c -- CPU restatement (plain C, float64) of chirpgp's filtering/smoothing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker / reported CPU baseline.
 * The product (chirpgp_b200/) never links, imports or falls back to it.
 *
 * Parity status: the real reference needs jax/jaxlib (unpinned, requirements.txt:3-4), which cannot be
 * installed here, and ships no golden vectors.  This restatement is pinned two ways:
 *   (1) against the UNMODIFIED reference sources executed over oracle/jaxshim (a torch-float64 stand-in
 *       for the JAX primitives) -- tests/golden/make_golden.py generates tests/golden/<case>.npz from them and
 *       tests/test_oracle_vs_golden.py checks every function here against those vectors;
 *   (2) against the reference's own JAX-free invariants (test/test_filters_smoothers.py:19-85 with its
 *       np.random.seed(666) data, test/test_models.py, test/test_m32.py, test/test_quadratures.py).
 * Element-wise parity with *real XLA* float64 rounding remains unpinned.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference/chirpgp/).
 * Layout: all matrices row-major; one chirp per call; or_batch() runs B chirps with OpenMP.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DMAX 16
#define OR_PI 3.141592653589793238462643383279502884

enum { OR_LINEAR = 0, OR_CHIRP = 1 };

typedef struct {
    int kind;                 /* OR_LINEAR or OR_CHIRP (chirp == harmonic with h = 1)              */
    int d;                    /* state dimension                                                    */
    int h;                    /* number of harmonics (OR_CHIRP): d = 2h + 2                         */
    int lascala;              /* 1: La Scala variant (no damping, no chirp noise; models.py:419-434) */
    double lam, b, ell, sigma, freq_scale;
    const double *F, *Sigma;  /* OR_LINEAR discrete model  x_k = F x_{k-1} + q,  q ~ N(0, Sigma)     */
    const double *A;          /* OR_LINEAR SDE drift matrix                                          */
    const double *Bm;         /* dispersion matrix (d x dw), used by CD variants                     */
    int dw;
} OrModel;

typedef struct {
    int d, n;
    const double *w;          /* (n)    */
    const double *xi;         /* (n, d) */
} OrSigma;

/* ---------------------------------------------------------------- small dense helpers */
static void matmul(int n, int k, int m, const double *A, const double *B, double *C) {   /* C(n,m) = A(n,k) B(k,m) */
    for (int i = 0; i < n; i++)
        for (int j = 0; j < m; j++) {
            double s = 0.;
            for (int l = 0; l < k; l++) s += A[i * k + l] * B[l * m + j];
            C[i * m + j] = s;
        }
}
static void matmul_nt(int n, int k, int m, const double *A, const double *B, double *C) { /* C(n,m) = A(n,k) B(m,k)^T */
    for (int i = 0; i < n; i++)
        for (int j = 0; j < m; j++) {
            double s = 0.;
            for (int l = 0; l < k; l++) s += A[i * k + l] * B[j * k + l];
            C[i * m + j] = s;
        }
}
static void matvec(int n, int m, const double *A, const double *x, double *y) {
    for (int i = 0; i < n; i++) {
        double s = 0.;
        for (int j = 0; j < m; j++) s += A[i * m + j] * x[j];
        y[i] = s;
    }
}
/* lower Cholesky reading only the lower triangle (LAPACK potrf('L') semantics, as
 * jax.scipy.linalg.cholesky(lower=True) / cho_factor do).  Non-PD input -> NaNs, never a trap. */
static void chol_lower(int d, const double *P, double *L) {
    memset(L, 0, sizeof(double) * d * d);
    for (int j = 0; j < d; j++) {
        double s = P[j * d + j];
        for (int k = 0; k < j; k++) s -= L[j * d + k] * L[j * d + k];
        double ljj = sqrt(s);
        L[j * d + j] = ljj;
        for (int i = j + 1; i < d; i++) {
            double t = P[i * d + j];
            for (int k = 0; k < j; k++) t -= L[i * d + k] * L[j * d + k];
            L[i * d + j] = t / ljj;
        }
    }
}
/* X = (L L^T)^{-1} Bm, Bm is (d, m) */
static void chol_solve(int d, int m, const double *L, const double *Bm, double *X) {
    double y[DMAX];
    for (int c = 0; c < m; c++) {
        for (int i = 0; i < d; i++) {
            double s = Bm[i * m + c];
            for (int k = 0; k < i; k++) s -= L[i * d + k] * y[k];
            y[i] = s / L[i * d + i];
        }
        for (int i = d - 1; i >= 0; i--) {
            double s = y[i];
            for (int k = i + 1; k < d; k++) s -= L[k * d + i] * X[k * m + c];
            X[i * m + c] = s / L[i * d + i];
        }
    }
}

/* ---------------------------------------------------------------- model functions */
static double softplus_naive(double x) { return log(exp(x) + 1.); }          /* models.py:50 (naive form)  */
static double sigmoid(double x) { double e = exp(x); return e / (e + 1.); }   /* d/dx of models.py:50       */

/* models.py:61-73 _m32_solution */
static void m32_solution(double ell, double sigma, double dt, double *Ft /*2x2*/, double *St /*2x2*/) {
    double gamma = sqrt(3.) / ell;
    double eta = dt * gamma;
    double beta = sigma * sigma * exp(-2 * eta);
    double ee = exp(-eta);
    Ft[0] = (1 + eta) * ee;             Ft[1] = dt * ee;
    Ft[2] = (-dt * (gamma * gamma)) * ee; Ft[3] = (1 - eta) * ee;
    St[0] = sigma * sigma - beta * (2 * eta + 2 * (eta * eta) + 1);
    St[1] = 2 * (dt * dt) * (gamma * gamma * gamma) * beta;
    St[2] = St[1];
    St[3] = (gamma * gamma) * (sigma * sigma + beta * (2 * eta - 2 * (eta * eta) - 1));
}

/* Conditional mean, its Jacobian (optional) and covariance (optional) of the discretised model.
 * OR_CHIRP: models.py:295-309 (disc_chirp_lcd), :369-384 (disc_harmonic_chirp_lcd), :423-432 (La Scala).
 * The Jacobian is the closed form of what jax.jacfwd(lambda u: cond_m_cov(u, dt)[0]) returns
 * (filters_smoothers.py:255, :342).  OR_LINEAR: mean F u, cov Sigma (test/test_filters_smoothers.py:68-70). */
static void disc_mean_cov(const OrModel *M, double dt, const double *u, double *mean, double *J, double *Sig) {
    int d = M->d;
    if (M->kind == OR_LINEAR) {
        matvec(d, d, M->F, u, mean);
        if (J) memcpy(J, M->F, sizeof(double) * d * d);
        if (Sig) memcpy(Sig, M->Sigma, sizeof(double) * d * d);
        return;
    }
    int h = M->h, v = d - 2;
    double lam = M->lascala ? 0. : M->lam;
    double gv = softplus_naive(u[v]);
    double w = 2 * OR_PI * gv * M->freq_scale;                    /* models.py:296 / :370 */
    double dw_du = 2 * OR_PI * sigmoid(u[v]) * M->freq_scale;
    double e = M->lascala ? 1. : exp(-lam * dt);
    double Fm[4], Sm[4];
    m32_solution(M->ell, M->sigma, dt, Fm, Sm);
    if (J) memset(J, 0, sizeof(double) * d * d);
    for (int k = 1; k <= h; k++) {
        double th = dt * k * w;                                   /* models.py:371 */
        double c = cos(th), s = sin(th);
        int r = 2 * (k - 1);
        double a00 = c * e, a01 = -s * e, a10 = s * e, a11 = c * e;
        mean[r] = a00 * u[r] + a01 * u[r + 1];
        mean[r + 1] = a10 * u[r] + a11 * u[r + 1];
        if (J) {
            J[r * d + r] = a00;       J[r * d + r + 1] = a01;
            J[(r + 1) * d + r] = a10; J[(r + 1) * d + r + 1] = a11;
            double dth = dt * k * dw_du;
            J[r * d + v] = e * (-s * u[r] - c * u[r + 1]) * dth;
            J[(r + 1) * d + v] = e * (c * u[r] - s * u[r + 1]) * dth;
        }
    }
    mean[v] = Fm[0] * u[v] + Fm[1] * u[v + 1];
    mean[v + 1] = Fm[2] * u[v] + Fm[3] * u[v + 1];
    if (J) {
        J[v * d + v] = Fm[0];       J[v * d + v + 1] = Fm[1];
        J[(v + 1) * d + v] = Fm[2]; J[(v + 1) * d + v + 1] = Fm[3];
    }
    if (Sig) {
        memset(Sig, 0, sizeof(double) * d * d);
        double q;
        if (M->lascala) q = 0.;
        else if (lam == 0.) q = M->b * M->b * dt;                                      /* models.py:303 */
        else q = M->b * M->b / (2 * lam) * (1 - exp(-2 * lam * dt));                   /* models.py:305 */
        for (int i = 0; i < 2 * h; i++) Sig[i * d + i] = q;
        Sig[v * d + v] = Sm[0];       Sig[v * d + v + 1] = Sm[1];
        Sig[(v + 1) * d + v] = Sm[2]; Sig[(v + 1) * d + v + 1] = Sm[3];
    }
}

/* SDE drift a(u) and (optional) Jacobian.  models.py:104-110 (chirp), :164-168 (harmonic), :246-252. */
static void sde_drift(const OrModel *M, const double *u, double *a, double *J) {
    int d = M->d;
    if (M->kind == OR_LINEAR) {
        matvec(d, d, M->A, u, a);
        if (J) memcpy(J, M->A, sizeof(double) * d * d);
        return;
    }
    int h = M->h, v = d - 2;
    double lam = M->lascala ? 0. : M->lam;
    double w = 2 * OR_PI * softplus_naive(u[v]) * M->freq_scale;
    double dw_du = 2 * OR_PI * sigmoid(u[v]) * M->freq_scale;
    double gamma = sqrt(3.) / M->ell;
    if (J) memset(J, 0, sizeof(double) * d * d);
    for (int k = 1; k <= h; k++) {
        int r = 2 * (k - 1);
        a[r] = -lam * u[r] + (-(w * k)) * u[r + 1];
        a[r + 1] = (w * k) * u[r] + (-lam) * u[r + 1];
        if (J) {
            J[r * d + r] = -lam;        J[r * d + r + 1] = -(w * k);
            J[(r + 1) * d + r] = w * k; J[(r + 1) * d + r + 1] = -lam;
            J[r * d + v] = -(dw_du * k) * u[r + 1];
            J[(r + 1) * d + v] = (dw_du * k) * u[r];
        }
    }
    a[v] = u[v + 1];
    a[v + 1] = -(gamma * gamma) * u[v] + (-2 * gamma) * u[v + 1];
    if (J) {
        J[v * d + v + 1] = 1.;
        J[(v + 1) * d + v] = -(gamma * gamma);
        J[(v + 1) * d + v + 1] = -2 * gamma;
    }
}
/* b b^T with b = M->Bm (d x dw).  For the chirp models the caller passes
 * diag(b, b, ..., 0, 2 sigma (sqrt3/ell)^1.5)  (models.py:112-113, :170-171). */
static void dispersion_gamma(const OrModel *M, double *Gam) { matmul_nt(M->d, M->dw, M->d, M->Bm, M->Bm, Gam); }

/* ---------------------------------------------------------------- shared steps */
/* filters_smoothers.py:55-68 _linear_update, with :44-45 (-norm.logpdf(y, pred, sqrt(S))) */
static double linear_update(int d, const double *mp, const double *Pp, const double *H, double Xi, double y,
                            double *mf, double *Pf) {
    double HP[DMAX], K[DMAX];
    for (int j = 0; j < d; j++) { double s = 0.; for (int i = 0; i < d; i++) s += H[i] * Pp[i * d + j]; HP[j] = s; }
    double S = 0.; for (int j = 0; j < d; j++) S += HP[j] * H[j];
    S += Xi;
    for (int i = 0; i < d; i++) { double s = 0.; for (int j = 0; j < d; j++) s += Pp[i * d + j] * H[j]; K[i] = s / S; }
    double pred = 0.; for (int i = 0; i < d; i++) pred += H[i] * mp[i];
    for (int i = 0; i < d; i++) mf[i] = mp[i] + K[i] * (y - pred);
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) Pf[i * d + j] = Pp[i * d + j] - (K[i] * K[j]) * S;
    double sc = sqrt(S), sc2 = sc * sc;
    return (log(2 * OR_PI * sc2) + (y - pred) * (y - pred) / sc2) / 2.;
}
/* filters_smoothers.py:71-85 _gaussian_smoother_common (DT is D^T); ms/Ps updated in place */
static void smoother_common(int d, const double *DT, const double *mf, const double *Pf, const double *mp,
                            const double *Pp, double *ms, double *Ps) {
    double L[DMAX * DMAX], X[DMAX * DMAX], G[DMAX * DMAX], dm[DMAX], dP[DMAX * DMAX], T1[DMAX * DMAX], T2[DMAX * DMAX];
    chol_lower(d, Pp, L);
    chol_solve(d, d, L, DT, X);
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) G[i * d + j] = X[j * d + i];
    for (int i = 0; i < d; i++) dm[i] = ms[i] - mp[i];
    for (int i = 0; i < d * d; i++) dP[i] = Ps[i] - Pp[i];
    double t[DMAX];
    matvec(d, d, G, dm, t);
    for (int i = 0; i < d; i++) ms[i] = mf[i] + t[i];
    matmul(d, d, d, G, dP, T1);
    matmul_nt(d, d, d, T1, G, T2);
    for (int i = 0; i < d * d; i++) Ps[i] = Pf[i] + T2[i];
}
/* filters_smoothers.py:88-121 _sgp_prediction; chi and evals optional outputs (n, d) */
static void sgp_prediction(const OrModel *M, const OrSigma *sg, double dt, const double *mf, const double *Pf,
                           double *mp, double *Pp, double *chi_out, double *ev_out) {
    int d = M->d, n = sg->n;
    double L[DMAX * DMAX], chi[DMAX], ev[DMAX], Sig[DMAX * DMAX];
    chol_lower(d, Pf, L);
    memset(mp, 0, sizeof(double) * d);
    memset(Pp, 0, sizeof(double) * d * d);
    for (int i = 0; i < n; i++) {
        for (int r = 0; r < d; r++) {                             /* quadratures.py:201 */
            double s = 0.;
            for (int c = 0; c < d; c++) s += L[r * d + c] * sg->xi[i * d + c];
            chi[r] = mf[r] + s;
        }
        disc_mean_cov(M, dt, chi, ev, NULL, Sig);
        for (int r = 0; r < d; r++) mp[r] += sg->w[i] * ev[r];
        for (int r = 0; r < d; r++) for (int c = 0; c < d; c++) Pp[r * d + c] += sg->w[i] * (ev[r] * ev[c] + Sig[r * d + c]);
        if (chi_out) memcpy(chi_out + (size_t)i * d, chi, sizeof(double) * d);
        if (ev_out) memcpy(ev_out + (size_t)i * d, ev, sizeof(double) * d);
    }
    for (int r = 0; r < d; r++) for (int c = 0; c < d; c++) Pp[r * d + c] -= mp[r] * mp[c];   /* :120 */
}
/* filters_smoothers.py:124-137 _cd_sgp_common */
static void cd_sgp_common(const OrModel *M, const OrSigma *sg, const double *Gam, const double *m, const double *P,
                          double *dm, double *dP) {
    int d = M->d, n = sg->n;
    double L[DMAX * DMAX], chi[DMAX], f[DMAX], Q[DMAX * DMAX];
    chol_lower(d, P, L);
    memset(dm, 0, sizeof(double) * d);
    memset(Q, 0, sizeof(double) * d * d);
    for (int i = 0; i < n; i++) {
        for (int r = 0; r < d; r++) {
            double s = 0.;
            for (int c = 0; c < d; c++) s += L[r * d + c] * sg->xi[i * d + c];
            chi[r] = m[r] + s;
        }
        sde_drift(M, chi, f, NULL);
        for (int r = 0; r < d; r++) dm[r] += sg->w[i] * f[r];
        for (int r = 0; r < d; r++) for (int c = 0; c < d; c++) Q[r * d + c] += sg->w[i] * ((chi[r] - m[r]) * f[c]);
    }
    for (int r = 0; r < d; r++) for (int c = 0; c < d; c++) dP[r * d + c] = Q[r * d + c] + Q[c * d + r] + Gam[r * d + c];
}

/* ---------------------------------------------------------------- RK4 (quadratures.py:34-54, :57-81) */
typedef void (*ode_fn)(void *ctx, const double *m, const double *P, double *dm, double *dP);
static void rk4_m_cov(int d, ode_fn f, void *ctx, double *m, double *P, double dt) {
    int dd = d * d;
    double k1m[DMAX], k2m[DMAX], k3m[DMAX], k4m[DMAX], tm[DMAX];
    double k1P[DMAX * DMAX], k2P[DMAX * DMAX], k3P[DMAX * DMAX], k4P[DMAX * DMAX], tP[DMAX * DMAX];
    f(ctx, m, P, k1m, k1P);
    for (int i = 0; i < d; i++) tm[i] = m[i] + dt * k1m[i] / 2;
    for (int i = 0; i < dd; i++) tP[i] = P[i] + dt * k1P[i] / 2;
    f(ctx, tm, tP, k2m, k2P);
    for (int i = 0; i < d; i++) tm[i] = m[i] + dt * k2m[i] / 2;
    for (int i = 0; i < dd; i++) tP[i] = P[i] + dt * k2P[i] / 2;
    f(ctx, tm, tP, k3m, k3P);
    for (int i = 0; i < d; i++) tm[i] = m[i] + dt * k3m[i];
    for (int i = 0; i < dd; i++) tP[i] = P[i] + dt * k3P[i];
    f(ctx, tm, tP, k4m, k4P);
    for (int i = 0; i < d; i++) m[i] = m[i] + dt * (k1m[i] + 2 * k2m[i] + 2 * k3m[i] + k4m[i]) / 6;
    for (int i = 0; i < dd; i++) P[i] = P[i] + dt * (k1P[i] + 2 * k2P[i] + 2 * k3P[i] + k4P[i]) / 6;
}

typedef struct {
    const OrModel *M; const OrSigma *sg; const double *Gam;
    const double *mf, *Pf;            /* smoother only */
    const double *LPf;                /* chol(Pf), smoother only */
} OdeCtx;

/* filters_smoothers.py:384-385 (cd_ekf odes) */
static void ode_cd_ekf(void *vc, const double *m, const double *P, double *dm, double *dP) {
    OdeCtx *c = (OdeCtx *)vc; int d = c->M->d;
    double J[DMAX * DMAX], PJt[DMAX * DMAX], JP[DMAX * DMAX];
    sde_drift(c->M, m, dm, J);
    matmul_nt(d, d, d, P, J, PJt);
    matmul(d, d, d, J, P, JP);
    for (int i = 0; i < d * d; i++) dP[i] = PJt[i] + JP[i] + c->Gam[i];
}
/* filters_smoothers.py:427-432 (cd_eks odes) */
static void ode_cd_eks(void *vc, const double *m, const double *P, double *dm, double *dP) {
    OdeCtx *c = (OdeCtx *)vc; int d = c->M->d;
    double J[DMAX * DMAX], a[DMAX], X[DMAX * DMAX], Mx[DMAX * DMAX], GamT[DMAX * DMAX], dmf[DMAX], z[DMAX], t[DMAX];
    double T1[DMAX * DMAX], T2[DMAX * DMAX];
    sde_drift(c->M, m, a, J);
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) GamT[i * d + j] = c->Gam[j * d + i];
    chol_solve(d, d, c->LPf, GamT, X);                                  /* Pf^{-1} gamma^T            */
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) Mx[i * d + j] = J[i * d + j] + X[j * d + i];
    for (int i = 0; i < d; i++) dmf[i] = m[i] - c->mf[i];
    chol_solve(d, 1, c->LPf, dmf, z);
    matvec(d, d, c->Gam, z, t);
    for (int i = 0; i < d; i++) dm[i] = a[i] + t[i];
    matmul(d, d, d, Mx, P, T1);
    matmul_nt(d, d, d, P, Mx, T2);
    for (int i = 0; i < d * d; i++) dP[i] = T1[i] + T2[i] - c->Gam[i];
}
/* filters_smoothers.py:569-570 (cd_sgp_filter odes) */
static void ode_cd_sgp(void *vc, const double *m, const double *P, double *dm, double *dP) {
    OdeCtx *c = (OdeCtx *)vc;
    cd_sgp_common(c->M, c->sg, c->Gam, m, P, dm, dP);
}
/* filters_smoothers.py:615-621 (cd_sgp_smoother odes) */
static void ode_cd_sgp_smoother(void *vc, const double *m, const double *P, double *dm, double *dP) {
    OdeCtx *c = (OdeCtx *)vc; int d = c->M->d;
    double G[DMAX * DMAX], _m[DMAX], _P[DMAX * DMAX], dmf[DMAX], T1[DMAX * DMAX], T2[DMAX * DMAX];
    chol_solve(d, d, c->LPf, c->Gam, G);                                /* G = Pf^{-1} gamma          */
    cd_sgp_common(c->M, c->sg, c->Gam, m, P, _m, _P);
    for (int i = 0; i < d; i++) dmf[i] = m[i] - c->mf[i];
    for (int i = 0; i < d; i++) { double s = 0.; for (int k = 0; k < d; k++) s += G[k * d + i] * dmf[k]; dm[i] = _m[i] + s; }
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) {
        double s = 0.; for (int k = 0; k < d; k++) s += G[k * d + i] * P[k * d + j]; T1[i * d + j] = s; }
    matmul(d, d, d, P, G, T2);
    for (int i = 0; i < d * d; i++) dP[i] = _P[i] + T1[i] + T2[i] - 2 * c->Gam[i];
}

/* ---------------------------------------------------------------- filters (single chirp) */
enum { V_KF = 0, V_RTS, V_EKF, V_EKS, V_SGP_FILTER, V_SGP_SMOOTHER, V_CD_EKF, V_CD_EKS, V_CD_SGP_FILTER,
       V_CD_SGP_SMOOTHER };

/* kf :145-184, ekf :222-264, sgp_filter :446-490, cd_ekf :352-397, cd_sgp_filter :534-582.
 * n_ell is the CUMULATIVE negative log-likelihood at every step (:180-184). */
static void run_filter(int variant, const OrModel *M, const OrSigma *sg, const double *H, double Xi, const double *m0,
                       const double *P0, double dt, int64_t T, const double *ys, double *mfs, double *Pfs, double *nell) {
    int d = M->d, dd = d * d;
    double mf[DMAX], Pf[DMAX * DMAX], mp[DMAX], Pp[DMAX * DMAX], J[DMAX * DMAX], Sig[DMAX * DMAX], T1[DMAX * DMAX];
    double Gam[DMAX * DMAX];
    memcpy(mf, m0, sizeof(double) * d);
    memcpy(Pf, P0, sizeof(double) * dd);
    OdeCtx ctx = { M, sg, Gam, NULL, NULL, NULL };
    if (variant == V_CD_EKF || variant == V_CD_SGP_FILTER) dispersion_gamma(M, Gam);
    double acc = 0.;
    for (int64_t t = 0; t < T; t++) {
        switch (variant) {
        case V_KF:                                                           /* :48-52 */
        case V_EKF:                                                          /* :255-257 */
            disc_mean_cov(M, dt, mf, mp, J, Sig);
            matmul(d, d, d, J, Pf, T1);
            matmul_nt(d, d, d, T1, J, Pp);
            for (int i = 0; i < dd; i++) Pp[i] += Sig[i];
            break;
        case V_SGP_FILTER:
            sgp_prediction(M, sg, dt, mf, Pf, mp, Pp, NULL, NULL);
            break;
        case V_CD_EKF:
            memcpy(mp, mf, sizeof(double) * d); memcpy(Pp, Pf, sizeof(double) * dd);
            rk4_m_cov(d, ode_cd_ekf, &ctx, mp, Pp, dt);
            break;
        case V_CD_SGP_FILTER:
            memcpy(mp, mf, sizeof(double) * d); memcpy(Pp, Pf, sizeof(double) * dd);
            rk4_m_cov(d, ode_cd_sgp, &ctx, mp, Pp, dt);
            break;
        }
        double inc = linear_update(d, mp, Pp, H, Xi, ys[t], mf, Pf);
        acc = acc + inc;
        if (mfs) memcpy(mfs + t * d, mf, sizeof(double) * d);
        if (Pfs) memcpy(Pfs + t * dd, Pf, sizeof(double) * dd);
        if (nell) nell[t] = acc;
    }
}

/* rts :187-219, eks :317-349, sgp_smoother :493-531, cd_eks :400-443, cd_sgp_smoother :585-632.
 * Reverse scan over (mfs[:-1], Pfs[:-1]) from (mfs[-1], Pfs[-1]); last row = filter's last (:140-142). */
static void run_smoother(int variant, const OrModel *M, const OrSigma *sg, double dt, int64_t T, const double *mfs,
                         const double *Pfs, double *mss, double *Pss) {
    int d = M->d, dd = d * d;
    double ms[DMAX], Ps[DMAX * DMAX], mp[DMAX], Pp[DMAX * DMAX], J[DMAX * DMAX], Sig[DMAX * DMAX], T1[DMAX * DMAX];
    double DT[DMAX * DMAX], Gam[DMAX * DMAX], LPf[DMAX * DMAX];
    double *chi = NULL, *ev = NULL;
    if (variant == V_SGP_SMOOTHER) { chi = malloc(sizeof(double) * sg->n * d); ev = malloc(sizeof(double) * sg->n * d); }
    memcpy(ms, mfs + (T - 1) * d, sizeof(double) * d);
    memcpy(Ps, Pfs + (T - 1) * dd, sizeof(double) * dd);
    memcpy(mss + (T - 1) * d, ms, sizeof(double) * d);
    memcpy(Pss + (T - 1) * dd, Ps, sizeof(double) * dd);
    OdeCtx ctx = { M, sg, Gam, NULL, NULL, LPf };
    if (variant == V_CD_EKS || variant == V_CD_SGP_SMOOTHER) dispersion_gamma(M, Gam);
    for (int64_t t = T - 2; t >= 0; t--) {
        const double *mf = mfs + t * d, *Pf = Pfs + t * dd;
        switch (variant) {
        case V_RTS:
        case V_EKS:
            disc_mean_cov(M, dt, mf, mp, J, Sig);
            matmul(d, d, d, J, Pf, T1);                 /* DT = J Pf  (:345, :212) */
            matmul_nt(d, d, d, T1, J, Pp);
            for (int i = 0; i < dd; i++) Pp[i] += Sig[i];
            smoother_common(d, T1, mf, Pf, mp, Pp, ms, Ps);
            break;
        case V_SGP_SMOOTHER: {
            sgp_prediction(M, sg, dt, mf, Pf, mp, Pp, chi, ev);
            /* D = sum_i w_i chi_i ev_i^T - mf mp^T (:525); DT = D^T */
            for (int r = 0; r < d; r++) for (int c = 0; c < d; c++) {
                double s = 0.;
                for (int i = 0; i < sg->n; i++) s += sg->w[i] * (chi[i * d + r] * ev[i * d + c]);
                DT[c * d + r] = s - mf[r] * mp[c];
            }
            smoother_common(d, DT, mf, Pf, mp, Pp, ms, Ps);
            break; }
        case V_CD_EKS:
            ctx.mf = mf; ctx.Pf = Pf; chol_lower(d, Pf, LPf);
            rk4_m_cov(d, ode_cd_eks, &ctx, ms, Ps, -dt);
            break;
        case V_CD_SGP_SMOOTHER:
            ctx.mf = mf; ctx.Pf = Pf; chol_lower(d, Pf, LPf);
            rk4_m_cov(d, ode_cd_sgp_smoother, &ctx, ms, Ps, -dt);
            break;
        }
        memcpy(mss + t * d, ms, sizeof(double) * d);
        memcpy(Pss + t * dd, Ps, sizeof(double) * dd);
    }
    free(chi); free(ev);
}

/* ---------------------------------------------------------------- exported API */
/* One call = B independent chirps (the vmap analogue, tetralith/jobs/crlb_ekf.py:68-72), OpenMP over chirps.
 * Strides (in doubles) of 0 mean "shared by all chirps".  Per-chirp hyper-parameters: hp (B|1, 5) =
 * (lam, b, ell, sigma, freq_scale).  Filters: in ys (B,T) -> out mfs (B,T,d), Pfs (B,T,d,d), nell (B,T).
 * Smoothers: in mfs/Pfs -> out mss/Pss.  Any output pointer may be NULL (filters) to skip storing. */
int or_batch(int variant, int64_t B, int64_t T, const OrModel *proto, const double *hp, int64_t hp_stride,
             const double *sig_w, const double *sig_xi, int n_sigma,
             const double *H, double Xi, const double *m0, int64_t m0_stride, const double *P0, int64_t P0_stride,
             double dt, const double *ys, int64_t ys_stride,
             const double *mfs_in, const double *Pfs_in,
             double *out_m, double *out_P, double *out_nell, int nthreads) {
    int d = proto->d;
    if (d > DMAX) return -1;
    OrSigma sg = { d, n_sigma, sig_w, sig_xi };
    int is_filter = (variant == V_KF || variant == V_EKF || variant == V_SGP_FILTER || variant == V_CD_EKF ||
                     variant == V_CD_SGP_FILTER);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t bi = 0; bi < B; bi++) {
        OrModel M = *proto;
        if (hp && M.kind == OR_CHIRP) {
            const double *p = hp + bi * hp_stride;
            M.lam = p[0]; M.b = p[1]; M.ell = p[2]; M.sigma = p[3]; M.freq_scale = p[4];
        }
        size_t o1 = (size_t)bi * T * d, o2 = (size_t)bi * T * d * d;
        if (is_filter)
            run_filter(variant, &M, &sg, H, Xi, m0 + bi * m0_stride, P0 + bi * P0_stride, dt, T, ys + bi * ys_stride,
                       out_m ? out_m + o1 : NULL, out_P ? out_P + o2 : NULL, out_nell ? out_nell + (size_t)bi * T : NULL);
        else
            run_smoother(variant, &M, &sg, dt, T, mfs_in + o1, Pfs_in + o2, out_m + o1, out_P + o2);
    }
    return 0;
}

/* ---------------------------------------------------------------- KPT model
 * ekf_for_kpt, filters_smoothers.py:267-314, with the measurement function of build_kpt_chirp_model, models.py:572-578:
 *   state x = [omega, a_1 .. a_nh, phase] (d = nh + 2),  h(x) = sum_k a_k sin(k g(x_0 + x_{d-1})),  g = softplus (:50).
 * Per step: (mp, Pp) = _linear_predict(F, Sigma) (:48-52, :299);  H = jacfwd(h)(mp) (:301) in closed form:
 *   dh/dx_0 = dh/dx_{d-1} = g'(x_0 + x_{d-1}) sum_k k a_k cos(k phi),  dh/da_k = sin(k phi);
 * S = H Pp H^T + Xi, K = Pp H^T / S, pred = h(mp), mf = mp + K (y - pred), Pf = Pp - K K^T S, n_ell -= logpdf (:302-308). */
static void run_ekf_kpt(int d, int nh, const double *F, const double *Sigma, double Xi, const double *m0, const double *P0,
                        int64_t T, const double *ys, double *mfs, double *Pfs, double *nell) {
    int dd = d * d;
    double mf[DMAX], Pf[DMAX * DMAX], mp[DMAX], Pp[DMAX * DMAX], T1[DMAX * DMAX], H[DMAX], HP[DMAX], K[DMAX];
    memcpy(mf, m0, sizeof(double) * d);
    memcpy(Pf, P0, sizeof(double) * dd);
    double acc = 0.;
    for (int64_t t = 0; t < T; t++) {
        matvec(d, d, F, mf, mp);
        matmul(d, d, d, F, Pf, T1);
        matmul_nt(d, d, d, T1, F, Pp);
        for (int i = 0; i < dd; i++) Pp[i] += Sigma[i];
        double arg = mp[0] + mp[d - 1], phi = softplus_naive(arg), dphi = sigmoid(arg);
        double pred = 0., dsum = 0.;
        for (int k = 1; k <= nh; k++) {
            double sn = sin(phi * k), cs = cos(phi * k);
            pred += mp[k] * sn;
            dsum += mp[k] * (cs * k);
            H[k] = sn;
        }
        H[0] = dsum * dphi; H[d - 1] = dsum * dphi;
        for (int j = 0; j < d; j++) { double sacc = 0.; for (int i = 0; i < d; i++) sacc += H[i] * Pp[i * d + j]; HP[j] = sacc; }
        double S = 0.; for (int j = 0; j < d; j++) S += HP[j] * H[j];
        S += Xi;
        for (int i = 0; i < d; i++) { double sacc = 0.; for (int j = 0; j < d; j++) sacc += Pp[i * d + j] * H[j]; K[i] = sacc / S; }
        for (int i = 0; i < d; i++) mf[i] = mp[i] + K[i] * (ys[t] - pred);
        for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) Pf[i * d + j] = Pp[i * d + j] - (K[i] * K[j]) * S;
        double sc = sqrt(S), sc2 = sc * sc;
        acc = acc + (log(2 * OR_PI * sc2) + (ys[t] - pred) * (ys[t] - pred) / sc2) / 2.;
        if (mfs) memcpy(mfs + t * d, mf, sizeof(double) * d);
        if (Pfs) memcpy(Pfs + t * dd, Pf, sizeof(double) * dd);
        if (nell) nell[t] = acc;
    }
}
/* B chirps; F, Sigma (d,d), m0 (d), P0 (d,d) shared or per chirp (stride 0 = shared) */
int or_ekf_kpt_batch(int64_t B, int64_t T, int d, int nh, const double *F, const double *Sigma, int64_t FS_stride, double Xi,
                     const double *m0, int64_t m0_stride, const double *P0, int64_t P0_stride, const double *ys,
                     int64_t ys_stride, double *out_m, double *out_P, double *out_nell, int nthreads) {
    if (d > DMAX || d != nh + 2) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t bi = 0; bi < B; bi++)
        run_ekf_kpt(d, nh, F + bi * FS_stride, Sigma + bi * FS_stride, Xi, m0 + bi * m0_stride, P0 + bi * P0_stride, T,
                    ys + bi * ys_stride, out_m ? out_m + (size_t)bi * T * d : NULL,
                    out_P ? out_P + (size_t)bi * T * d * d : NULL, out_nell ? out_nell + (size_t)bi * T : NULL);
    return 0;
}

/* model probes used by the tests (test/test_models.py, test/test_m32.py analogues) */
void or_m32_solution(double ell, double sigma, double dt, double *Ft, double *St) { m32_solution(ell, sigma, dt, Ft, St); }
void or_disc_mean_cov(const OrModel *M, double dt, const double *u, double *mean, double *J, double *Sig) {
    disc_mean_cov(M, dt, u, mean, J, Sig);
}
void or_sde_drift(const OrModel *M, const double *u, double *a, double *J) { sde_drift(M, u, a, J); }
int or_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
