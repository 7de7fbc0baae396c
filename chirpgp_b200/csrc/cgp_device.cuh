// cgp_device.cuh -- device-side building blocks shared by all chirpgp_b200 kernels (sm_100a, FP64 SIMT).
//
// Everything here is register-resident small-matrix code: D (state dim) is a compile-time constant and all
// loops are fully unrolled, so `double a[D][D]` arrays never touch local memory.  Tensor cores are not used:
// nothing on this path is a dense contraction (d <= 12, sequential time loop).
//
// Reference semantics followed (paths relative to /root/reference/chirpgp/):
//   linear_update      filters_smoothers.py:55-68 (+ :44-45, jax.scipy.stats.norm.logpdf operation order)
//   chol_lower         jax.scipy.linalg.cholesky(lower=True) / cho_factor: potrf('L'), reads the lower triangle
//   chol_solve         jax.scipy.linalg.cho_solve
//   ModelLCD           models.py:50 (naive softplus), :61-73, :295-309, :369-384, :423-432
//   ModelSDE           models.py:104-113, :164-171, :246-255
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/chirpgp_b200.h"
#include "cgp_math.cuh"

namespace cgp {

constexpr double kTwoPi = 6.283185307179586;   // 2 * math.pi in float64

#define CGP_DEV __device__ __forceinline__
#define CGP_UNROLL _Pragma("unroll")

// ------------------------------------------------------------------------------------------------ packing
// symmetric matrices are stored packed-lower: (r, c) with r >= c at r (r + 1) / 2 + c
CGP_DEV constexpr int sidx(int r, int c) { return r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r; }
template <int D> struct NSym { static constexpr int value = D * (D + 1) / 2; };

// ------------------------------------------------------------------------------------------------ loads / stores
template <int N> CGP_DEV void load_vec(const double *__restrict__ src, double (&dst)[N]) {
    if constexpr (N % 2 == 0) {
        CGP_UNROLL for (int i = 0; i < N; i += 2) {
            double2 v = *reinterpret_cast<const double2 *>(src + i);
            dst[i] = v.x; dst[i + 1] = v.y;
        }
    } else {
        CGP_UNROLL for (int i = 0; i < N; i++) dst[i] = src[i];
    }
}
template <int N> CGP_DEV void store_vec(double *__restrict__ dst, const double (&src)[N]) {
    if constexpr (N % 2 == 0) {
        CGP_UNROLL for (int i = 0; i < N; i += 2)
            *reinterpret_cast<double2 *>(dst + i) = make_double2(src[i], src[i + 1]);
    } else {
        CGP_UNROLL for (int i = 0; i < N; i++) dst[i] = src[i];
    }
}
template <int D> CGP_DEV void load_mat(const double *__restrict__ src, double (&dst)[D][D]) {
    if constexpr (D % 2 == 0) {
        CGP_UNROLL for (int r = 0; r < D; r++)
            CGP_UNROLL for (int c = 0; c < D; c += 2) {
                double2 v = *reinterpret_cast<const double2 *>(src + r * D + c);
                dst[r][c] = v.x; dst[r][c + 1] = v.y;
            }
    } else {
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) dst[r][c] = src[r * D + c];
    }
}
template <int D> CGP_DEV void store_mat(double *__restrict__ dst, const double (&src)[D][D]) {
    if constexpr (D % 2 == 0) {
        CGP_UNROLL for (int r = 0; r < D; r++)
            CGP_UNROLL for (int c = 0; c < D; c += 2)
                *reinterpret_cast<double2 *>(dst + r * D + c) = make_double2(src[r][c], src[r][c + 1]);
    } else {
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) dst[r * D + c] = src[r][c];
    }
}
// 256-bit global accesses (sm_100: LDG/STG.E.256).  One lane moves a whole 32-byte sector: a thread-per-chirp kernel whose
// lanes write to 32 different chirps would otherwise fill every sector with two 16-byte partial writes.
// Pointers must be 32-byte aligned.
CGP_DEV void ldg256(const double *__restrict__ q, double &a, double &b, double &c, double &d) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(q));
}
CGP_DEV void stg256(double *__restrict__ q, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(q), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
template <bool WIDE, int N> CGP_DEV void gload_vec(const double *__restrict__ src, double (&dst)[N]) {
    if constexpr (WIDE && N % 4 == 0) {
        CGP_UNROLL for (int i = 0; i < N; i += 4) ldg256(src + i, dst[i], dst[i + 1], dst[i + 2], dst[i + 3]);
    } else {
        load_vec<N>(src, dst);
    }
}
template <bool WIDE, int N> CGP_DEV void gstore_vec(double *__restrict__ dst, const double (&src)[N]) {
    if constexpr (WIDE && N % 4 == 0) {
        CGP_UNROLL for (int i = 0; i < N; i += 4) stg256(dst + i, src[i], src[i + 1], src[i + 2], src[i + 3]);
    } else {
        store_vec<N>(dst, src);
    }
}
template <bool WIDE, int D> CGP_DEV void gload_mat(const double *__restrict__ src, double (&dst)[D][D]) {
    if constexpr (WIDE && D % 4 == 0) {
        CGP_UNROLL for (int r = 0; r < D; r++)
            CGP_UNROLL for (int c = 0; c < D; c += 4) ldg256(src + r * D + c, dst[r][c], dst[r][c + 1], dst[r][c + 2], dst[r][c + 3]);
    } else {
        load_mat<D>(src, dst);
    }
}
template <bool WIDE, int D> CGP_DEV void gstore_mat(double *__restrict__ dst, const double (&src)[D][D]) {
    if constexpr (WIDE && D % 4 == 0) {
        CGP_UNROLL for (int r = 0; r < D; r++)
            CGP_UNROLL for (int c = 0; c < D; c += 4) stg256(dst + r * D + c, src[r][c], src[r][c + 1], src[r][c + 2], src[r][c + 3]);
    } else {
        store_mat<D>(dst, src);
    }
}

// Stores of one lane's record to GLOBAL memory that pick the 256-bit form whenever the destination allows it (checked at run
// time: 32-byte aligned, element count a multiple of 4) -- for kernels where the lanes of a warp write different records.
template <int N> CGP_DEV void gstore_vec_auto(double *__restrict__ dst, const double (&src)[N]) {
    if (N % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) gstore_vec<true, N>(dst, src);
    else store_vec<N>(dst, src);
}
template <int D> CGP_DEV void gstore_mat_auto(double *__restrict__ dst, const double (&src)[D][D]) {
    if (D % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) gstore_mat<true, D>(dst, src);
    else store_mat<D>(dst, src);
}
template <int D> CGP_DEV void gstore_sym_auto(double *__restrict__ dst, const double (&src)[D * (D + 1) / 2]);

// Cumulative nll of ONE chirp (thread-per-chirp kernels) written in whole 32-byte sectors: the values are collected four at a
// time at the positions the row occupies in memory ((b T + t) mod 4); aligned groups go out as one 256-bit store, the
// ragged ends of the row (and everything, if the buffer is not 32-byte aligned) as scalars.
struct NellRowWriter {
    double v0, v1, v2, v3;
    double *row;
    int64_t g0;
    bool wide;
    CGP_DEV void init(double *nell, int64_t b, int64_t T) {
        row = nell + b * T; g0 = b * T;
        wide = (reinterpret_cast<uintptr_t>(nell) & 31u) == 0;
        v0 = v1 = v2 = v3 = 0.;
    }
    CGP_DEV void put(int64_t t, int64_t T, double acc) {
        if (!wide) { row[t] = acc; return; }
        const int q = (int)((g0 + t) & 3);
        if (q == 0) v0 = acc; else if (q == 1) v1 = acc; else if (q == 2) v2 = acc; else v3 = acc;
        if (q == 3 && t >= 3) { stg256(row + t - 3, v0, v1, v2, v3); return; }
        if (q == 3 || t == T - 1) {                       // ragged start / end of the row: steps t - q .. t that exist
            if (t - q >= 0) row[t - q] = v0;
            if (q >= 1 && t - q + 1 >= 0) row[t - q + 1] = v1;
            if (q >= 2 && t - q + 2 >= 0) row[t - q + 2] = v2;
            if (q >= 3) row[t] = v3;
        }
    }
};

// packed-symmetric <-> full row-major memory
template <int D> CGP_DEV void load_sym(const double *__restrict__ src, double (&dst)[NSym<D>::value]) {
    // reads the lower triangle only (what cholesky would see)
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c <= r; c++) dst[sidx(r, c)] = src[r * D + c];
}
template <int D> CGP_DEV void store_sym(double *__restrict__ dst, const double (&src)[NSym<D>::value]) {
    if constexpr (D % 2 == 0) {
        CGP_UNROLL for (int r = 0; r < D; r++)
            CGP_UNROLL for (int c = 0; c < D; c += 2)
                *reinterpret_cast<double2 *>(dst + r * D + c) = make_double2(src[sidx(r, c)], src[sidx(r, c + 1)]);
    } else {
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) dst[r * D + c] = src[sidx(r, c)];
    }
}

template <int D> CGP_DEV void gstore_sym_auto(double *__restrict__ dst, const double (&src)[D * (D + 1) / 2]) {
    if (D % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
        CGP_UNROLL for (int r = 0; r < D; r++)
            CGP_UNROLL for (int c = 0; c < D; c += 4)
                stg256(dst + r * D + c, src[sidx(r, c)], src[sidx(r, c + 1)], src[sidx(r, c + 2)], src[sidx(r, c + 3)]);
    } else {
        store_sym<D>(dst, src);
    }
}

// ------------------------------------------------------------------------------------------------ dense helpers
template <int D> CGP_DEV void matmul(const double (&A)[D][D], const double (&B)[D][D], double (&C)[D][D]) {
    CGP_UNROLL for (int i = 0; i < D; i++)
        CGP_UNROLL for (int j = 0; j < D; j++) {
            double s = A[i][0] * B[0][j];
            CGP_UNROLL for (int k = 1; k < D; k++) s = fma(A[i][k], B[k][j], s);
            C[i][j] = s;
        }
}
template <int D> CGP_DEV void matmul_nt(const double (&A)[D][D], const double (&B)[D][D], double (&C)[D][D]) {
    CGP_UNROLL for (int i = 0; i < D; i++)
        CGP_UNROLL for (int j = 0; j < D; j++) {
            double s = A[i][0] * B[j][0];
            CGP_UNROLL for (int k = 1; k < D; k++) s = fma(A[i][k], B[j][k], s);
            C[i][j] = s;
        }
}
template <int D> CGP_DEV void matvec(const double (&A)[D][D], const double (&x)[D], double (&y)[D]) {
    CGP_UNROLL for (int i = 0; i < D; i++) {
        double s = A[i][0] * x[0];
        CGP_UNROLL for (int k = 1; k < D; k++) s = fma(A[i][k], x[k], s);
        y[i] = s;
    }
}

// Products with the Jacobian J of a model mean that skip its structural zeros (Model::jnz) at compile time.  The terms
// are taken in the same order as the dense loops, so for finite operands the results are bit-identical (a skipped term
// is an exact zero); d = 4 chirp model: 10 of 16 entries are non-zero.
#define CGP_JSUM(nzexpr, aexpr, bexpr)                                                            \
    double sacc = 0.; bool first = true;                                                          \
    CGP_UNROLL for (int k = 0; k < D; k++)                                                        \
        if (nzexpr) { sacc = first ? (aexpr) * (bexpr) : fma((aexpr), (bexpr), sacc); first = false; }
template <class Model, int D> CGP_DEV void jmul(const double (&J)[D][D], const double (&Bm)[D][D], double (&Cm)[D][D]) {          // J B
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j < D; j++) { CGP_JSUM(Model::jnz(i, k), J[i][k], Bm[k][j]) Cm[i][j] = sacc; }
}
template <class Model, int D> CGP_DEV void jmul_nt(const double (&J)[D][D], const double (&Bm)[D][D], double (&Cm)[D][D]) {       // J B^T
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j < D; j++) { CGP_JSUM(Model::jnz(i, k), J[i][k], Bm[j][k]) Cm[i][j] = sacc; }
}
template <class Model, int D> CGP_DEV void mul_jt(const double (&A)[D][D], const double (&J)[D][D], double (&Cm)[D][D]) {         // A J^T
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j < D; j++) { CGP_JSUM(Model::jnz(j, k), A[i][k], J[j][k]) Cm[i][j] = sacc; }
}
template <class Model, int D> CGP_DEV void mul_j(const double (&A)[D][D], const double (&J)[D][D], double (&Cm)[D][D]) {          // A J
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j < D; j++) { CGP_JSUM(Model::jnz(k, j), A[i][k], J[k][j]) Cm[i][j] = sacc; }
}
template <class Model, int D> CGP_DEV void jtmul(const double (&J)[D][D], const double (&Bm)[D][D], double (&Cm)[D][D]) {         // J^T B
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j < D; j++) { CGP_JSUM(Model::jnz(k, i), J[k][i], Bm[k][j]) Cm[i][j] = sacc; }
}
template <class Model, int D> CGP_DEV void jtvec(const double (&J)[D][D], const double (&x)[D], double (&y)[D]) {                 // J^T x
    CGP_UNROLL for (int i = 0; i < D; i++) { CGP_JSUM(Model::jnz(k, i), J[k][i], x[k]) y[i] = sacc; }
}
#undef CGP_JSUM

// lower Cholesky of a full matrix, reading its lower triangle only.  L full (upper part left untouched = 0
// must be provided by the caller if needed; only the lower triangle is ever read afterwards).
template <int D> CGP_DEV void chol_lower(const double (&P)[D][D], double (&L)[D][D]) {
    CGP_UNROLL for (int j = 0; j < D; j++) {
        double s = P[j][j];
        CGP_UNROLL for (int k = 0; k < j; k++) s = fma(-L[j][k], L[j][k], s);
        double ljj = sqrt(s);
        L[j][j] = ljj;
        CGP_UNROLL for (int i = j + 1; i < D; i++) {
            double t = P[i][j];
            CGP_UNROLL for (int k = 0; k < j; k++) t = fma(-L[i][k], L[j][k], t);
            L[i][j] = t / ljj;
        }
    }
}
template <int D> CGP_DEV void chol_lower_sym(const double (&P)[NSym<D>::value], double (&L)[NSym<D>::value]) {
    CGP_UNROLL for (int j = 0; j < D; j++) {
        double s = P[sidx(j, j)];
        CGP_UNROLL for (int k = 0; k < j; k++) s = fma(-L[sidx(j, k)], L[sidx(j, k)], s);
        double ljj = sqrt(s);
        L[sidx(j, j)] = ljj;
        CGP_UNROLL for (int i = j + 1; i < D; i++) {
            double t = P[sidx(i, j)];
            CGP_UNROLL for (int k = 0; k < j; k++) t = fma(-L[sidx(i, k)], L[sidx(j, k)], t);
            L[sidx(i, j)] = t / ljj;
        }
    }
}
// x <- (L L^T)^{-1} x with L full-lower
template <int D> CGP_DEV void chol_solve_vec(const double (&L)[D][D], double (&x)[D]) {
    CGP_UNROLL for (int i = 0; i < D; i++) {
        double s = x[i];
        CGP_UNROLL for (int k = 0; k < i; k++) s = fma(-L[i][k], x[k], s);
        x[i] = s / L[i][i];
    }
    CGP_UNROLL for (int i = D - 1; i >= 0; i--) {
        double s = x[i];
        CGP_UNROLL for (int k = i + 1; k < D; k++) s = fma(-L[k][i], x[k], s);
        x[i] = s / L[i][i];
    }
}
// X <- (L L^T)^{-1} Bm, column by column
template <int D> CGP_DEV void chol_solve_mat(const double (&L)[D][D], const double (&Bm)[D][D], double (&X)[D][D]) {
    CGP_UNROLL for (int c = 0; c < D; c++) {
        double col[D];
        CGP_UNROLL for (int i = 0; i < D; i++) col[i] = Bm[i][c];
        chol_solve_vec<D>(L, col);
        CGP_UNROLL for (int i = 0; i < D; i++) X[i][c] = col[i];
    }
}
// rsqrt-scaled Cholesky variants (columns scaled by rsqrt(pivot): <= 2 ulp from the divide-by-sqrt form)
template <int D> CGP_DEV void chol_lower_sym_rsqrt(const double (&P)[NSym<D>::value], double (&L)[NSym<D>::value]) {
    CGP_UNROLL for (int j = 0; j < D; j++) {
        double s = P[sidx(j, j)];
        CGP_UNROLL for (int k = 0; k < j; k++) s = fma(-L[sidx(j, k)], L[sidx(j, k)], s);
        const double r = fast_rsqrt(s);
        L[sidx(j, j)] = s * r;
        CGP_UNROLL for (int i = j + 1; i < D; i++) {
            double t = P[sidx(i, j)];
            CGP_UNROLL for (int k = 0; k < j; k++) t = fma(-L[sidx(i, k)], L[sidx(j, k)], t);
            L[sidx(i, j)] = t * r;
        }
    }
}
// the same without the last pivot: L[D-1][D-1] = piv * fast_rsqrt(piv) is left to the caller (bit-identical; lets the caller
// place the last rsqrt where its latency is hidden)
template <int D> CGP_DEV void chol_lower_sym_rsqrt_head(const double (&P)[NSym<D>::value], double (&L)[NSym<D>::value], double &piv) {
    CGP_UNROLL for (int j = 0; j < D; j++) {
        double s = P[sidx(j, j)];
        CGP_UNROLL for (int k = 0; k < j; k++) s = fma(-L[sidx(j, k)], L[sidx(j, k)], s);
        if (j == D - 1) { piv = s; L[sidx(j, j)] = 0.; break; }
        const double r = fast_rsqrt(s);
        L[sidx(j, j)] = s * r;
        CGP_UNROLL for (int i = j + 1; i < D; i++) {
            double t = P[sidx(i, j)];
            CGP_UNROLL for (int k = 0; k < j; k++) t = fma(-L[sidx(i, k)], L[sidx(j, k)], t);
            L[sidx(i, j)] = t * r;
        }
    }
}
// full-storage variant that also returns 1 / L_jj (for the triangular solves)
template <int D> CGP_DEV void chol_lower_rsqrt(const double (&P)[D][D], double (&L)[D][D], double (&rinv)[D]) {
    CGP_UNROLL for (int j = 0; j < D; j++) {
        double s = P[j][j];
        CGP_UNROLL for (int k = 0; k < j; k++) s = fma(-L[j][k], L[j][k], s);
        const double r = fast_rsqrt(s);
        rinv[j] = r;
        L[j][j] = s * r;
        CGP_UNROLL for (int i = j + 1; i < D; i++) {
            double t = P[i][j];
            CGP_UNROLL for (int k = 0; k < j; k++) t = fma(-L[i][k], L[j][k], t);
            L[i][j] = t * r;
        }
    }
}
template <int D> CGP_DEV void chol_solve_vec_rinv(const double (&L)[D][D], const double (&rinv)[D], double (&x)[D]) {
    CGP_UNROLL for (int i = 0; i < D; i++) {
        double s = x[i];
        CGP_UNROLL for (int k = 0; k < i; k++) s = fma(-L[i][k], x[k], s);
        x[i] = s * rinv[i];
    }
    CGP_UNROLL for (int i = D - 1; i >= 0; i--) {
        double s = x[i];
        CGP_UNROLL for (int k = i + 1; k < D; k++) s = fma(-L[k][i], x[k], s);
        x[i] = s * rinv[i];
    }
}

template <int D> CGP_DEV void sym_to_full(const double (&S)[NSym<D>::value], double (&F)[D][D]) {
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) F[r][c] = S[sidx(r, c)];
}
template <int D> CGP_DEV void lower_to_full(const double (&S)[NSym<D>::value], double (&F)[D][D]) {
    CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) F[r][c] = (c <= r) ? S[sidx(r, c)] : 0.;
}

// -norm.logpdf(y, pred, sqrt(S)) in jax.scipy.stats.norm.logpdf's operation order (filters_smoothers.py:44-45):
// (log(2 pi sc^2) + r^2 / sc^2) / 2 with sc = sqrt(S); sqrt, log and the division through the short-chain routines.
CGP_DEV double nll_increment(double S, double r) {
    const double sc = sqrt(S), sc2 = sc * sc;
    return fma(r * r, fast_rcp(sc2), fast_log_pos(kTwoPi * sc2)) * 0.5;
}
// ------------------------------------------------------------------------------------------------ measurement update
// filters_smoothers.py:55-68; returns the nll increment  (log(2 pi sc^2) + (y - pred)^2 / sc^2) / 2, sc = sqrt(S)
template <int D>
CGP_DEV double linear_update(const double (&mp)[D], const double (&Pp)[D][D], const double (&H)[D], double Xi, double y,
                             double (&mf)[D], double (&Pf)[D][D]) {
    double S = 0.;
    CGP_UNROLL for (int j = 0; j < D; j++) {
        double hp = H[0] * Pp[0][j];
        CGP_UNROLL for (int i = 1; i < D; i++) hp = fma(H[i], Pp[i][j], hp);
        S = (j == 0) ? hp * H[0] : fma(hp, H[j], S);
    }
    S += Xi;
    const double rS = fast_rcp(S);                    // K = Pp h / S with one reciprocal (<= 1 ulp per entry)
    double K[D];
    CGP_UNROLL for (int i = 0; i < D; i++) {
        double s = Pp[i][0] * H[0];
        CGP_UNROLL for (int j = 1; j < D; j++) s = fma(Pp[i][j], H[j], s);
        K[i] = s * rS;
    }
    double pred = H[0] * mp[0];
    CGP_UNROLL for (int i = 1; i < D; i++) pred = fma(H[i], mp[i], pred);
    double r = y - pred;
    CGP_UNROLL for (int i = 0; i < D; i++) mf[i] = fma(K[i], r, mp[i]);
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j < D; j++) Pf[i][j] = Pp[i][j] - (K[i] * K[j]) * S;
    return nll_increment(S, r);
}
// measurement update of ekf_for_kpt (filters_smoothers.py:301-308): as linear_update with the measurement row H = dh/dx(mp)
// and the predicted measurement pred = h(mp) given separately.
template <int D>
CGP_DEV double nonlinear_update(const double (&mp)[D], const double (&Pp)[D][D], const double (&H)[D], double pred, double Xi,
                                double y, double (&mf)[D], double (&Pf)[D][D]) {
    double S = 0.;
    CGP_UNROLL for (int j = 0; j < D; j++) {
        double hp = H[0] * Pp[0][j];
        CGP_UNROLL for (int i = 1; i < D; i++) hp = fma(H[i], Pp[i][j], hp);
        S = (j == 0) ? hp * H[0] : fma(hp, H[j], S);
    }
    S += Xi;
    const double rS = fast_rcp(S);
    double K[D];
    CGP_UNROLL for (int i = 0; i < D; i++) {
        double s = Pp[i][0] * H[0];
        CGP_UNROLL for (int j = 1; j < D; j++) s = fma(Pp[i][j], H[j], s);
        K[i] = s * rS;
    }
    const double r = y - pred;
    CGP_UNROLL for (int i = 0; i < D; i++) mf[i] = fma(K[i], r, mp[i]);
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j < D; j++) Pf[i][j] = Pp[i][j] - (K[i] * K[j]) * S;
    return nll_increment(S, r);
}
// same on packed-symmetric covariances (exactly symmetric inputs stay exactly symmetric)
template <int D>
CGP_DEV double linear_update_sym(const double (&mp)[D], const double (&Pp)[NSym<D>::value], const double (&H)[D], double Xi,
                                 double y, double (&mf)[D], double (&Pf)[NSym<D>::value]) {
    double PH[D];
    CGP_UNROLL for (int i = 0; i < D; i++) {
        double s = Pp[sidx(i, 0)] * H[0];
        CGP_UNROLL for (int j = 1; j < D; j++) s = fma(Pp[sidx(i, j)], H[j], s);
        PH[i] = s;
    }
    double S = PH[0] * H[0];
    CGP_UNROLL for (int j = 1; j < D; j++) S = fma(PH[j], H[j], S);
    S += Xi;
    const double rS = fast_rcp(S);
    double K[D];
    CGP_UNROLL for (int i = 0; i < D; i++) K[i] = PH[i] * rS;
    double pred = H[0] * mp[0];
    CGP_UNROLL for (int i = 1; i < D; i++) pred = fma(H[i], mp[i], pred);
    double r = y - pred;
    CGP_UNROLL for (int i = 0; i < D; i++) mf[i] = fma(K[i], r, mp[i]);
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j <= i; j++)
        Pf[sidx(i, j)] = Pp[sidx(i, j)] - (K[i] * K[j]) * S;
    return nll_increment(S, r);
}

// measurement update on packed-symmetric covariance with one reciprocal; H generic or the unit vector e_1.
// Returns the innovation variance S and residual r = y - H mp; the nll increment is formed from them later
// (nll_increment), off the critical path of the state recursion.
template <int D, bool H_E1>
CGP_DEV void linear_update_fast(const double (&mp)[D], const double (&Pp)[NSym<D>::value], const double (&H)[D], double Xi,
                                double y, double (&mf)[D], double (&Pf)[NSym<D>::value], double &S_out, double &r_out) {
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
    const double rS = fast_rcp(S);
    double K[D];
    CGP_UNROLL for (int i = 0; i < D; i++) K[i] = PH[i] * rS;
    const double r = y - pred;
    CGP_UNROLL for (int i = 0; i < D; i++) mf[i] = fma(K[i], r, mp[i]);
    // Pf = Pp - K K^T S (:66) with K_j S = (Pp h)_j: one fma per entry (K_i (Pp h)_j instead of (K_i K_j) S, <= 1 ulp apart)
    CGP_UNROLL for (int i = 0; i < D; i++) CGP_UNROLL for (int j = 0; j <= i; j++)
        Pf[sidx(i, j)] = fma(-K[i], PH[j], Pp[sidx(i, j)]);
    S_out = S;
    r_out = r;
}

// ------------------------------------------------------------------------------------------------ models
// Softplus and its derivative sharing one exp: g = log(e^x + 1) (naive, as models.py:50), g' = e^x / (e^x + 1).
CGP_DEV void softplus_and_sigmoid(double x, double &gv, double &sg) {
    fast_softplus_sigmoid(x, gv, sg);
}

// Discrete linear model (u, dt) -> (F u, Sigma).  consts = [F | Sigma].
template <int D_> struct ModelLinearDisc {
    static constexpr int D = D_;
    static constexpr int NH = 0;
    static constexpr int kNumConsts = 2 * D_ * D_;
    static constexpr bool kLinear = true;
    struct Trig {};
    double F[D][D], Sg[D][D];
    CGP_DEV void load(const double *__restrict__ c, double /*dt*/) {
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int k = 0; k < D; k++) {
            F[r][k] = c[r * D + k];
            Sg[r][k] = c[D * D + r * D + k];
        }
    }
    static CGP_DEV constexpr bool has_sig(int, int) { return true; }
    static CGP_DEV constexpr bool jnz(int, int) { return true; }        // Jacobian of the mean: dense
    CGP_DEV double sig(int r, int c) const { return Sg[r][c]; }
    CGP_DEV Trig prep(const double (&)[D]) const { return Trig{}; }
    CGP_DEV void mean_with(const Trig &, const double (&u)[D], double (&m)[D]) const { matvec<D>(F, u, m); }
    CGP_DEV void mean(const double (&u)[D], double (&m)[D]) const { matvec<D>(F, u, m); }
    CGP_DEV void mean_tail(const double (&u)[D], double (&m)[D]) const { matvec<D>(F, u, m); }
    CGP_DEV void mean_jac(const double (&u)[D], double (&m)[D], double (&J)[D][D]) const {
        matvec<D>(F, u, m);
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int k = 0; k < D; k++) J[r][k] = F[r][k];
    }
};

// Locally-conditional discretisation of the (harmonic) chirp SDE, NH harmonics, D = 2 NH + 2, V = D - 2.
template <int NH_> struct ModelLCD {
    static constexpr int NH = NH_;
    static constexpr int D = 2 * NH_ + 2;
    static constexpr int V = D - 2;
    static constexpr int kNumConsts = CGP_NC_LCD;
    static constexpr bool kLinear = false;
    struct Trig { double c[NH_], s[NH_]; };      // e-scaled rotation entries: c = cos(theta_k) e, s = sin(theta_k) e
    double e, f00, f01, f10, f11, q, s00, s01, s11, fs, dt;
    CGP_DEV void load(const double *__restrict__ k, double dt_) {
        e = k[0]; f00 = k[1]; f01 = k[2]; f10 = k[3]; f11 = k[4]; q = k[5]; s00 = k[6]; s01 = k[7]; s11 = k[8];
        fs = k[9]; dt = dt_;
    }
    static CGP_DEV constexpr bool has_sig(int r, int c) {
        return (r == c) || (r == V && c == V + 1) || (r == V + 1 && c == V);
    }
    // structural non-zeros of the Jacobian of the mean (mean_jac): the 2 x 2 rotation blocks, their column V, the Matern block
    static CGP_DEV constexpr bool jnz(int r, int c) {
        return r < V ? ((r / 2 == c / 2) || c == V) : (c >= V);
    }
    CGP_DEV double sig(int r, int c) const {
        if (r == c) return r < V ? q : (r == V ? s00 : s11);
        return s01;
    }
    // trig depends on u[V] only (angles dt * k * w, w = 2 pi g(u_V) freq_scale; models.py:296-298, :370-372)
    template <bool WARP_UNIFORM = false> CGP_DEV Trig prep_v(double uv) const {
        return prep_g(WARP_UNIFORM ? fast_softplus_warp(uv) : fast_softplus(uv));
    }
    // the same from gv = g(u_V) (callers that pick the softplus branch themselves)
    CGP_DEV Trig prep_g(double gv) const {
        Trig t;
        double w = (kTwoPi * gv) * fs;
        double s1, c1;
        fast_sincos(dt * w, &s1, &c1);
        double sk = s1, ck = c1;
        CGP_UNROLL for (int k = 0; k < NH; k++) {
            // harmonic k + 1 rotates by (k + 1) theta: angle addition instead of NH sincos evaluations
            // (the reference evaluates cos/sin(dt k w) per harmonic, models.py:371-372; the difference is <= k ulp)
            if (k > 0) {
                const double cn = fma(ck, c1, -(sk * s1)), sn = fma(sk, c1, ck * s1);
                ck = cn; sk = sn;
            }
            t.c[k] = ck * e; t.s[k] = sk * e;
        }
        return t;
    }
    CGP_DEV Trig prep(const double (&u)[D]) const { return prep_v(u[V]); }
    CGP_DEV void mean_with(const Trig &t, const double (&u)[D], double (&m)[D]) const {
        CGP_UNROLL for (int k = 0; k < NH; k++) {
            m[2 * k] = fma(-t.s[k], u[2 * k + 1], t.c[k] * u[2 * k]);
            m[2 * k + 1] = fma(t.c[k], u[2 * k + 1], t.s[k] * u[2 * k]);
        }
        m[V] = fma(f01, u[V + 1], f00 * u[V]);
        m[V + 1] = fma(f11, u[V + 1], f10 * u[V]);
    }
    CGP_DEV void mean(const double (&u)[D], double (&m)[D]) const { mean_with(prep(u), u, m); }
    // only the Matern rows (the chirp rows do not depend on u[V + 1])
    CGP_DEV void mean_tail(const double (&u)[D], double (&m)[D]) const {
        m[V] = fma(f01, u[V + 1], f00 * u[V]);
        m[V + 1] = fma(f11, u[V + 1], f10 * u[V]);
    }
    // closed form of jax.jacfwd(lambda u: cond_m_cov(u, dt)[0]) (filters_smoothers.py:255, :342)
    CGP_DEV void mean_jac(const double (&u)[D], double (&m)[D], double (&J)[D][D]) const {
        double gv, sg;
        softplus_and_sigmoid(u[V], gv, sg);
        double w = (kTwoPi * gv) * fs, dw = (kTwoPi * sg) * fs;
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) J[r][c] = 0.;
        CGP_UNROLL for (int k = 0; k < NH; k++) {
            double sn, cs, dtk = dt * (double)(k + 1);
            fast_sincos(dtk * w, &sn, &cs);
            double ce = cs * e, se = sn * e, dth = dtk * dw;
            double u0 = u[2 * k], u1 = u[2 * k + 1];
            m[2 * k] = fma(-se, u1, ce * u0);
            m[2 * k + 1] = fma(ce, u1, se * u0);
            J[2 * k][2 * k] = ce;     J[2 * k][2 * k + 1] = -se;
            J[2 * k + 1][2 * k] = se; J[2 * k + 1][2 * k + 1] = ce;
            J[2 * k][V] = -m[2 * k + 1] * dth;      // e(-s u0 - c u1) dtheta/du_V
            J[2 * k + 1][V] = m[2 * k] * dth;       // e( c u0 - s u1) dtheta/du_V
        }
        m[V] = fma(f01, u[V + 1], f00 * u[V]);
        m[V + 1] = fma(f11, u[V + 1], f10 * u[V]);
        J[V][V] = f00; J[V][V + 1] = f01; J[V + 1][V] = f10; J[V + 1][V + 1] = f11;
    }
};

// Linear SDE drift u -> A u.  consts = [A].
template <int D_> struct ModelLinearSDE {
    static constexpr int D = D_;
    static constexpr int NH = 0;
    static constexpr int kNumConsts = D_ * D_;
    static constexpr bool kLinear = true;
    double A[D][D];
    CGP_DEV void load(const double *__restrict__ c) {
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int k = 0; k < D; k++) A[r][k] = c[r * D + k];
    }
    CGP_DEV void drift(const double (&u)[D], double (&a)[D]) const { matvec<D>(A, u, a); }
    CGP_DEV void drift_jac(const double (&u)[D], double (&a)[D], double (&J)[D][D]) const {
        matvec<D>(A, u, a);
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int k = 0; k < D; k++) J[r][k] = A[r][k];
    }
};

// (Harmonic) chirp SDE drift a(u) = A(u_V) u (models.py:104-110, :164-168).  consts = [lam, gamma^2, 2 gamma, fs].
template <int NH_> struct ModelSDE {
    static constexpr int NH = NH_;
    static constexpr int D = 2 * NH_ + 2;
    static constexpr int V = D - 2;
    static constexpr int kNumConsts = CGP_NC_SDE;
    static constexpr bool kLinear = false;
    double lam, g2, tg, fs;
    CGP_DEV void load(const double *__restrict__ k) { lam = k[0]; g2 = k[1]; tg = k[2]; fs = k[3]; }
    CGP_DEV void drift_w(double w, const double (&u)[D], double (&a)[D]) const {
        CGP_UNROLL for (int k = 0; k < NH; k++) {
            double wk = w * (double)(k + 1);
            a[2 * k] = fma(-wk, u[2 * k + 1], -lam * u[2 * k]);
            a[2 * k + 1] = fma(-lam, u[2 * k + 1], wk * u[2 * k]);
        }
        a[V] = u[V + 1];
        a[V + 1] = fma(-tg, u[V + 1], -g2 * u[V]);
    }
    CGP_DEV double omega(double uv) const { return (kTwoPi * fast_softplus(uv)) * fs; }
    CGP_DEV void drift(const double (&u)[D], double (&a)[D]) const { drift_w(omega(u[V]), u, a); }
    // closed form of jax.jacfwd(a) (filters_smoothers.py:382, :425)
    CGP_DEV void drift_jac(const double (&u)[D], double (&a)[D], double (&J)[D][D]) const {
        double gv, sg;
        softplus_and_sigmoid(u[V], gv, sg);
        double w = (kTwoPi * gv) * fs, dw = (kTwoPi * sg) * fs;
        drift_w(w, u, a);
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) J[r][c] = 0.;
        CGP_UNROLL for (int k = 0; k < NH; k++) {
            double wk = w * (double)(k + 1), dwk = dw * (double)(k + 1);
            J[2 * k][2 * k] = -lam;     J[2 * k][2 * k + 1] = -wk;
            J[2 * k + 1][2 * k] = wk;   J[2 * k + 1][2 * k + 1] = -lam;
            J[2 * k][V] = -dwk * u[2 * k + 1];
            J[2 * k + 1][V] = dwk * u[2 * k];
        }
        J[V][V + 1] = 1.;
        J[V + 1][V] = -g2;
        J[V + 1][V + 1] = -tg;
    }
};

// ------------------------------------------------------------------------------------------------ lane-group reductions
// Butterfly all-reduce over a group of G consecutive lanes (G power of two <= 32).  Every lane ends with the
// bit-identical sum (IEEE addition commutes), so replicated per-lane state never diverges.
template <int G> CGP_DEV double group_allreduce(double v) {
    CGP_UNROLL for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Shared-memory variant for many values per lane: NA values x G lanes are transposed through `sm`
// ([NA][G + 1] partials followed by NA results), every lane sums NA / G of them in a fixed tree order and all lanes
// read the totals back.  ~ (NA + 2 NA / G * G/2 ...) instructions instead of 3 NA log2(G) for the butterfly.
template <int NA, int G> struct GroupSmem { static constexpr int kDoubles = NA * (G + 1) + ((NA + 1) & ~1); };
template <int NA, int G> CGP_DEV void group_sum_smem(double (&a)[NA], double *sm, int lane) {
    constexpr int PITCH = G + 1;
    double *res = sm + NA * PITCH;
    CGP_UNROLL for (int k = 0; k < NA; k++) sm[k * PITCH + lane] = a[k];
    __syncwarp();
    CGP_UNROLL for (int k0 = 0; k0 < NA; k0 += G) {
        const int k = k0 + lane;
        const bool ok = k < NA;
        double v[G];
        CGP_UNROLL for (int j = 0; j < G; j++) v[j] = sm[(ok ? k : 0) * PITCH + j];
        CGP_UNROLL for (int w2 = 1; w2 < G; w2 <<= 1)
            CGP_UNROLL for (int j = 0; j + w2 < G; j += 2 * w2) v[j] += v[j + w2];
        if (ok) res[k] = v[0];
    }
    __syncwarp();
    CGP_UNROLL for (int k = 0; k < NA; k++) a[k] = res[k];
    __syncwarp();
}

}  // namespace cgp
