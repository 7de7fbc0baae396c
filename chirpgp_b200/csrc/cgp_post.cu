// cgp_post.cu -- post-processing step right after the smoothers:
// chirpgp.quadratures.gaussian_expectation (quadratures.py:234-274) for its default integrand func = g (the softplus of
// models.py:50) and d = 1, i.e. the instantaneous-frequency estimate  E[g(V_k)],  V_k ~ N(ms_k, chol_k^2),  that every demo
// / job forms from the smoothing result (demos/ghfs_mle.py:87-89: ms = smoothing_mean[:, 2], chol = sqrt(smoothing_cov[:, 2, 2])).
// One thread per time step, the Gauss-Hermite table (quadratures.py:157-196, d = 1: xi = sqrt(2) roots, w = w1d / sqrt(pi))
// is passed as data.  HBM-bound: 16 B in (strided reads straight out of mss / Pss are allowed), 8 B out per step.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/chirpgp_b200.h"
#include "cgp_math.cuh"

namespace {
constexpr int kMaxOrder = 64;
struct GhTable { double w[kMaxOrder], xi[kMaxOrder]; };      // by value: lives in the kernel-parameter constant bank

__global__ void __launch_bounds__(256) expect_softplus_kernel(int64_t n, const double *__restrict__ ms, int64_t ms_stride,
                                                              const double *__restrict__ sd, int64_t sd_stride, int sd_is_variance,
                                                              int order, const __grid_constant__ GhTable tab,
                                                              double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m = ms[i * ms_stride];
    double c = sd[i * sd_stride];
    if (sd_is_variance) c = sqrt(c);
    // expectation_from_nodes (quadratures.py:203-216): sum_s w_s func(chi_s), chi_s = m + chol xi_s (:201), in table order
    double acc = 0.;
    for (int s = 0; s < order; s++) acc = fma(tab.w[s], cgp::fast_softplus(fma(c, tab.xi[s], m)), acc);
    out[i] = acc;
}
}  // namespace

extern "C" int cgp_gaussian_expectation_softplus_f64(int64_t n, const double *ms, int64_t ms_stride, const double *sd,
                                                     int64_t sd_stride, int sd_is_variance, const double *w_host,
                                                     const double *xi_host, int order, double *out, void *stream) {
    if (n < 1 || !ms || !sd || !out || !w_host || !xi_host || ms_stride < 1 || sd_stride < 1) return CGP_ERR_BAD_ARG;
    if (order < 1 || order > kMaxOrder) return CGP_ERR_UNSUPPORTED;
    GhTable tab;
    for (int s = 0; s < kMaxOrder; s++) { tab.w[s] = s < order ? w_host[s] : 0.; tab.xi[s] = s < order ? xi_host[s] : 0.; }
    expect_softplus_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, ms, ms_stride, sd, sd_stride,
                                                                                         sd_is_variance, order, tab, out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}
