// cgp_peak.cu -- DFMA-only microbenchmark used as the FP64 roofline denominator (MEASURED_PEAKS.json has no
// FP64 figure).  8 independent FMA chains per thread; flops = threads * iters * 8 * 2.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {
__global__ void __launch_bounds__(256) dfma_chain_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3., x4 = x0 + 4., x5 = x0 + 5., x6 = x0 + 6.,
           x7 = x0 + 7.;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace

extern "C" {
// Launches the DFMA kernel on `stream`; `out` must hold blocks * 256 doubles.  Returns flops issued, or < 0.
double cgp_bench_dfma(double *out, int blocks, int iters, void *stream) {
    if (!out || blocks < 1 || iters < 1) return -1.;
    dfma_chain_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, iters, 0.999999, 1e-6);
    if (cudaGetLastError() != cudaSuccess) return -2.;
    return (double)blocks * 256. * (double)iters * 16.;
}
}

// ---- test hook for cgp_math.cuh: evaluates one of the fast elementary functions element-wise (device pointers)
#include "cgp_math.cuh"
namespace {
__global__ void math_probe_kernel(int kind, int64_t n, const double *__restrict__ x, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    double r = 0., s, c;
    switch (kind) {
        case 0: r = cgp::fast_exp(v); break;
        case 1: r = cgp::fast_softplus(v); break;
        case 2: cgp::fast_sincos(v, &s, &c); r = s; break;
        case 3: cgp::fast_sincos(v, &s, &c); r = c; break;
        case 4: r = cgp::fast_rsqrt(v); break;
        case 5: r = cgp::fast_rcp(v); break;
        case 6: cgp::fast_softplus_sigmoid(v, s, c); r = c; break;
        case 7: cgp::fast_softplus_sigmoid(v, s, c); r = s; break;
        default: r = v;
    }
    out[i] = r;
}
}  // namespace
extern "C" int cgp_test_math(int kind, int64_t n, const double *x, double *out, void *stream) {
    if (n < 1 || !x || !out) return -1;
    math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, n, x, out);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
