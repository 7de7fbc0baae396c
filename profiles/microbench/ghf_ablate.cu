// ghf_ablate.cu -- timing-only ablations of ghf_filter_kernel (numerically meaningless when DBG != 0):
// which part of the per-step dependency chain costs what.  1 warp per SM (B = 148) => pure latency.
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../chirpgp_b200/csrc/cgp_fast.cuh"
using namespace cgp;
template <int DBG> float run(const CgpProblem &p, const FilterIO &io) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gh_warp_filter_kernel<GhPredictLCD<1, 3, DBG>, false, true><<<(unsigned)p.B, 32>>>(p, io);
    cudaEventRecord(e0);
    gh_warp_filter_kernel<GhPredictLCD<1, 3, DBG>, false, true><<<(unsigned)p.B, 32>>>(p, io);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("DBG=%d  %8.3f ms  %7.0f cycles/step  (%s)\n", DBG, ms, ms * 1e-3 * 1.965e9 / p.T, cudaGetErrorString(cudaGetLastError()));
    return ms;
}
int main(int argc, char **argv) {
    const int64_t B = argc > 1 ? atoi(argv[1]) : 148, T = 3141;
    const int D = 4, n = 81;
    // GH order-3 table (symmetric nodes are fine for timing)
    std::vector<double> w(n), xi(n * D);
    const double r1[3] = {0., 1.7320508075688772, -1.7320508075688772}, w1[3] = {2. / 3, 1. / 6, 1. / 6};
    for (int i = 0; i < n; i++) { double ww = 1; int t = i; for (int j = 0; j < D; j++) { xi[i * D + j] = r1[t % 3]; ww *= w1[t % 3]; t /= 3; } w[i] = ww; }
    double consts[10] = {0.9999, 1.0, 9.98e-4, -2.99e-3, 0.9965, 1e-5, 6.9e-9, 1.03e-5, 2.07e-2, 1.0};
    double m0[4] = {0, 0, 7, 0}, P0[16] = {0.1, 0, 0, 0, 0, 0.1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 3}, H[4] = {0, 1, 0, 0};
    std::vector<double> ys(B * T);
    for (size_t i = 0; i < ys.size(); i++) ys[i] = sin(0.05 * (i % T));
    double *dw, *dxi, *dc, *dm0, *dP0, *dH, *dys, *mfs, *Pfs, *nell;
    cudaMalloc(&dw, n * 8); cudaMalloc(&dxi, n * D * 8); cudaMalloc(&dc, 80); cudaMalloc(&dm0, 32); cudaMalloc(&dP0, 128); cudaMalloc(&dH, 32);
    cudaMalloc(&dys, B * T * 8); cudaMalloc(&mfs, B * T * 4 * 8); cudaMalloc(&Pfs, B * T * 16 * 8); cudaMalloc(&nell, B * T * 8);
    cudaMemcpy(dw, w.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dxi, xi.data(), n * D * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dc, consts, 80, cudaMemcpyHostToDevice); cudaMemcpy(dm0, m0, 32, cudaMemcpyHostToDevice);
    cudaMemcpy(dP0, P0, 128, cudaMemcpyHostToDevice); cudaMemcpy(dH, H, 32, cudaMemcpyHostToDevice);
    cudaMemcpy(dys, ys.data(), B * T * 8, cudaMemcpyHostToDevice);
    CgpProblem p = {};
    p.B = B; p.T = T; p.model = CGP_MODEL_LCD; p.d = 4; p.num_harmonics = 1; p.n_sigma = n; p.sigma_kind = 1; p.gh_order = 3; p.ys_repeat = 1;
    p.consts = dc; p.m0 = dm0; p.P0 = dP0; p.H = dH; p.sig_w = dw; p.sig_xi = dxi; p.Xi = 0.1; p.dt = 1e-3;
    FilterIO io = {dys, mfs, Pfs, nell, 0};
    printf("B = %ld\n", (long)B);
    run<0>(p, io); run<1>(p, io); run<2>(p, io); run<3>(p, io);
    FilterIO io2 = {dys, nullptr, nullptr, nell, 1};
    printf("nll-only: "); run<0>(p, io2);
    return 0;
}
