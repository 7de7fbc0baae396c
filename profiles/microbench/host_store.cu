// Can a latency-bound kernel stream its (small) results straight into pinned host memory while it runs, i.e. does the
// device->host transfer of a readout overlap the smoother sweep that produces it?
// Mimics the sweep of config 2: 500 warps (2 chirps each), 3141 steps in tiles of 16; every tile takes `spin` cycles of
// dependent work and then each half-warp writes 2 x 128 bytes (16 steps x {freq, v_var}).  50 MB in total.
//   a) destination in device memory                       -> kernel time
//   b) destination in device memory + cudaMemcpyAsync D2H -> kernel + copy (what the product did before)
//   c) destination = mapped pinned host memory            -> kernel time with the stores crossing PCIe as they are issued
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o host_store host_store.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(32) sweep_like(double *freq, double *var, int B, int T, int spin) {
    const int lane = threadIdx.x, half = lane >> 4, l = lane & 15;
    const long b = (long)blockIdx.x * 2 + half;
    if (b >= B) return;
    double x = 1. + lane;
    for (long hi = T; hi > 0; hi -= 16) {
        const long lo = hi >= 16 ? hi - 16 : 0;
        const long long t0 = clock64();
        while (clock64() - t0 < spin) x = fma(x, 1.0000001, 1e-9);
        if (lo + l < hi) {
            freq[b * T + lo + l] = x;
            var[b * T + lo + l] = x + 1.;
        }
    }
}

int main(int argc, char **argv) {
    const int B = 1000, T = 3141;
    const size_t n = (size_t)B * T;
    double *d_f, *d_v, *h_f, *h_v, *h2_f, *h2_v;
    cudaMalloc(&d_f, n * 8); cudaMalloc(&d_v, n * 8);
    cudaHostAlloc(&h_f, n * 8, cudaHostAllocDefault); cudaHostAlloc(&h_v, n * 8, cudaHostAllocDefault);
    cudaHostAlloc(&h2_f, n * 8, cudaHostAllocDefault); cudaHostAlloc(&h2_v, n * 8, cudaHostAllocDefault);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int spins[] = {0, 2400, 4800, 9600};
    for (int spin : spins) {
        float ta = 1e9f, tb = 1e9f, tc = 1e9f, ms;
        for (int rep = 0; rep < 6; rep++) {
            cudaEventRecord(e0);
            sweep_like<<<(B + 1) / 2, 32>>>(d_f, d_v, B, T, spin);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < ta) ta = ms;
            cudaEventRecord(e0);
            sweep_like<<<(B + 1) / 2, 32>>>(d_f, d_v, B, T, spin);
            cudaMemcpyAsync(h2_f, d_f, n * 8, cudaMemcpyDeviceToHost); cudaMemcpyAsync(h2_v, d_v, n * 8, cudaMemcpyDeviceToHost);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < tb) tb = ms;
            cudaEventRecord(e0);
            sweep_like<<<(B + 1) / 2, 32>>>(h_f, h_v, B, T, spin);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < tc) tc = ms;
        }
        double bad = 0;
        for (size_t i = 0; i < n; i++) bad += (h_f[i] != h2_f[i]) + (h_v[i] != h2_v[i]);
        printf("spin %5d cycles/tile: device dst %.3f ms | device dst + 2 x D2H copy %.3f ms | pinned-host dst %.3f ms (%.1f GB/s)  mismatches %.0f\n",
               spin, ta, tb, tc, 2 * n * 8 / tc / 1e6, bad);
    }
    return 0;
}
