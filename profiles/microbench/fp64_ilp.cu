// fp64_ilp.cu -- can ONE warp overlap independent DFMA chains?  cycles per DFMA for K independent chains, for 1..4
// warps resident on one SMSP (blockDim = 32 * 4 * W puts W warps on each of the 4 SMSPs).
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int K> __global__ void chains(double *out, long long *cyc, double a, double b) {
    double x[K];
#pragma unroll
    for (int k = 0; k < K; k++) x[k] = 1.0 + threadIdx.x * 1e-3 + k;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int k = 0; k < K; k++) x[k] = fma(x[k], a, b);
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; k++) s += x[k];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int K> void run(int warps_per_smsp) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    chains<K><<<1, 32 * 4 * warps_per_smsp>>>(out, cyc, 0.999, 1e-3);
    chains<K><<<1, 32 * 4 * warps_per_smsp>>>(out, cyc, 0.999, 1e-3);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SMSP %d  chains %d : %6.2f cycles per DFMA per warp  (%.2f cycles per DFMA per SMSP)\n", warps_per_smsp, K,
           (double)h / (N * K), (double)h / (N * K * warps_per_smsp));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w = 1; w <= 4; w *= 2) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
