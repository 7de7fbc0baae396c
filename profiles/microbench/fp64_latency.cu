// fp64_latency.cu -- dependent-issue latency (cycles) of the FP64 building blocks the filters are made of, measured
// with clock64() on one warp.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N 4096
template <int OP> __device__ __forceinline__ double step(double x, double a, double b) {
    if (OP == 0) return fma(x, a, b);
    if (OP == 1) return x * a;
    if (OP == 2) return x + b;
    if (OP == 3) return rsqrt(x) + b;            // b ~ 1 keeps x in range
    if (OP == 4) return 1. / x + b;
    if (OP == 5) return sqrt(x) + b;
    if (OP == 6) return exp(x * 1e-3);
    if (OP == 7) return log(x + 2.);
    if (OP == 8) { double s, c; sincos(x, &s, &c); return s + c * 0.5; }
    if (OP == 9) return log(exp(x) + 1.);        // softplus
    if (OP == 10) return __shfl_xor_sync(0xffffffffu, x, 1) + b;
    if (OP == 11) return x / a + b;
    return x;
}
template <int OP> __global__ void lat(double *out, long long *cyc, double a, double b) {
    double x = 1.0 + threadIdx.x * 1e-3;
    for (int i = 0; i < 64; i++) x = step<OP>(x, a, b);
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = step<OP>(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput: 8 independent chains, many warps
template <int OP> __global__ void thr(double *out, double a, double b) {
    double x[8];
    for (int k = 0; k < 8; k++) x[k] = 1.0 + threadIdx.x * 1e-3 + k;
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = step<OP>(x[k], a, b);
    double s = 0;
    for (int k = 0; k < 8; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char *name, double a, double b) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 8 * 256 * 8); cudaMalloc(&cyc, 8);
    lat<OP><<<1, 32>>>(out, cyc, a, b);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    thr<OP><<<148 * 8, 256>>>(out, a, b);
    cudaEventRecord(e0);
    thr<OP><<<148 * 8, 256>>>(out, a, b);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148. * 8 * 256 * 8 * N;
    printf("%-12s dependent latency %7.1f cycles/op   throughput %8.2f Gop/s (%.2f warp-op/clk/SM @1.965GHz)\n", name, (double)h / N,
           ops / ms / 1e6, ops / 32 / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("dfma", 0.999, 1e-3);
    run<1>("dmul", 1.0000001, 0);
    run<2>("dadd", 0, 1e-9);
    run<3>("rsqrt+add", 0, 1.0);
    run<4>("rcp+add", 0, 1.0);
    run<5>("sqrt+add", 0, 1.0);
    run<6>("exp", 0, 0);
    run<7>("log", 0, 0);
    run<8>("sincos", 0, 0);
    run<9>("softplus", 0, 0);
    run<10>("shfl64+add", 0, 1e-9);
    run<11>("div+add", 1.0000001, 1e-9);
    return 0;
}
