// FP64 pipe cost of partially filled warps on B200: does a DFMA whose upper half-warp (or more) is inactive occupy the
// 16-lane FP64 unit of an SM sub-partition for one pass instead of two?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lanes fp64_lanes.cu && ./fp64_lanes
// Every SM sub-partition gets WPS warps; in each warp only lanes < L run ILP independent DFMA chains.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double *out, int iters, int L) {
    const int lane = threadIdx.x & 31;
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double b = 1.0000001, c = 1e-9;
    if (lane < L) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) a[i] = fma(a[i], b, c);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    double *out;
    cudaMalloc(&out, 148 * 1024 * sizeof(double) * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps : {1, 2, 4}) {
        for (int L : {32, 16, 8, 1}) {
            const int block = 128 * wps;     // wps warps per sub-partition
            k<8><<<148, block>>>(out, 100, L);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k<8><<<148, block>>>(out, iters, L);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double instr = (double)iters * 8;                       // DFMA warp-instructions per warp
            const double cyc = ms * 1e-3 * 1.965e9;
            printf("warps/SMSP=%d active lanes=%2d  %.3f ms  %.2f cycles per warp-DFMA per SMSP-warp  (%.2f TFLOP/s useful)\n", wps, L, ms,
                   cyc / (instr * wps), 148.0 * block / 32 * L * instr * 2 / (ms * 1e-3) / 1e12);
        }
    }
    return 0;
}
