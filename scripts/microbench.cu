// microbench.cu -- B200 fp64 / shared-memory ceilings that bound the fused gate kernel.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_dmul_dadd(double* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = x[i] * a; x[i] = x[i] + b; }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_lds(double* out, int iters) {
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = make_double2(i, -i);
    __syncthreads();
    double2 acc = make_double2(0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) { double2 v = sm[(idx + k * 256) & 4095]; acc.x += v.x; acc.y += v.y; }
        idx = (idx + 1) & 4095;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 16 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int ctas_per_sm : {1, 2, 4, 8}) {
        int grid = 148 * ctas_per_sm, iters = 20000;
        k_dfma<<<grid, 256>>>(out, 100, 1.0000001, 1e-9);
        cudaEventRecord(e0); k_dfma<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8 * iters * 256.0 * grid;
        printf("DFMA  ctas/SM=%d: %.2f TFLOP/s (%.1f DFMA/clk/SM at 1.965 GHz)\n", ctas_per_sm, flops / ms / 1e9,
               flops / 2 / (ms * 1e-3) / 148 / 1.965e9);
        cudaEventRecord(e0); k_dmul_dadd<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMUL+DADD ctas/SM=%d: %.2f Tinstr-lanes/s (%.1f ops/clk/SM)\n", ctas_per_sm, flops / ms / 1e9,
               flops / (ms * 1e-3) / 148 / 1.965e9);
    }
    cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int ctas_per_sm : {1, 2, 3}) {
        int grid = 148 * ctas_per_sm, iters = 2000;
        k_lds<<<grid, 256, 65536>>>(out, 10);
        cudaEventRecord(e0); k_lds<<<grid, 256, 65536>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double bytes = 16.0 * 16 * iters * 256.0 * grid;
        printf("LDS.128 ctas/SM=%d: %.1f B/clk/SM\n", ctas_per_sm, bytes / (ms * 1e-3) / 148 / 1.965e9);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
