// FP64 pipe peak on the B200: dependent-chain DFMA streams, 8 independent accumulators per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma dfma.cu
#include <cuda_runtime.h>
#include <cstdio>
__global__ void dfma_kernel(double* out, double a, double b, int iters)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
// dependent-chain latency of DADD / DMUL / DFMA: one warp per SM, 4096 dependent operations, clock64 around them
template <int OP>
__global__ void chain_kernel(double* out, long long* cyc, double a, double b)
{
    double x = threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 4096; ++i) {
        if (OP == 0) x = __dadd_rn(x, a);
        else if (OP == 1) x = __dmul_rn(x, a);
        else x = __fma_rn(x, a, b);
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
static void chain(const char* name, double* d, int warps)
{
    long long* c; cudaMalloc(&c, 148 * sizeof(long long));
    chain_kernel<OP><<<148, 32 * warps>>>(d, c, 1.0000001, 1e-9);
    chain_kernel<OP><<<148, 32 * warps>>>(d, c, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
    printf("%s dependent chain, %2d warps/SM: %.2f cycles per operation\n", name, warps, h[0] / 4096.0);
    cudaFree(c);
}
int main()
{
    {
        double* d0; cudaMalloc(&d0, 148 * 1024 * sizeof(double));
        for (int w : {1, 4, 8, 16}) { chain<0>("DADD", d0, w); chain<1>("DMUL", d0, w); chain<2>("DFMA", d0, w); }
        cudaFree(d0);
    }
    double* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(double));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 16;
    for (int ctas : {4, 8}) {
        dfma_kernel<<<148 * ctas, 256>>>(d, 1.0000001, 1e-9, iters);
        cudaEventRecord(e0);
        dfma_kernel<<<148 * ctas, 256>>>(d, 1.0000001, 1e-9, iters);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * iters * 148.0 * ctas * 256;
        printf("DFMA %d CTAs/SM x 256 thr: %.3f ms  %.2f TFLOP/s fp64 (%.2f T DFMA/s)\n", ctas, ms, flops / ms / 1e9, flops / 2 / ms / 1e9);
    }
    return 0;
}
