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
int main()
{
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
