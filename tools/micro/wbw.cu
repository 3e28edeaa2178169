// Write-bandwidth ceilings on the B200: cudaMemset, plain STG.128 streaming, and TMA bulk stores issued
// the way the evaluator issues them (one warp per CTA, 8.5 KB chunks from shared memory, 2 in flight).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wbw wbw.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void stg_kernel(double2* __restrict__ dst, size_t n16)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = make_double2(1.0, 2.0);
}

template <int NBUF>
__global__ void __launch_bounds__(32) tma_kernel(double* __restrict__ dst, size_t nchunks, int chunk_doubles)
{
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x;
    for (int i = lane; i < NBUF * chunk_doubles; i += 32) sm[i] = (double)i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
    int b = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * chunk_doubles),
                         "r"(sbase + (unsigned)b * chunk_doubles * 8u), "r"((unsigned)chunk_doubles * 8u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        b = (b + 1) % NBUF;
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Pattern of the evaluator: every warp owns whole rows of `row_doubles` (one evaluation's Jacobian values)
// and writes them as consecutive chunks; rows are dealt round-robin, so G write heads sit row_doubles apart.
template <int NBUF>
__global__ void __launch_bounds__(32) tma_rows_kernel(double* __restrict__ dst, size_t nrows, int row_doubles, int chunk_doubles,
                                                       int rows_per_group)
{
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x;
    for (int i = lane; i < NBUF * chunk_doubles; i += 32) sm[i] = (double)i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
    int b = 0;
    // rows_per_group == 1: row r = blockIdx + i*grid.  rows_per_group == g: g consecutive CTAs share a row
    // (each takes every g-th chunk), so the number of concurrent write heads drops by g.
    const int g = rows_per_group;
    const size_t nchunk_row = row_doubles / chunk_doubles;
    for (size_t r = blockIdx.x / g; r < nrows; r += gridDim.x / g) {
        for (size_t c = blockIdx.x % g; c < nchunk_row; c += g) {
            if (lane == 0) {
                asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + r * row_doubles + c * chunk_doubles),
                             "r"(sbase + (unsigned)b * chunk_doubles * 8u), "r"((unsigned)chunk_doubles * 8u) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            b = (b + 1) % NBUF;
            __syncwarp();
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
    const size_t bytes = (size_t)4 << 30;
    double* d;
    CK(cudaMalloc(&d, bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    auto report = [&](const char* name) {
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms / 5, bytes / (ms / 5 * 1e-3) / 1e9);
    };
    for (int w = 0; w < 2; ++w) CK(cudaMemset(d, 0, bytes));
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) CK(cudaMemsetAsync(d, 0, bytes));
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    report("cudaMemset");
    for (int tpb : {256, 1024}) {
        stg_kernel<<<148 * (2048 / tpb), tpb>>>((double2*)d, bytes / 16);
        cudaEventRecord(e0);
        for (int r = 0; r < 5; ++r) stg_kernel<<<148 * (2048 / tpb), tpb>>>((double2*)d, bytes / 16);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        char nm[64]; snprintf(nm, 64, "STG.128 grid-stride, %d thr/CTA", tpb);
        report(nm);
    }
    for (int cd : {529, 1058, 2116, 4232}) {
        const int chunk = cd & ~1;
        const size_t nchunks = bytes / 8 / chunk;
        for (int ctas : {4, 8, 16}) {
            const size_t smem = 2 * chunk * 8;
            CK(cudaFuncSetAttribute(tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (smem * ctas > 220 * 1024) continue;
            tma_kernel<2><<<148 * ctas, 32, smem>>>(d, nchunks, chunk);
            cudaEventRecord(e0);
            for (int r = 0; r < 5; ++r) tma_kernel<2><<<148 * ctas, 32, smem>>>(d, nchunks, chunk);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            char nm[96]; snprintf(nm, 96, "TMA bulk store %5d B x2 bufs, %2d warps/SM", chunk * 8, ctas);
            report(nm);
        }
    }
    {
        const int chunk = 1056, row = chunk * 30;     // ~253 KB rows of 30 chunks, like one evaluation
        const size_t nrows = bytes / 8 / row;
        const size_t smem = 2 * chunk * 8;
        CK(cudaFuncSetAttribute(tma_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int g : {1, 2, 4, 8, 30}) {
            for (int ctas : {8}) {
                const int grid = 148 * ctas / g * g;
                tma_rows_kernel<2><<<grid, 32, smem>>>(d, nrows, row, chunk, g);
                cudaEventRecord(e0);
                for (int r = 0; r < 5; ++r) tma_rows_kernel<2><<<grid, 32, smem>>>(d, nrows, row, chunk, g);
                cudaEventRecord(e1); CK(cudaDeviceSynchronize());
                cudaEventElapsedTime(&ms, e0, e1);
                const double wbytes = (double)nrows * row * 8;
                printf("TMA rows of %d B, chunk %d B, %d warps/SM, %2d warps per row: %8.3f ms %8.1f GB/s\n", row * 8, chunk * 8, ctas, g,
                       ms / 5, wbytes / (ms / 5 * 1e-3) / 1e9);
            }
        }
    }
    CK(cudaGetLastError());
    return 0;
}
