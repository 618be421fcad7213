// Micro-benchmark: cycles per tcgen05.mma (kind::f16, SS, no-swizzle K-major) for several shapes.
#include <cstdio>
#include <cuda_runtime.h>
#include "../metagenomic-deepfri_b200/csrc/tc_ptx.cuh"
using namespace mdf::tc;

template <int M, int N>
__global__ void bench(long long *out, int reps, uint32_t lbo_a, uint32_t lbo_b)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_f16(M, N);
        const uint64_t ad = umma_smem_desc(smem_u32(smem), lbo_a, 128);
        const uint64_t bd = umma_smem_desc(smem_u32(smem) + 64 * 1024, lbo_b, 128);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll 8
            for (int ks = 0; ks < 16; ++ks)
                umma_f16(slot, ad + (uint64_t)(ks * (2 * lbo_a >> 4)), bd + (uint64_t)(ks * (2 * lbo_b >> 4)), idesc, ks != 0);
        }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(slot);
}

template <int M, int N>
void run(long long *d)
{
    const int reps = 16;
    cudaFuncSetAttribute(bench<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    // A: M rows x 256 K -> lbo = (M/8)*128 ; B: N rows x 256 K -> lbo = (N/8)*128
    bench<M, N><<<1, 128, 200 * 1024>>>(d, reps, (M / 8) * 128, (N / 8) * 128);
    long long h[2];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("M=%3d N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%.0f MAC/cyc) %s\n", M, N, h[0] / (16.0 * reps), h[1] / (16.0 * reps),
           (double)M * N * 16 * 16 * reps / h[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    long long *d;
    cudaMalloc(&d, 64);
    run<64, 32>(d); run<64, 64>(d); run<64, 128>(d); run<64, 256>(d);
    run<128, 32>(d); run<128, 64>(d); run<128, 128>(d); run<128, 256>(d);
    run<64, 32>(d);
    return 0;
}
