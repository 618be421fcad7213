// Micro-benchmark: cycles per tcgen05.mma for the shapes the fused LSTM kernel issues, operands resident in
// shared memory in the 128-row x 64-k tile layout (LBO 2048, SBO 128):
//   cta_group::1  M=128 N=128 / N=256   (one CTA)
//   cta_group::2  M=256 N=256 / N=128   (CTA pair, leader issues)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pair_bench tools/pair_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../metagenomic-deepfri_b200/csrc/tc_ptx.cuh"
using namespace mdf::tc;

template <bool PAIR, int N>
__global__ void bench(long long *out, int reps)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    const bool leader = !PAIR || cluster_ctarank() == 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) { if (PAIR) tmem_alloc_pair<512>(&slot); else tmem_alloc<512>(&slot); }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    tcgen05_fence_after();
    if (threadIdx.x == 0 && leader) {
        const uint32_t idesc = umma_idesc_f16(PAIR ? 256 : 128, N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 128 * 1024;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int kb = 0; kb < 8; ++kb) {                   // A: 8 tiles of 16 KiB, B: ring of 4 tiles
                const uint64_t ad = umma_smem_desc(a0 + kb * TILE_BYTES, TILE_LBO, TILE_SBO);
                const uint64_t bd = umma_smem_desc(b0 + (kb & 3) * TILE_BYTES, TILE_LBO, TILE_SBO);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (PAIR) umma_f16_pair(slot, ad + (uint64_t)(ks * 256), bd + (uint64_t)(ks * 256), idesc, (kb | ks) != 0);
                    else umma_f16(slot, ad + (uint64_t)(ks * 256), bd + (uint64_t)(ks * 256), idesc, (kb | ks) != 0);
                }
            }
        }
        long long t1 = clock64();
        if (PAIR) umma_commit_pair(&bar, 1); else umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    if (threadIdx.x < 32) { if (PAIR) tmem_dealloc_pair<512>(slot); else tmem_dealloc<512>(slot); }
}

template <bool PAIR, int N>
void run(long long *d)
{
    const int reps = 16;
    auto kern = bench<PAIR, N>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(PAIR ? 2 : 1);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaMemset(d, 0, 64);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, d, reps);
    long long h[2];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaGetLastError();
    const double n_mma = 32.0 * reps;
    const double macs = (PAIR ? 256.0 : 128.0) * N * 16 * n_mma / (PAIR ? 2 : 1);   // per SM
    printf("cta_group::%d M=%3d N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%.0f MAC/cyc/SM) %s\n", PAIR ? 2 : 1, PAIR ? 256 : 128, N,
           h[0] / n_mma, h[1] / n_mma, macs / h[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    long long *d;
    cudaMalloc(&d, 64);
    run<false, 128>(d); run<false, 256>(d);
    run<true, 128>(d); run<true, 256>(d);
    run<false, 128>(d); run<true, 256>(d);
    return 0;
}
