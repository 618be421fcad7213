// Micro-benchmark: issue cost (cycles of the issuing warp) of the building blocks of a tcgen05 issuer loop on a CTA pair:
// elected 4-MMA k-block, tcgen05.commit (multicast), mbarrier.try_wait on a completed phase, with and without 16 busy
// "epilogue" warps on the same SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/issue_bench tools/issue_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../metagenomic-deepfri_b200/csrc/tc_ptx.cuh"
using namespace mdf::tc;

#ifndef NOISE_UNROLL
#define NOISE_UNROLL 64
#endif
// NOISE_UNROLL sets the code footprint of the noise warps' loop body (3 instructions = 48 bytes per step)
__device__ __forceinline__ float busy(float x, int n)
{
#pragma unroll 1
    for (int i = 0; i < n * 64 / NOISE_UNROLL; ++i) {
#pragma unroll
        for (int j = 0; j < NOISE_UNROLL; ++j) x = fmaf(x, 1.0001f, ex2_ftz(x * 0.001f));
    }
    return x;
}

// mode: 0 = 4-MMA block, 1 = 4-MMA block + 1 commit, 2 = + 2 commits, 3 = no MMA, 1 commit, 4 = no MMA, 2 commits, 5 = try_wait only,
//       6 = elect block with nothing enabled, 7 = 4-MMA block + 2 commits + 2 try_waits (the full k-block)
__global__ void bench(long long *out, int reps, int mode, int noise, float *sink)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar[8];
    __shared__ uint32_t slot;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    const bool leader = cluster_ctarank() == 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc_pair<512>(&slot);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    if (warp == 0) {
        if (leader) {
            const uint32_t idesc = umma_idesc_f16(256, 256);
            const uint64_t a0 = umma_smem_desc(smem_u32(smem), TILE_LBO, TILE_SBO), b0 = umma_smem_desc(smem_u32(smem) + 128 * 1024, TILE_LBO, TILE_SBO);
            const uint32_t bar0 = smem_u32(&bar[0]);
            // one completed phase on bar[4], bar[5] for the try_wait tests
            if (lane == 0) { mbar_arrive(&bar[4]); mbar_arrive(&bar[5]); }
            __syncwarp();
            long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                for (int kb = 0; kb < 8; ++kb) {
                    const uint64_t ad = a0 + (uint64_t)(kb * 1024), bd = b0 + (uint64_t)((kb & 3) * 1024);
                    if (mode == 5 || mode == 7) mbar_wait2_addr(bar0 + 32, 0, bar0 + 40, 0);
                    if (mode <= 2 || mode == 7)
                        umma_f16_pair_kblock_elect(slot, ad, bd, idesc, kb != 0, 0u, 0u, mode >= 1 ? bar0 + 8u * (kb & 3) : 0u, (uint16_t)3,
                                                   (mode == 2 || mode == 7) ? bar0 + 8u * (kb & 3) : 0u, (uint16_t)3);
                    else if (mode == 3 || mode == 4 || mode == 6)
                        umma_pair_kblock_nomma_elect(0u, 0u, mode != 6 ? bar0 + 8u * (kb & 3) : 0u, (uint16_t)3, mode == 4 ? bar0 + 8u * (kb & 3) : 0u, (uint16_t)3);
                }
            }
            long long t1 = clock64();
            umma_commit_pair_elect(&bar[6], 1);
            mbar_wait(&bar[6], 0);
            long long t2 = clock64();
            if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    } else if (noise) {
        float x = busy((float)threadIdx.x, reps * noise);
        if (x == 123.456f) sink[0] = x;
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_pair<512>(slot);
}

int main()
{
    long long *d; float *sink;
    cudaMalloc(&d, 64); cudaMalloc(&sink, 64);
    const int reps = 64;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char *names[] = {"4 MMA", "4 MMA + 1 commit", "4 MMA + 2 commits", "1 commit", "2 commits", "2 try_wait (done)", "empty elect block", "2 try_wait + 4 MMA + 2 commits"};
    for (int noise = 0; noise <= 8; noise += 8)
        for (int mode = 0; mode < 8; ++mode) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2);
            cfg.blockDim = dim3(17 * 32);
            cfg.dynamicSmemBytes = 200 * 1024;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cudaMemset(d, 0, 64);
            cudaError_t e = cudaLaunchKernelEx(&cfg, bench, d, reps, mode, noise, sink);
            long long h[2];
            cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaGetLastError();
            printf("noise %d  %-34s issue %.1f cyc/k-block, complete %.1f cyc/k-block %s\n", noise, names[mode], h[0] / (8.0 * reps), h[1] / (8.0 * reps),
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
