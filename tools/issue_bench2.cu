// Micro-benchmark 2: issue overhead of different ways to let one thread of a warp issue tcgen05 instructions (CTA pair).
// MMAs use N = 16 (8 tensor cycles each) so that the issue path, not the tensor pipe, is what is measured.
#include <cstdio>
#include <cuda_runtime.h>
#include "../metagenomic-deepfri_b200/csrc/tc_ptx.cuh"
using namespace mdf::tc;

__device__ __forceinline__ uint32_t elect_one()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void mma_plain(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_plain(uint32_t bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void wait4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, uint32_t par)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p0, p1, p2, p3;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p0, [%1], %5;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p1, [%2], %5;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p2, [%3], %5;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p3, [%4], %5;\n\t"
                     "and.pred p0, p0, p1;\n\tand.pred p2, p2, p3;\n\tand.pred p0, p0, p2;\n\t"
                     "selp.u32 %0, 1, 0, p0;\n\t}" : "=r"(ok) : "r"(b0), "r"(b1), "r"(b2), "r"(b3), "r"(par) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void wait2(uint32_t b0, uint32_t b1, uint32_t par)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p0, p1;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p0, [%1], %3;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p1, [%2], %3;\n\t"
                     "and.pred p0, p0, p1;\n\t"
                     "selp.u32 %0, 1, 0, p0;\n\t}" : "=r"(ok) : "r"(b0), "r"(b1), "r"(par) : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint32_t test_wait1(uint32_t b0, uint32_t par)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p0;\n\tmbarrier.test_wait.parity.shared::cta.b64 p0, [%1], %2;\n\tselp.u32 %0, 1, 0, p0;\n\t}" : "=r"(ok) : "r"(b0), "r"(par) : "memory");
    return ok;
}

// mode 0: converged warp, one elect asm block per k-block (4 MMAs + 2 commits)          [current fast path]
// mode 1: converged warp, `if (elect_one())` region around plain instructions, per k-block (CUTLASS style, coarser)
// mode 2: single live lane (others exited), plain instructions
// mode 3: mode 0 with two k-blocks per iteration (8 MMAs + 4 commits)
// mode 4: only wait2 (single asm);  5: only wait4 (single asm);  6: only 2x mbar_wait_addr (separate);  7: elect_one() alone; 8: test_wait x2
// mode 9: mode 2 + wait2 per k-block;  10: mode 0 + wait2 per k-block;  11: single live lane, 4 MMAs only; 12: single live lane, 2 commits only
__global__ void bench(long long *out, int reps, int mode)
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
    if (warp == 1 && leader) {
        const uint32_t idesc = umma_idesc_f16(256, 16);
        const uint64_t a0 = umma_smem_desc(smem_u32(smem), TILE_LBO, TILE_SBO), b0 = umma_smem_desc(smem_u32(smem) + 128 * 1024, TILE_LBO, TILE_SBO);
        const uint32_t bar0 = smem_u32(&bar[0]);
        const uint32_t tm = slot;
        if (lane == 0) { for (int i = 4; i < 8; ++i) mbar_arrive(&bar[i]); }
        __syncwarp();
        const bool single = mode == 2 || mode == 9 || mode == 11 || mode == 12;
        if (!single || lane == 0) {
            long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                for (int kb = 0; kb < 8; ++kb) {
                    const uint64_t ad = a0 + (uint64_t)(kb * 1024), bd = b0 + (uint64_t)((kb & 3) * 1024);
                    const uint32_t eb = bar0 + 8u * (kb & 3);
                    if (mode == 0 || mode == 10) {
                        if (mode == 10) wait2(bar0 + 32, bar0 + 40, 0);
                        umma_f16_pair_kblock_elect(tm, ad, bd, idesc, kb != 0, 0u, 0u, eb, (uint16_t)3, eb, (uint16_t)3);
                    } else if (mode == 1) {
                        if (elect_one()) {
                            mma_plain(tm, ad, bd, idesc, kb != 0); mma_plain(tm, ad + 256, bd + 256, idesc, 1);
                            mma_plain(tm, ad + 512, bd + 512, idesc, 1); mma_plain(tm, ad + 768, bd + 768, idesc, 1);
                            commit_plain(eb, 3); commit_plain(eb, 3);
                        }
                        __syncwarp();
                    } else if (single) {
                        if (mode == 9) wait2(bar0 + 32, bar0 + 40, 0);
                        if (mode != 12) {
                            mma_plain(tm, ad, bd, idesc, kb != 0); mma_plain(tm, ad + 256, bd + 256, idesc, 1);
                            mma_plain(tm, ad + 512, bd + 512, idesc, 1); mma_plain(tm, ad + 768, bd + 768, idesc, 1);
                        }
                        if (mode != 11) { commit_plain(eb, 3); commit_plain(eb, 3); }
                    } else if (mode == 3) {
                        if (kb & 1) continue;
                        umma_f16_pair_kblock_elect(tm, ad, bd, idesc, kb != 0, 0u, 0u, eb, (uint16_t)3, eb, (uint16_t)3);
                        umma_f16_pair_kblock_elect(tm, ad + 1024, bd + 1024, idesc, 1, 0u, 0u, eb, (uint16_t)3, eb, (uint16_t)3);
                    } else if (mode == 4) wait2(bar0 + 32, bar0 + 40, 0);
                    else if (mode == 5) wait4(bar0 + 32, bar0 + 40, bar0 + 48, bar0 + 56, 0);
                    else if (mode == 6) { mbar_wait_addr(bar0 + 32, 0); mbar_wait_addr(bar0 + 40, 0); }
                    else if (mode == 7) { if (elect_one()) out[4] = 1; __syncwarp(); }
                    else if (mode == 8) { while (!(test_wait1(bar0 + 32, 0) & test_wait1(bar0 + 40, 0))) { } }
                }
            }
            long long t1 = clock64();
            if (lane == 0) { out[0] = t1 - t0; }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_pair<512>(slot);
}

int main()
{
    long long *d;
    cudaMalloc(&d, 64);
    const int reps = 64;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char *names[] = {"elect block: 4 MMA + 2 commits", "if(elect) region: 4 MMA + 2 commits", "single live lane: 4 MMA + 2 commits", "elect blocks, 2 k-blocks/iter (per k-block)",
                           "wait2 (one asm)", "wait4 (one asm)", "2 x mbar_wait (separate)", "elect_one alone", "2 x test_wait", "single lane: wait2 + 4 MMA + 2 commits",
                           "elect block + wait2", "single lane: 4 MMA", "single lane: 2 commits"};
    for (int mode = 0; mode < 13; ++mode) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2);
        cfg.blockDim = dim3(64);
        cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaMemset(d, 0, 64);
        cudaError_t e = cudaLaunchKernelEx(&cfg, bench, d, reps, mode);
        long long h[2];
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaGetLastError();
        printf("%-46s %.1f cyc/k-block %s\n", names[mode], h[0] / (8.0 * reps), e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
