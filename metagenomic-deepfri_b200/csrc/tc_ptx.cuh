// sm_100a primitives used by the tensor-core engine: mbarrier, bulk-TMA copies, TMEM management,
// tcgen05.mma / ld / commit and the UMMA shared-memory + instruction descriptors.
// Inline PTX only; bit layouts follow the public CUTLASS sm100 headers
// (cute/arch/mma_sm100_desc.hpp: SmemDescriptor, InstrDescriptor).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace mdf {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a fully converged warp (CUTLASS cute::elect_one_sync).  The single-thread tcgen05 / TMA instructions take
// their operands from uniform registers: issued under `if (lane == 0)` the compiler cannot prove uniformity and wraps
// every one of them in a per-lane "waterfall" loop (ELECT / BRA.U.ANY, ~200 cycles per MMA measured); issued by the
// elected lane of a warp that runs the surrounding loop converged, they are a single predicated instruction.
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred != 0;
}

// 16-byte store to shared memory by shared-window address (a generic pointer rebuilt from uintptr arithmetic makes the
// compiler emit generic ST.E instead of STS)
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}

// Wait that backs off with nanosleep between polls: the B200 runs these kernels at its power cap, and a warp spinning on
// try_wait burns issue energy for the whole time its producer needs - use for waits that are expected to be long.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity, uint32_t ns)
{
    while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

// ---------------------------------------------------------------- bulk TMA (1-D, no tensor map)
// global -> shared, completion signalled on an mbarrier through complete_tx::bytes
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// multicast form: the bytes land at the same shared-memory offset in every CTA of `cta_mask` (cluster ranks) and complete_tx is
// signalled on the barrier at the same offset in each of them - one L2 read feeds all the CTAs that stream the same operand
__device__ __forceinline__ void bulk_g2s_mc(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint16_t cta_mask)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (tensor core / TMA)
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- TMEM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot)   // whole warp
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr)      // whole warp (the allocating one)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread = its warp's lane)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor, no swizzle (LayoutType::SWIZZLE_NONE), K-major operand stored as
// 8-row x 16-byte core matrices (128 contiguous bytes each):
//   lbo = byte distance between core matrices adjacent along K
//   sbo = byte distance between core matrices adjacent along M/N (next 8 rows)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);            // [0,14)  start address >> 4
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;      // [16,30) leading byte offset >> 4
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;      // [32,46) stride byte offset >> 4
    d |= (uint64_t)1 << 46;                                // [46,48) descriptor version = 1 (sm_100)
    return d;                                              // base_offset 0, lbo_mode 0, layout_type 0
}

// Instruction descriptor for kind::f16: A,B = f16 (K-major), D = f32
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N)
{
    return (1u << 4)                      // c_format = F32
           | (0u << 7) | (0u << 10)       // a_format = b_format = F16
           | (0u << 15) | (0u << 16)      // a_major = b_major = K
           | ((uint32_t)(N >> 3) << 17)   // n_dim
           | ((uint32_t)(M >> 4) << 24);  // m_dim
}

// D[tmem] (+)= A[smem] . B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

// Converged-warp forms: executed by ALL 32 lanes with identical operands, the instruction itself predicated on the
// elected lane inside the asm block - no C++ branch, so the operands stay in uniform registers (no waterfall loop).
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t *bar)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar)) : "memory");
}

// single-CTA MMAs, completion signalled on the barrier at this offset in every CTA of `cta_mask` (stages shared through multicast)
__device__ __forceinline__ void umma_commit_mc_elect(uint64_t *bar, uint16_t cta_mask)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

__device__ __forceinline__ void umma_f16_pair_elect(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair_elect(uint64_t *bar, uint16_t cta_mask)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
// One k-block of a CTA-pair pass issued by a converged warp under ONE elect.sync: [re-arm the operand barrier] + the four
// K = 16 MMAs of a 64-wide k-block + commit -> weight stage free [+ commit -> operand chunk free].  Barrier arguments are
// shared-window addresses (0 = skip).  The per-instruction elect forms above cost ~50 issue cycles per MMA (ELECT, VOTEU,
// R2UR of every descriptor); a single issuing warp then needs longer to issue a k-block than the tensor pipe needs to run it.
__device__ __forceinline__ void umma_f16_pair_kblock_elect(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc_first,
                                                           uint32_t rearm_bar, uint32_t rearm_bytes, uint32_t empty_bar, uint16_t empty_mask,
                                                           uint32_t free_bar, uint16_t free_mask)
{
    asm volatile(
        "{\n\t.reg .pred q, p0, pt, pr, pe, pf;\n\t"
        ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p0, %4, 0;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "setp.ne.b32 pr, %5, 0;\n\t"
        "and.pred pr, pr, q;\n\t"
        "setp.ne.b32 pe, %7, 0;\n\t"
        "and.pred pe, pe, q;\n\t"
        "setp.ne.b32 pf, %9, 0;\n\t"
        "and.pred pf, pf, q;\n\t"
        "add.s64 a1, %1, 256;\n\t"
        "add.s64 a2, %1, 512;\n\t"
        "add.s64 a3, %1, 768;\n\t"
        "add.s64 b1, %2, 256;\n\t"
        "add.s64 b2, %2, 512;\n\t"
        "add.s64 b3, %2, 768;\n\t"
        "@pr mbarrier.arrive.expect_tx.shared::cta.b64 _, [%5], %6;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, pt;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, pt;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, pt;\n\t"
        "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%7], %8;\n\t"
        "@pf tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%9], %10;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc_first), "r"(rearm_bar), "r"(rearm_bytes), "r"(empty_bar), "h"(empty_mask),
          "r"(free_bar), "h"(free_mask) : "memory");
}
// Wait for one or two mbarrier phases in one asm block (both polls are issued before either result is consumed; the label is
// local to the braces, so the block can be inlined any number of times).
__device__ __forceinline__ void mbar_wait2_asm(uint32_t bar_a, uint32_t parity_a, uint32_t bar_b, uint32_t parity_b)
{
    asm volatile(
        "{\n\t.reg .pred p1, p2;\n\t"
        "MDF_WAIT2:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p1, [%0], %1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p2, [%2], %3;\n\t"
        "and.pred p1, p1, p2;\n\t"
        "@!p1 bra MDF_WAIT2;\n\t}"
        ::"r"(bar_a), "r"(parity_a), "r"(bar_b), "r"(parity_b) : "memory");
}
__device__ __forceinline__ void mbar_wait1_asm(uint32_t bar_a, uint32_t parity_a)
{
    asm volatile(
        "{\n\t.reg .pred p1;\n\t"
        "MDF_WAIT1:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p1, [%0], %1;\n\t"
        "@!p1 bra MDF_WAIT1;\n\t}"
        ::"r"(bar_a), "r"(parity_a) : "memory");
}
// keeps a hoisted value in a register: the compiler would otherwise rematerialise shared-window addresses (S2UR + ULEA) at every use
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) { uint32_t r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ uint64_t pin_u64(uint64_t v) { uint64_t r; asm volatile("mov.u64 %0, %1;" : "=l"(r) : "l"(v)); return r; }
// Single-CTA variant for a 128 x 256 tile held as two N = 128 accumulators (the adjacency product): eight MMAs of one 64-wide k-block
// (A . B0 -> d0, A . B1 -> d0 + 128) + commit -> stage free [+ commit -> accumulator full] under one elect.sync.
__device__ __forceinline__ void umma_f16_kblock_2n128_elect(uint32_t tmem_d, uint64_t a_desc, uint64_t b0_desc, uint64_t b1_desc, uint32_t idesc,
                                                            uint32_t acc_first, uint32_t empty_bar, uint32_t full_bar)
{
    asm volatile(
        "{\n\t.reg .pred q, p0, pt, pf;\n\t"
        ".reg .b64 a1, a2, a3, b1, b2, b3, c1, c2, c3;\n\t"
        ".reg .b32 d1;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p0, %5, 0;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "setp.ne.b32 pf, %7, 0;\n\t"
        "and.pred pf, pf, q;\n\t"
        "add.u32 d1, %0, 128;\n\t"
        "add.s64 a1, %1, 256;\n\t"
        "add.s64 a2, %1, 512;\n\t"
        "add.s64 a3, %1, 768;\n\t"
        "add.s64 b1, %2, 256;\n\t"
        "add.s64 b2, %2, 512;\n\t"
        "add.s64 b3, %2, 768;\n\t"
        "add.s64 c1, %3, 256;\n\t"
        "add.s64 c2, %3, 512;\n\t"
        "add.s64 c3, %3, 768;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %4, p0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [d1], %1, %3, %4, p0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %4, pt;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [d1], a1, c1, %4, pt;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %4, pt;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [d1], a2, c2, %4, pt;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %4, pt;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [d1], a3, c3, %4, pt;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
        "@pf tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b0_desc), "l"(b1_desc), "r"(idesc), "r"(acc_first), "r"(empty_bar), "r"(full_bar) : "memory");
}
// the same block without its MMAs (timing experiments)
__device__ __forceinline__ void umma_pair_kblock_nomma_elect(uint32_t rearm_bar, uint32_t rearm_bytes, uint32_t empty_bar, uint16_t empty_mask,
                                                             uint32_t free_bar, uint16_t free_mask)
{
    asm volatile(
        "{\n\t.reg .pred q, pr, pe, pf;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 pr, %0, 0;\n\t"
        "and.pred pr, pr, q;\n\t"
        "setp.ne.b32 pe, %2, 0;\n\t"
        "and.pred pe, pe, q;\n\t"
        "setp.ne.b32 pf, %4, 0;\n\t"
        "and.pred pf, pf, q;\n\t"
        "@pr mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
        "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%2], %3;\n\t"
        "@pf tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%4], %5;\n\t}"
        ::"r"(rearm_bar), "r"(rearm_bytes), "r"(empty_bar), "h"(empty_mask), "r"(free_bar), "h"(free_mask) : "memory");
}
// try_wait on a shared-window address; two barriers polled together (their latencies overlap)
__device__ __forceinline__ bool mbar_try_wait_addr(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait_addr(bar, parity)) { }
}
__device__ __forceinline__ void mbar_wait2_addr(uint32_t bar_a, uint32_t parity_a, uint32_t bar_b, uint32_t parity_b)
{
    bool a = mbar_try_wait_addr(bar_a, parity_a), b = mbar_try_wait_addr(bar_b, parity_b);
    while (!(a & b)) {
        if (!a) a = mbar_try_wait_addr(bar_a, parity_a);
        if (!b) b = mbar_try_wait_addr(bar_b, parity_b);
    }
}
// arrive + expect_tx by the elected lane of a converged warp
__device__ __forceinline__ void mbar_arrive_expect_tx_elect(uint64_t *bar, uint32_t bytes)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2) and clusters
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()     // every thread of every CTA of the cluster
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of this cluster.  Default (.release.cta)
// semantics as in CUTLASS ClusterBarrier::arrive(cta_id): a cluster-scope release would drain this SM's
// outstanding global stores (and invalidate L1) on every call, and what the arrival orders here is
// async-proxy data (TMA writes, tcgen05.ld reads) that is already complete when the local barrier fires.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t cta)
{
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *smem_slot)   // one warp in EACH CTA of the pair, same slot offset
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 from each CTA's smem] . B[N rows: N/2 from each CTA's smem]^T ; leader CTA only
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// all previously issued pair MMAs arrive on the mbarrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

// 16 KiB tile (a [32 x 256] box of a flat [bytes/512][256] u16 tensor map) -> this CTA's shared memory; the
// completion bytes are credited to the barrier at the same offset in the LEADER CTA of the pair (peer bit of the
// shared::cluster address cleared, as CUTLASS SM100_TMA_2SM_LOAD does), which is what lets the leader's MMA
// issuer wait for both halves of an operand without a software relay.  Executed by both CTAs.
__device__ __forceinline__ void tma_tile_g2s_pair(void *dst_smem, const void *tmap, int32_t row512, uint64_t *bar)
{
    const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(0), "r"(row512), "r"(mbar) : "memory");
}

// Multicast form: the tile lands at the same shared-memory offset in every CTA of `cta_mask` (cluster ranks) and its bytes
// are credited to the barrier at this offset in the leader of each destination CTA's pair (CUTLASS
// SM100_TMA_2SM_LOAD_MULTICAST) - one L2 read feeds all the CTAs that need the same operand tile.
__device__ __forceinline__ void tma_tile_g2s_pair_mc(void *dst_smem, const void *tmap, int32_t row512, uint64_t *bar, uint16_t cta_mask)
{
    const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
                 " [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(0), "r"(row512), "r"(mbar), "h"(cta_mask) : "memory");
}

// Same copy issued by the LEADER on behalf of CTA `dst_cta` of its pair: the tile lands at the same offset of that CTA's
// shared memory and the bytes are credited to the leader's own barrier - the peer needs no producer thread and its
// stage is refilled without waiting for a commit to cross the pair.
__device__ __forceinline__ void tma_tile_g2s_pair_to(void *dst_smem, uint32_t dst_cta, const void *tmap, int32_t row512, uint64_t *bar)
{
    uint32_t dst;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(smem_u32(dst_smem)), "r"(dst_cta));
    const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tmap), "r"(0), "r"(row512), "r"(mbar) : "memory");
}

// ---------------------------------------------------------------- fast transcendental primitives (MUFU, flush-to-zero)
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---------------------------------------------------------------- operand tile images
// A [rows x K] fp16 operand lives in HBM as 128-row x 64-k tiles of 16 KiB, each stored exactly as
// the shared-memory image the tensor core reads (no-swizzle K-major core matrices):
//   tile(rt, kb) at ((rt * KB) + kb) * 16384 bytes
//   inside a tile: byte = ((k%64)/8 * 16 + (r%128)/8) * 128 + (r%8) * 16 + (k%8) * 2
// => descriptor LBO = 2048, SBO = 128; one bulk copy moves a whole tile.
constexpr int TILE_ROWS = 128, TILE_K = 64, TILE_BYTES = 16384;
constexpr uint32_t TILE_LBO = 2048, TILE_SBO = 128;

__host__ __device__ __forceinline__ size_t image_offset_bytes(int64_t r, int k, int KB)
{
    const int64_t rt = r >> 7;
    const int kb = k >> 6;
    return ((size_t)(rt * KB + kb)) * TILE_BYTES + (size_t)((((k & 63) >> 3) * 16 + ((r & 127) >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b)
{
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

}  // namespace tc
}  // namespace mdf
