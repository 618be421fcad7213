// Persistent batched LSTM recurrence on tcgen05 (engine 1).
//
// One *group* of H/16 CTAs covers all hidden units; CTA s owns units [16s, 16s+16), i.e. 64 gate rows
// (gate-major: row = gate*16 + u, ONNX gate order i,o,f,c).  Its slice of the recurrent matrix R is
// split into hi + lo fp16 terms and stays RESIDENT IN SHARED MEMORY for the whole kernel as the MMA
// A operand (2 x 64 x H x 2 B = 128 KiB at H = 512).  Proteins run in sub-batches of 32 (the MMA N
// dimension); each group keeps two sub-batches in flight ("slots") so that one slot's tensor-core
// work overlaps the other's h exchange.  Per step and slot:
//     gates^T[64 x 32] (TMEM, fp32) = R_hi . h_{t-1}^T + R_lo . h_{t-1}^T         (2 x H/16 MMAs, M=64 N=32 K=16)
//     epilogue: TMEM -> registers -> smem staging -> cell update (c in registers, fp32) ->
//               h_t (fp16) to the group's exchange buffer + to the [Tp x H] operand image
//     exchange: release/acquire counter per (group, slot) in global memory; the next step's loaders
//               copy the full h_t [32 x H] back into shared memory as the MMA B operand.
// Input pre-activations (x_t W^T + b) come from a resident table (layer 1: one-hot input = row
// gather) or from the fp32 [Tp x 4H] output of the input GEMM (layers >= 2), both in
// [unit][gate] order so a cell reads one float4.
#include <algorithm>

#include "gemm_tc.cuh"
#include "lstm_tc.cuh"

namespace mdf {
namespace tc {

constexpr int LSTM_BS = 32;          // proteins per sub-batch (MMA N)
constexpr int LSTM_TC_THREADS = 256; // warps 0-3: h loaders + MMA issue, warps 4-7: epilogue
constexpr int GS_STRIDE = 33;

struct LstmTcArgs {
    int H, n, n_groups, cpg, n_sub;
    const __half *Rimg;       // [cpg][2][64 x H] operand images (hi, lo)
    const float *tab;         // [cpg][26][16][4] layer-1 pre-activation table, or nullptr
    const float *pre;         // [Tp][4H] fp32, [unit][gate] order, or nullptr
    const uint8_t *idx_pad;   // [Tp]
    const int *order;         // [n] protein ids, length-descending
    const int64_t *seq_off;   // [n+1] packed offsets (lengths)
    const int64_t *seg_off;   // [n+1] padded row offsets
    __half *Himg;             // [Tp x H] output operand image
    __half *hbuf;             // [n_groups][2 slots][2 parities][32 x H] exchange buffers (smem image layout)
    unsigned *flags;          // [n_groups][2]
};

struct SlotState {
    int sb;      // sub-batch index, -1 = idle
    int t;       // next step
    int Lmax;    // steps of this sub-batch
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(unsigned *p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x)
{
    // tanh(x) = 1 - 2 / (exp(2x) + 1); exact at the limits, ~1e-7 relative elsewhere
    const float e = __expf(2.0f * x);
    return 1.0f - 2.0f / (e + 1.0f);
}

// sub-batch j = proteins order[32j .. 32j+31]; group g owns j = g, g + n_groups, ...
__device__ __forceinline__ void slot_load(SlotState &s, int &next_sb, const LstmTcArgs &a)
{
    if (next_sb < a.n_sub) {
        s.sb = next_sb;
        s.t = 0;
        const int p0 = a.order[next_sb * LSTM_BS];               // longest protein of the sub-batch
        s.Lmax = (int)(a.seq_off[p0 + 1] - a.seq_off[p0]);
        next_sb += a.n_groups;
        if (s.Lmax == 0) s.sb = -1;                              // sorted descending: nothing left to do
    } else {
        s.sb = -1;
    }
}

__global__ void __launch_bounds__(LSTM_TC_THREADS, 1) lstm_tc_kernel(LstmTcArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar_tfull[2];      // MMAs of a slot finished -> epilogue may read TMEM
    __shared__ uint64_t bar_hfree[2];      // MMAs of a slot finished -> loaders may overwrite its h operand
    __shared__ uint32_t tmem_slot;
    __shared__ int sub_pid[2][LSTM_BS], sub_len[2][LSTM_BS];
    __shared__ long long sub_row[2][LSTM_BS];

    const int H = a.H, KS = H / 16;
    const int g = blockIdx.x / a.cpg, s = blockIdx.x % a.cpg;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t r_term_bytes = 64u * H * 2u, hb_bytes = (uint32_t)LSTM_BS * H * 2u;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sR = smem;                                   // [2][64 x H] fp16
    uint8_t *sH = sR + 2 * r_term_bytes;                  // [2 slots][32 x H] fp16
    float *gs = reinterpret_cast<float *>(sH + 2 * hb_bytes);        // [64][33] gate staging
    float *tabS = gs + 64 * GS_STRIDE;                    // [26][16][4]

    // ---- one-time setup: resident weights, barriers, TMEM
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(a.Rimg) + (size_t)s * 2 * r_term_bytes);
        uint4 *dst = reinterpret_cast<uint4 *>(sR);
        for (uint32_t e = tid; e < 2 * r_term_bytes / 16; e += LSTM_TC_THREADS) dst[e] = src[e];
        if (a.tab)
            for (int e = tid; e < 26 * 64; e += LSTM_TC_THREADS) tabS[e] = a.tab[(size_t)s * 26 * 64 + e];
    }
    if (tid == 0) {
        for (int k = 0; k < 2; ++k) { mbar_init(&bar_tfull[k], 1); mbar_init(&bar_hfree[k], 1); }
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<64>(&tmem_slot);
    fence_proxy_async_smem();             // resident R was written through the generic proxy
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;

    SlotState slot[2];
    int next_sb = g;
    slot_load(slot[0], next_sb, a);
    slot_load(slot[1], next_sb, a);
    unsigned done[2] = {0u, 0u};           // steps completed on each slot (all sub-batches)
    uint32_t mma_count[2] = {0u, 0u};      // MMA rounds issued on each slot -> mbarrier parities
    unsigned *flag = a.flags + g * 2;
    __half *hbuf_g = a.hbuf + (size_t)g * 4 * LSTM_BS * H;

    if (warp < 4) {
        // =========================================================== loaders + MMA issue (128 threads)
        constexpr uint32_t idesc = umma_idesc_f16(64, LSTM_BS);
        while (slot[0].sb >= 0 || slot[1].sb >= 0) {
            for (int k = 0; k < 2; ++k) {
                SlotState &sl = slot[k];
                if (sl.sb < 0) continue;
                // every CTA of the group has finished all earlier steps on this slot
                if (tid == 0) {
                    const unsigned target = done[k] * (unsigned)a.cpg;
                    while (ld_acquire(flag + k) < target) { }
                }
                named_bar_sync(1, 128);
                if (sl.t > 0) {
                    if (mma_count[k] > 0) mbar_wait(&bar_hfree[k], (mma_count[k] - 1) & 1);   // operand buffer free
                    const uint4 *src = reinterpret_cast<const uint4 *>(hbuf_g + (size_t)(k * 2 + ((sl.t - 1) & 1)) * LSTM_BS * H);
                    uint4 *dst = reinterpret_cast<uint4 *>(sH + k * hb_bytes);
                    const int nchunk = (int)(hb_bytes / 16);
#pragma unroll 4
                    for (int e = tid; e < nchunk; e += 128) dst[e] = __ldcg(src + e);
                    fence_proxy_async_smem();
                    named_bar_sync(1, 128);
                    if (tid == 0) {
                        tcgen05_fence_after();
                        const uint32_t d = tmem_base + (uint32_t)(k * LSTM_BS);
                        const uint32_t sa = smem_u32(sR), sb = smem_u32(sH + k * hb_bytes);
                        for (int term = 0; term < 2; ++term)
                            for (int ks = 0; ks < KS; ++ks) {
                                const uint64_t ad = umma_smem_desc(sa + term * r_term_bytes + ks * 2 * 1024, 1024, 128);
                                const uint64_t bd = umma_smem_desc(sb + ks * 2 * 512, 512, 128);
                                umma_f16(d, ad, bd, idesc, (term | ks) != 0);
                            }
                        umma_commit(&bar_hfree[k]);
                        umma_commit(&bar_tfull[k]);
                    }
                    ++mma_count[k];
                }
                ++done[k];
                if (++sl.t == sl.Lmax) slot_load(sl, next_sb, a);
            }
        }
    } else {
        // =========================================================== epilogue (128 threads)
        const int et = tid - 128;
        const int ew = warp & 3;                 // TMEM sub-partition = gate index
        const int pn = et & 31, q = et >> 5;     // cell ownership: protein pn, units 4q .. 4q+3
        float cstate[2][4] = {};
        int cur_sb[2] = {-1, -1};
        const int H4 = 4 * H;
        while (slot[0].sb >= 0 || slot[1].sb >= 0) {
            for (int k = 0; k < 2; ++k) {
                SlotState &sl = slot[k];
                if (sl.sb < 0) continue;
                if (cur_sb[k] != sl.sb) {        // new sub-batch on this slot: protein table + zero state
                    // the exchange buffers of this slot are reused: every CTA of the group must have finished
                    // the previous sub-batch (its last loads) before step 0 publishes into them
                    if (et == 0) {
                        const unsigned target = done[k] * (unsigned)a.cpg;
                        while (ld_acquire(flag + k) < target) { }
                    }
                    named_bar_sync(2, 128);
                    if (et < LSTM_BS) {
                        const int j = sl.sb * LSTM_BS + et;
                        int pid = -1, len = 0;
                        long long row = 0;
                        if (j < a.n) {
                            pid = a.order[j];
                            len = (int)(a.seq_off[pid + 1] - a.seq_off[pid]);
                            row = a.seg_off[pid];
                        }
                        sub_pid[k][et] = pid; sub_len[k][et] = len; sub_row[k][et] = row;
                    }
                    named_bar_sync(2, 128);
                    cur_sb[k] = sl.sb;
#pragma unroll
                    for (int j = 0; j < 4; ++j) cstate[k][j] = 0.0f;
                }
                const int t = sl.t;
                const bool active = t < sub_len[k][pn];
                const long long row = sub_row[k][pn] + t;
                // prefetch the input pre-activations of my 4 cells ([unit][gate] order -> one float4 per cell)
                float4 pre[4];
                if (active) {
                    if (a.pre) {
                        const float4 *p = reinterpret_cast<const float4 *>(a.pre + (size_t)row * H4 + (size_t)(s * 16 + 4 * q) * 4);
#pragma unroll
                        for (int j = 0; j < 4; ++j) pre[j] = __ldg(p + j);
                    } else {
                        const float4 *p = reinterpret_cast<const float4 *>(tabS + ((int)a.idx_pad[row] * 16 + 4 * q) * 4);
#pragma unroll
                        for (int j = 0; j < 4; ++j) pre[j] = p[j];
                    }
                }
                if (t > 0) {
                    mbar_wait(&bar_tfull[k], mma_count[k] & 1);
                    ++mma_count[k];
                    tcgen05_fence_after();
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(k * LSTM_BS), r);
                    tmem_ld_wait();
                    if (lane < 16) {
                        float *dst = gs + (ew * 16 + lane) * GS_STRIDE;
#pragma unroll
                        for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(r[j]);
                    }
                    tcgen05_fence_before();
                }
                named_bar_sync(2, 128);
                if (active) {
                    float hv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int u = 4 * q + j;
                        float gi = pre[j].x, go = pre[j].y, gf = pre[j].z, gc = pre[j].w;
                        if (t > 0) {
                            gi += gs[(0 * 16 + u) * GS_STRIDE + pn];
                            go += gs[(1 * 16 + u) * GS_STRIDE + pn];
                            gf += gs[(2 * 16 + u) * GS_STRIDE + pn];
                            gc += gs[(3 * 16 + u) * GS_STRIDE + pn];
                        }
                        const float c = sigmoid_f(gf) * cstate[k][j] + sigmoid_f(gi) * tanh_f(gc);
                        cstate[k][j] = c;
                        hv[j] = sigmoid_f(go) * tanh_f(c);
                    }
                    uint2 pk;
                    pk.x = pack_half2(hv[0], hv[1]);
                    pk.y = pack_half2(hv[2], hv[3]);
                    // exchange buffer (smem image layout of the B operand): unit 16s + 4q .. +3 of protein pn
                    const int unit = s * 16 + 4 * q;
                    uint8_t *hb = reinterpret_cast<uint8_t *>(hbuf_g + (size_t)(k * 2 + (t & 1)) * LSTM_BS * H);
                    *reinterpret_cast<uint2 *>(hb + ((unit >> 3) * 4 + (pn >> 3)) * 128 + (pn & 7) * 16 + (unit & 7) * 2) = pk;
                    // operand image for the downstream GEMMs
                    *reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(a.Himg) + image_offset_bytes(row, unit, H / TILE_K)) = pk;
                }
                named_bar_sync(2, 128);          // all h_t stores of this CTA issued, gate staging free again
                if (et == 0) {
                    __threadfence();
                    red_release(flag + k, 1u);
                }
                ++done[k];
                if (++sl.t == sl.Lmax) slot_load(sl, next_sb, a);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem_base);
}

size_t lstm_tc_smem_bytes(int H)
{
    return (size_t)2 * 64 * H * 2 + (size_t)2 * LSTM_BS * H * 2 + (size_t)64 * GS_STRIDE * 4 + 26 * 64 * 4 + 1024;
}

int lstm_tc_groups(const mdf_ctx *ctx, int H) { return std::max(1, ctx->sm_count / (H / 16)); }

size_t lstm_tc_scratch_bytes(const mdf_ctx *ctx, int H)
{
    return (size_t)lstm_tc_groups(ctx, H) * 4 * LSTM_BS * H * 2 + 4096;
}

int launch_lstm_tc(mdf_ctx *ctx, int H, int n, const __half *Rimg, const float *tab, const float *pre,
                   const uint8_t *idx_pad, const int *order, const int64_t *seq_off, const int64_t *seg_off,
                   __half *Himg, void *scratch)
{
    if (n <= 0) return MDF_OK;
    LstmTcArgs a;
    a.H = H; a.n = n;
    a.cpg = H / 16;
    a.n_sub = cdiv(n, LSTM_BS);
    a.n_groups = std::min(lstm_tc_groups(ctx, H), a.n_sub);
    a.Rimg = Rimg; a.tab = tab; a.pre = pre; a.idx_pad = idx_pad; a.order = order;
    a.seq_off = seq_off; a.seg_off = seg_off; a.Himg = Himg;
    a.flags = reinterpret_cast<unsigned *>(scratch);
    a.hbuf = reinterpret_cast<__half *>(reinterpret_cast<uint8_t *>(scratch) + 4096);
    MDF_CUDA(cudaMemsetAsync(scratch, 0, 4096, ctx->stream));
    const size_t smem = lstm_tc_smem_bytes(H);
    MDF_CUDA(cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {&a};
    MDF_CUDA(cudaLaunchCooperativeKernel((void *)lstm_tc_kernel, dim3(a.n_groups * a.cpg), dim3(LSTM_TC_THREADS),
                                         args, smem, ctx->stream));
    ctx->launches++;
    return MDF_OK;
}

}  // namespace tc
}  // namespace mdf
