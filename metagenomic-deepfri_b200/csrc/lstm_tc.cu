// Persistent batched LSTM recurrence on tcgen05 (engine 1).
//
// One *group* of H/16 CTAs covers all hidden units; CTA s owns units [16s, 16s+16), i.e. 64 gate rows
// (gate-major: row = gate*16 + u, ONNX gate order i,o,f,c).  Its slice of the recurrent matrix R is
// split into hi + lo fp16 terms and stays RESIDENT IN SHARED MEMORY for the whole kernel as the MMA
// A operand (2 x 64 x H x 2 B = 128 KiB at H = 512).  Proteins run in sub-batches of 32 (the MMA N
// dimension).  Every group drives two independent *slots*, each with its own list of sub-batches, its
// own issuer warp, epilogue warps, TMEM accumulator and exchange buffers, so one slot's tensor-core
// work overlaps the other slot's h exchange.  Per step and slot:
//     issuer  : wait until every CTA of the group has published h_{t-1} (acquire counter in global
//               memory) -> one bulk-TMA copy of h_{t-1} [32 x H] fp16 into shared memory (MMA B operand)
//               -> 2 x H/16 MMAs (M=64, N=32, K=16): gates^T[64 x 32] = (R_hi + R_lo) . h_{t-1}^T in TMEM
//     epilogue: TMEM -> registers -> smem staging -> cell update (c in registers, fp32) -> h_t (fp16) to the
//               group's exchange buffer and to the [Tp x H] operand image -> release counter
// Input pre-activations (x_t W^T + b) come from a resident table (layer 1: one-hot input = row
// gather) or from the fp32 [Tp x 4H] output of the input GEMM (layers >= 2), both in
// [unit][gate] order so a cell reads one float4.
#include <algorithm>
#include <stdlib.h>
#include <vector>

#include "gemm_tc.cuh"
#include "lstm_tc.cuh"

namespace mdf {
namespace tc {

constexpr int LSTM_BS = 32;           // proteins per sub-batch (MMA N)
constexpr int LSTM_TC_THREADS = 384;  // warps 0,1: issuers of slot 0,1; warps 4-7 / 8-11: epilogue of slot 0 / 1
constexpr int GS_STRIDE = 33;

struct LstmTcArgs {
    int H, n, n_lists, cpg, n_sub;
    int alternate;            // 1: Rimg terms are (R_a, R_b) used on alternating steps; 0: (hi, lo) both every step
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
    long long *trace;         // optional [trace_items][8] clock64 stamps of CTA 0 / slot 0 (MDF_LSTM_TRACE=1)
    int trace_items;
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
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x)
{
    // tanh(x) = 1 - 2 / (exp(2x) + 1): exact at both limits, absolute error ~1e-7
    return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f);
}

// Sub-batch j = proteins order[32j .. 32j+31] (length-descending).  List v = group*2 + slot owns the
// sub-batches j = v, v + n_lists, ...
struct SubBatch {
    int sb, Lmax;
};
__device__ __forceinline__ SubBatch next_sub_batch(int &cursor, const LstmTcArgs &a)
{
    SubBatch r{-1, 0};
    if (cursor < a.n_sub) {
        const int p0 = a.order[cursor * LSTM_BS];                // longest protein of the sub-batch
        r.Lmax = (int)(a.seq_off[p0 + 1] - a.seq_off[p0]);
        if (r.Lmax > 0) r.sb = cursor;                           // sorted descending: zero length = nothing left
        cursor += a.n_lists;
    }
    return r;
}

__global__ void __launch_bounds__(LSTM_TC_THREADS, 1) lstm_tc_kernel(LstmTcArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar_tfull[2];      // MMAs of a slot finished -> its epilogue may read TMEM
    __shared__ uint64_t bar_hfull[2];      // bulk copy of a slot's h operand landed
    __shared__ uint32_t tmem_slot;
    __shared__ int sub_len[2][LSTM_BS];
    __shared__ long long sub_row[2][LSTM_BS];

    const int H = a.H, KS = H / 16;
    const int g = blockIdx.x / a.cpg, s = blockIdx.x % a.cpg;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t r_term_bytes = 64u * H * 2u, hb_bytes = (uint32_t)LSTM_BS * H * 2u;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sR = smem;                                   // [2][64 x H] fp16
    uint8_t *sH = sR + 2 * r_term_bytes;                  // [2 slots][32 x H] fp16
    float *gs_all = reinterpret_cast<float *>(sH + 2 * hb_bytes);    // [2 slots][64][33] gate staging
    float *tabS = gs_all + 2 * 64 * GS_STRIDE;            // [26][16][4]

    // ---- one-time setup: resident weights, barriers, TMEM
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(a.Rimg) + (size_t)s * 2 * r_term_bytes);
        uint4 *dst = reinterpret_cast<uint4 *>(sR);
        for (uint32_t e = tid; e < 2 * r_term_bytes / 16; e += LSTM_TC_THREADS) dst[e] = src[e];
        if (a.tab)
            for (int e = tid; e < 26 * 64; e += LSTM_TC_THREADS) tabS[e] = a.tab[(size_t)s * 26 * 64 + e];
    }
    if (tid == 0) {
        for (int k = 0; k < 2; ++k) { mbar_init(&bar_tfull[k], 1); mbar_init(&bar_hfull[k], 1); }
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<64>(&tmem_slot);
    fence_proxy_async_smem();             // resident R was written through the generic proxy
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp < 2) {
        // =========================================================== issuer of slot k = warp (one lane)
        const int k = warp;
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(64, LSTM_BS);
            unsigned *flag = a.flags + g * 2 + k;
            const __half *hb = a.hbuf + (size_t)(g * 2 + k) * 2 * LSTM_BS * H;
            const uint32_t sb_addr = smem_u32(sH + k * hb_bytes);
            const uint32_t d = tmem_base + (uint32_t)(k * LSTM_BS);
            const uint64_t a_hi = umma_smem_desc(smem_u32(sR), 1024, 128);
            const uint64_t a_lo = umma_smem_desc(smem_u32(sR) + r_term_bytes, 1024, 128);
            const uint64_t b0 = umma_smem_desc(sb_addr, 512, 128);
            int cursor = g * 2 + k;
            unsigned done = 0;          // steps completed on this slot (all sub-batches)
            uint32_t rounds = 0;        // MMA rounds issued -> mbarrier parity
            int item = 0;
            for (SubBatch sbt = next_sub_batch(cursor, a); sbt.sb >= 0; sbt = next_sub_batch(cursor, a)) {
                for (int t = 0; t < sbt.Lmax; ++t, ++done, ++item) {
                    const bool tr = a.trace && blockIdx.x == 0 && k == 0 && item < a.trace_items;
                    if (tr) a.trace[item * 8 + 0] = clock64();
                    // every CTA of the group has finished all earlier steps of this slot
                    const unsigned target = done * (unsigned)a.cpg;
                    while (ld_acquire(flag) < target) { }
                    if (tr) a.trace[item * 8 + 1] = clock64();
                    if (t == 0) continue;                        // h_{-1} = 0: the epilogue uses the input term alone
                    fence_proxy_async_all();                     // peers' generic-proxy stores -> this async-proxy read
                    mbar_arrive_expect_tx(&bar_hfull[k], hb_bytes);
                    bulk_g2s(sH + k * hb_bytes, hb + (size_t)((t - 1) & 1) * LSTM_BS * H, hb_bytes, &bar_hfull[k]);
                    mbar_wait(&bar_hfull[k], rounds & 1);
                    if (tr) a.trace[item * 8 + 2] = clock64();
                    tcgen05_fence_after();
                    if (a.alternate) {
                        // time-dithered weights: R_a = fp16(R) on odd steps, R_b = fp16(2R - R_a) on even steps;
                        // R_a + R_b = 2R to ~2^-22, so the rounding error of the weights changes sign every
                        // step instead of accumulating coherently along the sequence (half the MMAs of hi+lo)
                        const uint64_t aw = (t & 1) ? a_hi : a_lo;
#pragma unroll 8
                        for (int ks = 0; ks < KS; ++ks)
                            umma_f16(d, aw + (uint64_t)(ks * 128), b0 + (uint64_t)(ks * 64), idesc, ks != 0);
                    } else {
#pragma unroll 8
                        for (int ks = 0; ks < KS; ++ks)
                            umma_f16(d, a_hi + (uint64_t)(ks * 128), b0 + (uint64_t)(ks * 64), idesc, ks != 0);
#pragma unroll 8
                        for (int ks = 0; ks < KS; ++ks)
                            umma_f16(d, a_lo + (uint64_t)(ks * 128), b0 + (uint64_t)(ks * 64), idesc, true);
                    }
                    umma_commit(&bar_tfull[k]);
                    ++rounds;
                    if (tr) a.trace[item * 8 + 3] = clock64();
                }
            }
        }
    } else if (warp >= 4) {
        // =========================================================== epilogue of slot k (128 threads)
        const int k = (warp - 4) >> 2;
        const int et = tid - 128 - k * 128;
        const int ew = warp & 3;                 // TMEM sub-partition = gate index
        const int pn = et & 31, q = et >> 5;     // cell ownership: protein pn, units 4q .. 4q+3
        const int bar_id = 2 + k;
        float *gs = gs_all + k * 64 * GS_STRIDE;
        unsigned *flag = a.flags + g * 2 + k;
        uint8_t *hb = reinterpret_cast<uint8_t *>(a.hbuf + (size_t)(g * 2 + k) * 2 * LSTM_BS * H);
        const int H4 = 4 * H;
        const int unit = s * 16 + 4 * q;
        int cursor = g * 2 + k;
        unsigned done = 0;
        uint32_t rounds = 0;
        int item = 0;
        for (SubBatch sbt = next_sub_batch(cursor, a); sbt.sb >= 0; sbt = next_sub_batch(cursor, a)) {
            // The exchange buffers of this slot are reused: every CTA of the group must have finished the
            // previous sub-batch (its last loads) before step 0 publishes into them.
            if (et == 0) {
                const unsigned target = done * (unsigned)a.cpg;
                while (ld_acquire(flag) < target) { }
            }
            named_bar_sync(bar_id, 128);
            if (et < LSTM_BS) {
                const int j = sbt.sb * LSTM_BS + et;
                int len = 0;
                long long row = 0;
                if (j < a.n) {
                    const int pid = a.order[j];
                    len = (int)(a.seq_off[pid + 1] - a.seq_off[pid]);
                    row = a.seg_off[pid];
                }
                sub_len[k][et] = len; sub_row[k][et] = row;
            }
            named_bar_sync(bar_id, 128);
            const int my_len = sub_len[k][pn];
            const long long row0 = sub_row[k][pn];
            float cstate[4] = {0.f, 0.f, 0.f, 0.f};
            for (int t = 0; t < sbt.Lmax; ++t, ++done, ++item) {
                const bool tr = a.trace && blockIdx.x == 0 && k == 0 && et == 0 && item < a.trace_items;
                if (tr) a.trace[item * 8 + 4] = clock64();
                const bool active = t < my_len;
                const long long row = row0 + t;
                // prefetch the input pre-activations of my 4 cells ([unit][gate] order -> one float4 per cell)
                float4 pre[4];
                if (active) {
                    if (a.pre) {
                        const float4 *p = reinterpret_cast<const float4 *>(a.pre + (size_t)row * H4 + (size_t)unit * 4);
#pragma unroll
                        for (int j = 0; j < 4; ++j) pre[j] = __ldg(p + j);
                    } else {
                        const float4 *p = reinterpret_cast<const float4 *>(tabS + ((int)a.idx_pad[row] * 16 + 4 * q) * 4);
#pragma unroll
                        for (int j = 0; j < 4; ++j) pre[j] = p[j];
                    }
                }
                if (t > 0) {
                    mbar_wait(&bar_tfull[k], rounds & 1);
                    ++rounds;
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
                named_bar_sync(bar_id, 128);
                if (tr) a.trace[item * 8 + 5] = clock64();
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
                        const float c = sigmoid_f(gf) * cstate[j] + sigmoid_f(gi) * tanh_f(gc);
                        cstate[j] = c;
                        hv[j] = sigmoid_f(go) * tanh_f(c);
                    }
                    uint2 pk;
                    pk.x = pack_half2(hv[0], hv[1]);
                    pk.y = pack_half2(hv[2], hv[3]);
                    // exchange buffer (smem image layout of the B operand): units `unit .. unit+3` of protein pn
                    *reinterpret_cast<uint2 *>(hb + (size_t)(t & 1) * hb_bytes + ((unit >> 3) * 4 + (pn >> 3)) * 128 +
                                               (pn & 7) * 16 + (unit & 7) * 2) = pk;
                    // operand image for the downstream GEMMs
                    *reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(a.Himg) + image_offset_bytes(row, unit, H / TILE_K)) = pk;
                }
                named_bar_sync(bar_id, 128);     // all h_t stores of this CTA issued, gate staging free again
                if (tr) a.trace[item * 8 + 6] = clock64();
                if (et == 0) red_release(flag, 1u);   // release: cumulative over the stores ordered by the barrier
                if (tr) a.trace[item * 8 + 7] = clock64();
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem_base);
}

size_t lstm_tc_smem_bytes(int H)
{
    return (size_t)2 * 64 * H * 2 + (size_t)2 * LSTM_BS * H * 2 + (size_t)2 * 64 * GS_STRIDE * 4 + 26 * 64 * 4 + 1024;
}

int lstm_tc_groups(const mdf_ctx *ctx, int H) { return std::max(1, ctx->sm_count / (H / 16)); }

size_t lstm_tc_scratch_bytes(const mdf_ctx *ctx, int H)
{
    return (size_t)lstm_tc_groups(ctx, H) * 4 * LSTM_BS * H * 2 + 4096;
}

int launch_lstm_tc(mdf_ctx *ctx, int H, int n, const __half *Rimg, const float *tab, const float *pre,
                   const uint8_t *idx_pad, const int *order, const int64_t *seq_off, const int64_t *seg_off,
                   __half *Himg, void *scratch, int alternate)
{
    if (n <= 0) return MDF_OK;
    LstmTcArgs a;
    a.H = H; a.n = n; a.alternate = alternate;
    a.cpg = H / 16;
    a.n_sub = cdiv(n, LSTM_BS);
    const int n_groups = std::min(lstm_tc_groups(ctx, H), cdiv(a.n_sub, 2));
    a.n_lists = n_groups * 2;
    a.Rimg = Rimg; a.tab = tab; a.pre = pre; a.idx_pad = idx_pad; a.order = order;
    a.seq_off = seq_off; a.seg_off = seg_off; a.Himg = Himg;
    a.flags = reinterpret_cast<unsigned *>(scratch);
    a.hbuf = reinterpret_cast<__half *>(reinterpret_cast<uint8_t *>(scratch) + 4096);
    MDF_CUDA(cudaMemsetAsync(scratch, 0, 4096, ctx->stream));
    a.trace = nullptr; a.trace_items = 0;
    const bool want_trace = getenv("MDF_LSTM_TRACE") != nullptr;
    if (want_trace) {
        a.trace_items = 4096;
        MDF_CUDA(cudaMalloc((void **)&a.trace, (size_t)a.trace_items * 8 * sizeof(long long)));
        MDF_CUDA(cudaMemsetAsync(a.trace, 0, (size_t)a.trace_items * 8 * sizeof(long long), ctx->stream));
    }
    const size_t smem = lstm_tc_smem_bytes(H);
    MDF_CUDA(cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {&a};
    MDF_CUDA(cudaLaunchCooperativeKernel((void *)lstm_tc_kernel, dim3(n_groups * a.cpg), dim3(LSTM_TC_THREADS),
                                         args, smem, ctx->stream));
    ctx->launches++;
    if (want_trace) {
        std::vector<long long> h((size_t)a.trace_items * 8);
        MDF_CUDA(cudaStreamSynchronize(ctx->stream));
        MDF_CUDA(cudaMemcpy(h.data(), a.trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(a.trace);
        double sum[8] = {0}; int cnt = 0; long long first = 0, last = 0; int nrec = 0;
        for (int i = 0; i < a.trace_items; ++i) { if (h[i * 8 + 7]) { if (!first) first = h[i * 8]; last = h[i * 8 + 7]; ++nrec; } }
        for (int i = 64; i < a.trace_items; ++i) {
            const long long *t = &h[(size_t)i * 8];
            if (!t[7] || !t[3] || !t[2]) continue;
            sum[0] += t[1] - t[0];   // issuer: flag wait
            sum[1] += t[2] - t[1];   // issuer: bulk copy of h
            sum[2] += t[3] - t[2];   // issuer: MMA issue
            sum[3] += t[5] - t[4];   // epilogue: wait MMA + tcgen05.ld + staging
            sum[4] += t[6] - t[5];   // epilogue: cell update + stores + bar
            sum[5] += t[7] - t[6];   // epilogue: release
            sum[6] += t[5] - t[3];   // MMA issue done -> epilogue has the gates
            sum[7] += t[1] - t[7 - 8];   // previous release (this CTA) -> flag seen complete
            ++cnt;
        }
        if (cnt)
            fprintf(stderr, "[lstm trace] slot-0 items %d (avg over %d): cycles/item %.0f | issuer: flagwait %.0f copy %.0f issue %.0f | "
                            "epi: waitmma+ld %.0f cell %.0f release %.0f | issue->gates %.0f | own release->flag complete %.0f\n",
                    nrec, cnt, nrec ? (double)(last - first) / nrec : 0.0, sum[0] / cnt, sum[1] / cnt, sum[2] / cnt, sum[3] / cnt,
                    sum[4] / cnt, sum[5] / cnt, sum[6] / cnt, sum[7] / cnt);
    }
    return MDF_OK;
}

}  // namespace tc
}  // namespace mdf
