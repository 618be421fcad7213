// Large-batch LSTM recurrence on tcgen05 with STREAMED weights (engine 1, batches >= ~2k proteins).
//
// The resident-weight kernel (lstm_tc.cu) is limited to 32 proteins per MMA by shared memory and is
// latency-bound; with thousands of proteins in flight it pays to turn the partition around:
//   * a group of H/128 CTAs covers the hidden units; CTA s owns units [128s, 128s+128) = 512 gate rows;
//   * the step operand h_{t-1} of 128 proteins [128 x H] fp16 is RESIDENT in shared memory (128 KiB at
//     H = 512) as the MMA *A* operand (M = 128 proteins), and the CTA's slice of R (512 x H, 512 KiB per
//     term) is streamed from L2 through a bulk-TMA ring as the *B* operand (N = 128 gate rows of one gate).
//     The weight stream does not depend on h, so it is prefetched across the step boundary while the group
//     exchanges h_t;
//   * four 128x128 fp32 accumulators in TMEM, one per gate (i, o, f, c): TMEM lane = protein, column =
//     unit, so an epilogue thread finds all four gates of a cell in its own lane, owns one protein, and
//     reads / writes runs of consecutive units (float4 pre-activation loads, 16-byte h stores straight into
//     the operand images);
//   * per step: 4 x H/16 MMAs of 128x128x16 -> epilogue (tcgen05.ld -> cell update with c in registers ->
//     h_t fp16 to the group's exchange buffer and to the [Tp x H] operand image) -> release counter; the
//     next step's issuer acquires the counter and bulk-copies the full h_t back into shared memory.
// Weights use the same time-dithered fp16 pair (R_a, R_b) as the resident kernel.
#include <algorithm>
#include <stdlib.h>
#include <vector>

#include "gemm_tc.cuh"
#include "lstm_tc.cuh"

namespace mdf {
namespace tc {

constexpr int LS_N = 128;            // proteins per sub-batch (MMA M)
constexpr int LS_UNITS = 128;        // hidden units per CTA
constexpr int LS_THREADS = 320;      // warp 0: weight producer, warp 1: h copy + MMA issue, warps 2-9: epilogue
constexpr int LS_STAGES = 5;

struct LstmStreamArgs {
    int H, n, n_groups, cpg, n_sub;
    const __half *Rimg[2];    // [4H rows (cta, gate, unit) x H] operand images, indexed by step parity
    const float *tab;         // [26][H][4] layer-1 pre-activation table ([unit][gate] order), or nullptr
    const float *pre;         // [Tp][4H] fp32, [unit][gate] order, or nullptr
    const uint8_t *idx_pad;   // [Tp]
    const int *order;         // [n] protein ids, length-descending
    const int64_t *seq_off;   // [n+1]
    const int64_t *seg_off;   // [n+1] padded row offsets
    __half *Himg;             // [Tp x H] output operand image
    __half *hbuf;             // [n_groups][2 parities][128 x H] exchange buffers (tile images, rows = proteins)
    unsigned *flags;          // [n_groups]
    long long *trace;         // optional clock64 stamps of CTA 0 (MDF_LSTM_TRACE=1)
    int trace_items;
};

__device__ __forceinline__ unsigned ls_ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ls_red_release(unsigned *p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void ls_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// LSTM cell with shared denominators: 5 exp + 2 reciprocals instead of 5 + 5.
//   c' = sigmoid(f) c + sigmoid(i) tanh(g) = [c B C + (C - 2) A] / (A B C),  A = 1+e^-f, B = 1+e^-i, C = e^2g + 1
//   h  = sigmoid(o) tanh(c')             = (E - 2) / (D E),                  D = 1+e^-o, E = e^2c' + 1
// Pre-activations are clamped so that the products stay finite in fp32 (sigmoid/tanh change by < 1e-10).
__device__ __forceinline__ void ls_cell(float xi, float xo, float xf, float xg, float c_prev, float &c_out, float &h_out)
{
    const float A = 1.0f + __expf(-fminf(fmaxf(xf, -25.0f), 25.0f));
    const float B = 1.0f + __expf(-fminf(fmaxf(xi, -25.0f), 25.0f));
    const float C = 1.0f + __expf(2.0f * fminf(fmaxf(xg, -12.0f), 12.0f));
    const float c = __fdividef(c_prev * (B * C) + (C - 2.0f) * A, A * (B * C));
    const float D = 1.0f + __expf(-fminf(fmaxf(xo, -25.0f), 25.0f));
    const float E = 1.0f + __expf(2.0f * fminf(fmaxf(c, -12.0f), 12.0f));
    c_out = c;
    h_out = __fdividef(E - 2.0f, D * E);
}

struct LsSub {
    int sb, Lmax;
};
__device__ __forceinline__ LsSub ls_next(int &cursor, const LstmStreamArgs &a)
{
    LsSub r{-1, 0};
    if (cursor < a.n_sub) {
        const int p0 = a.order[cursor * LS_N];
        r.Lmax = (int)(a.seq_off[p0 + 1] - a.seq_off[p0]);
        if (r.Lmax > 0) r.sb = cursor;
        cursor += a.n_groups;
    }
    return r;
}

__global__ void __launch_bounds__(LS_THREADS, 1) lstm_stream_kernel(LstmStreamArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar_full[LS_STAGES], bar_empty[LS_STAGES], bar_hfull, bar_tfull;
    __shared__ uint32_t tmem_slot;

    const int H = a.H, KB = H / TILE_K;                   // k-blocks (8 at H = 512)
    const int g = blockIdx.x / a.cpg, s = blockIdx.x % a.cpg;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sH = smem;                                   // [KB tiles][16 KiB]  h operand, rows = proteins
    uint8_t *sW = sH + (size_t)KB * TILE_BYTES;           // [LS_STAGES][16 KiB] weight ring
    const uint32_t h_bytes = (uint32_t)KB * TILE_BYTES;

    if (tid == 0) {
        for (int i = 0; i < LS_STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        mbar_init(&bar_hfull, 1);
        mbar_init(&bar_tfull, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(&tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;
    unsigned *flag = a.flags + g;

    if (warp == 0) {
        // =========================================================== weight producer
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            int cursor = g;
            for (LsSub sbt = ls_next(cursor, a); sbt.sb >= 0; sbt = ls_next(cursor, a)) {
                for (int t = 1; t < sbt.Lmax; ++t) {
                    const uint8_t *src = reinterpret_cast<const uint8_t *>(a.Rimg[t & 1]) + (size_t)(s * 4) * KB * TILE_BYTES;
                    for (int kb = 0; kb < KB; ++kb)
                        for (int gate = 0; gate < 4; ++gate) {
                            mbar_wait(&bar_empty[st], ph ^ 1);
                            mbar_arrive_expect_tx(&bar_full[st], TILE_BYTES);
                            bulk_g2s(sW + (size_t)st * TILE_BYTES, src + ((size_t)gate * KB + kb) * TILE_BYTES, TILE_BYTES, &bar_full[st]);
                            if (++st == LS_STAGES) { st = 0; ph ^= 1; }
                        }
                }
            }
        }
    } else if (warp == 1) {
        // =========================================================== h copy + MMA issue
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, 128);
            int st = 0; uint32_t ph = 0;
            unsigned done = 0;
            uint32_t rounds = 0;
            int cursor = g;
            const uint32_t sh_addr = smem_u32(sH);
            const uint8_t *hb = reinterpret_cast<const uint8_t *>(a.hbuf) + (size_t)g * 2 * h_bytes;
            for (LsSub sbt = ls_next(cursor, a); sbt.sb >= 0; sbt = ls_next(cursor, a)) {
                for (int t = 0; t < sbt.Lmax; ++t, ++done) {
                    const bool tr = a.trace && blockIdx.x == 0 && (int)done < a.trace_items;
                    if (tr) a.trace[done * 8 + 0] = clock64();
                    const unsigned target = done * (unsigned)a.cpg;
                    while (ls_ld_acquire(flag) < target) { __nanosleep(64); }
                    if (tr) a.trace[done * 8 + 1] = clock64();
                    if (t == 0) continue;
                    asm volatile("fence.proxy.async;" ::: "memory");     // peers' generic stores -> this async-proxy read
                    mbar_arrive_expect_tx(&bar_hfull, h_bytes);
                    const uint8_t *src = hb + (size_t)((t - 1) & 1) * h_bytes;
                    for (int kb = 0; kb < KB; ++kb)
                        bulk_g2s(sH + (size_t)kb * TILE_BYTES, src + (size_t)kb * TILE_BYTES, TILE_BYTES, &bar_hfull);
                    mbar_wait(&bar_hfull, rounds & 1);
                    if (tr) a.trace[done * 8 + 2] = clock64();
                    tcgen05_fence_after();
                    for (int kb = 0; kb < KB; ++kb) {
                        const uint64_t hd = umma_smem_desc(sh_addr + kb * TILE_BYTES, TILE_LBO, TILE_SBO);   // A: proteins x 64 k
                        for (int gate = 0; gate < 4; ++gate) {
                            mbar_wait(&bar_full[st], ph);
                            tcgen05_fence_after();
                            const uint64_t wd = umma_smem_desc(smem_u32(sW + (size_t)st * TILE_BYTES), TILE_LBO, TILE_SBO);   // B: gate rows x 64 k
#pragma unroll
                            for (int ks = 0; ks < TILE_K / 16; ++ks)
                                umma_f16(tmem_base + (uint32_t)(gate * 128), hd + (uint64_t)(ks * 256), wd + (uint64_t)(ks * 256),
                                         idesc, (kb | ks) != 0);
                            umma_commit(&bar_empty[st]);
                            if (++st == LS_STAGES) { st = 0; ph ^= 1; }
                        }
                    }
                    umma_commit(&bar_tfull);
                    ++rounds;
                    if (tr) a.trace[done * 8 + 3] = clock64();
                }
            }
        }
    } else {
        // =========================================================== epilogue: thread = one protein x 64 units
        const int et = tid - 64;
        const int lb = (warp & 3) * 32;
        const int p = lb + lane;                            // protein inside the sub-batch = TMEM lane
        const int half = (warp - 2) >> 2;                   // units [64*half, 64*half + 64) of this CTA
        const int unit0 = s * LS_UNITS + half * 64;         // first global hidden unit of this thread
        // byte offset of (row 0, k = unit0) inside a tile image: unit0 is a multiple of 64 -> start of k-block unit0/64
        const size_t kblk = (size_t)(unit0 >> 6) * TILE_BYTES;
        uint8_t *hbw = reinterpret_cast<uint8_t *>(a.hbuf) + (size_t)g * 2 * h_bytes + kblk + (size_t)(p >> 3) * 128 + (size_t)(p & 7) * 16;
        const float4 *pre4 = reinterpret_cast<const float4 *>(a.pre) + unit0;
        const float4 *tab4 = reinterpret_cast<const float4 *>(a.tab) + unit0;
        float cst[64];
        unsigned done = 0;
        uint32_t rounds = 0;
        int cursor = g;
        for (LsSub sbt = ls_next(cursor, a); sbt.sb >= 0; sbt = ls_next(cursor, a)) {
            // The exchange buffers are reused: every CTA of the group must have finished the previous sub-batch
            // (its last loads) before step 0 publishes into them.
            if (et == 0) {
                const unsigned target = done * (unsigned)a.cpg;
                while (ls_ld_acquire(flag) < target) { __nanosleep(64); }
            }
            ls_bar_sync(1, 256);
            int len = 0;
            long long row0 = 0;
            {
                const int j = sbt.sb * LS_N + p;
                if (j < a.n) {
                    const int pid = a.order[j];
                    len = (int)(a.seq_off[pid + 1] - a.seq_off[pid]);
                    row0 = a.seg_off[pid];
                }
            }
#pragma unroll
            for (int j = 0; j < 64; ++j) cst[j] = 0.0f;
            for (int t = 0; t < sbt.Lmax; ++t, ++done) {
                const bool tr = a.trace && blockIdx.x == 0 && et == 0 && (int)done < a.trace_items;
                if (tr) a.trace[done * 8 + 4] = clock64();
                const bool active = t < len;
                const long long row = row0 + t;
                const float4 *src = nullptr;                // 64 consecutive float4: the pre-activations of my cells
                if (active) src = a.pre ? pre4 + (size_t)row * H : tab4 + (size_t)a.idx_pad[row] * H;
                // the inputs do not depend on the recurrence: pull my 1 KiB into L2 while the MMAs of this step run
                if (active && a.pre) {
#pragma unroll
                    for (int j = 0; j < 64; j += 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + j));
                }
                if (t > 0) {
                    mbar_wait(&bar_tfull, rounds & 1);
                    ++rounds;
                    tcgen05_fence_after();
                }
                if (tr) a.trace[done * 8 + 5] = clock64();
                uint8_t *hdst = hbw + (size_t)(t & 1) * h_bytes;
                uint8_t *idst = reinterpret_cast<uint8_t *>(a.Himg) + (size_t)(row >> 7) * KB * TILE_BYTES + kblk +
                                (size_t)((((int)row & 127) >> 3) * 128 + ((int)row & 7) * 16);
                float4 pre_cur[8], pre_nxt[8];
                if (active) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) pre_cur[j] = __ldg(src + j);
                }
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 8) {
                    if (active && c0 + 8 < 64) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) pre_nxt[j] = __ldg(src + c0 + 8 + j);
                    }
                    uint32_t gi[8], go[8], gf[8], gc[8];
                    if (t > 0) {                              // warp-uniform: tcgen05.ld is .sync.aligned
                        const uint32_t tb = tmem_base + ((uint32_t)lb << 16) + (uint32_t)(half * 64 + c0);
                        tmem_ld_32x32b_x8(tb + 0 * 128, gi);
                        tmem_ld_32x32b_x8(tb + 1 * 128, go);
                        tmem_ld_32x32b_x8(tb + 2 * 128, gf);
                        tmem_ld_32x32b_x8(tb + 3 * 128, gc);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) gi[j] = go[j] = gf[j] = gc[j] = 0u;
                    }
                    if (active) {
                        float hv[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 pre = pre_cur[j];
                            float c;
                            ls_cell(__uint_as_float(gi[j]) + pre.x, __uint_as_float(go[j]) + pre.y, __uint_as_float(gf[j]) + pre.z,
                                    __uint_as_float(gc[j]) + pre.w, cst[c0 + j], c, hv[j]);
                            cst[c0 + j] = c;
                        }
                        uint4 pk;
                        pk.x = pack_half2(hv[0], hv[1]); pk.y = pack_half2(hv[2], hv[3]);
                        pk.z = pack_half2(hv[4], hv[5]); pk.w = pack_half2(hv[6], hv[7]);
                        // units unit0+c0 .. +7 of my protein: one 16-byte chunk of the exchange tile and of the H image
                        *reinterpret_cast<uint4 *>(hdst + (c0 >> 3) * 2048) = pk;
                        *reinterpret_cast<uint4 *>(idst + (c0 >> 3) * 2048) = pk;
#pragma unroll
                        for (int j = 0; j < 8; ++j) pre_cur[j] = pre_nxt[j];
                    }
                }
                tcgen05_fence_before();
                ls_bar_sync(1, 256);
                if (tr) a.trace[done * 8 + 6] = clock64();
                if (et == 0) ls_red_release(flag, 1u);
                if (tr) a.trace[done * 8 + 7] = clock64();
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

bool lstm_stream_supported(int H) { return H % LS_UNITS == 0 && (size_t)(H / TILE_K + LS_STAGES) * TILE_BYTES + 2048 <= 227 * 1024; }

size_t lstm_stream_scratch_bytes(const mdf_ctx *ctx, int H)
{
    const int groups = std::max(1, ctx->sm_count / std::max(1, H / LS_UNITS));
    return (size_t)groups * 2 * LS_N * H * 2 + 4096;
}

int launch_lstm_stream(mdf_ctx *ctx, int H, int n, const __half *Ra, const __half *Rb, const float *tab, const float *pre,
                       const uint8_t *idx_pad, const int *order, const int64_t *seq_off, const int64_t *seg_off,
                       __half *Himg, void *scratch)
{
    if (n <= 0) return MDF_OK;
    LstmStreamArgs a;
    a.H = H; a.n = n;
    a.cpg = H / LS_UNITS;
    a.n_sub = cdiv(n, LS_N);
    a.n_groups = std::min(std::max(1, ctx->sm_count / a.cpg), a.n_sub);
    a.Rimg[1] = Ra; a.Rimg[0] = Rb;            // odd steps use R_a, even steps R_b (same convention as lstm_tc.cu)
    a.tab = tab; a.pre = pre; a.idx_pad = idx_pad; a.order = order;
    a.seq_off = seq_off; a.seg_off = seg_off; a.Himg = Himg;
    a.flags = reinterpret_cast<unsigned *>(scratch);
    a.hbuf = reinterpret_cast<__half *>(reinterpret_cast<uint8_t *>(scratch) + 4096);
    MDF_CUDA(cudaMemsetAsync(scratch, 0, 4096, ctx->stream));
    a.trace = nullptr; a.trace_items = 0;
    const bool want_trace = getenv("MDF_LSTM_TRACE") != nullptr;
    if (want_trace) {
        a.trace_items = 2048;
        MDF_CUDA(cudaMalloc((void **)&a.trace, (size_t)a.trace_items * 8 * sizeof(long long)));
        MDF_CUDA(cudaMemsetAsync(a.trace, 0, (size_t)a.trace_items * 8 * sizeof(long long), ctx->stream));
    }
    const size_t smem = (size_t)(H / TILE_K + LS_STAGES) * TILE_BYTES + 1024;
    MDF_CUDA(cudaFuncSetAttribute(lstm_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {&a};
    MDF_CUDA(cudaLaunchCooperativeKernel((void *)lstm_stream_kernel, dim3(a.n_groups * a.cpg), dim3(LS_THREADS), args, smem,
                                         ctx->stream));
    ctx->launches++;
    if (want_trace) {
        std::vector<long long> h((size_t)a.trace_items * 8);
        MDF_CUDA(cudaStreamSynchronize(ctx->stream));
        MDF_CUDA(cudaMemcpy(h.data(), a.trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(a.trace);
        double sum[8] = {0}; int cnt = 0;
        for (int i = 32; i < a.trace_items; ++i) {
            const long long *t = &h[(size_t)i * 8];
            if (!t[7] || !t[3] || !t[2] || !h[(size_t)(i - 1) * 8 + 7]) continue;
            sum[0] += t[1] - t[0];   // issuer: flag wait
            sum[1] += t[2] - t[1];   // issuer: h copy
            sum[2] += t[3] - t[2];   // issuer: MMA issue (paced by the weight stream)
            sum[3] += t[5] - t[3];   // last MMA issued -> epilogue sees the accumulators
            sum[4] += t[6] - t[5];   // epilogue: cell update + stores
            sum[5] += t[7] - t[6];   // epilogue: release
            sum[6] += t[7] - h[(size_t)(i - 1) * 8 + 7];   // full step
            ++cnt;
        }
        if (cnt)
            fprintf(stderr, "[lstm stream trace] avg over %d steps: step %.0f cyc | flagwait %.0f hcopy %.0f mma-issue %.0f mma-drain %.0f "
                            "epilogue %.0f release %.0f\n", cnt, sum[6] / cnt, sum[0] / cnt, sum[1] / cnt, sum[2] / cnt, sum[3] / cnt,
                    sum[4] / cnt, sum[5] / cnt);
    }
    return MDF_OK;
}

}  // namespace tc
}  // namespace mdf
