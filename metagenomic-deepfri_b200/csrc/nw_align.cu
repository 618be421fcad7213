// Batched global alignment (Needleman-Wunsch with affine gaps) on the GPU: SURVEY.md §8f row 4.
//
// Reference: mDeepFRI/alignment.py:163-221 hands every (query, best MMseqs2 hit) pair to PyOpal (Opal's SIMD NW, gap open 10 /
// extend 1, VTML80) inside a multiprocessing.Pool (alignment.py:314) and turns the returned M/X/I/D string into the gapped
// strings (insert_gaps, alignment.py:38-62) that the alignment-transfer kernel consumes.  Recurrences, gap model and the
// traceback order are stated in include/mdf_b200.h and below; the parity tests compare scores AND alignment strings bit for bit
// with the CPU restatement kept beside the tests.
//
// Kernel layout (integer work, bound by instruction issue and shared-memory reads, no tensor cores):
//   * one warp per pair; the target columns of a panel are dealt to the 32 lanes in strips of W consecutive columns
//     (W = 4 / 8 / 16 / 32 chosen from the target length; targets longer than 32 W columns take several panels);
//   * the warp walks the query rows as a systolic wavefront: at step s lane l fills row s - l of its strip, with H / E of the
//     row above in registers and the three values it needs from its left neighbour (H and F of the cell to the left, H of the
//     diagonal cell) arriving by warp shuffle; the rightmost strip's column is kept per row for the next panel;
//   * substitution scores come from a per-warp TARGET PROFILE in shared memory (for every residue r and strip column k the
//     score S[r][t_k]): a row needs one 4..32-byte shared-memory read per lane instead of W table lookups;
//   * every cell leaves a 4-bit code (source of H: diagonal / E / F; "E was extended"; "F was extended") - W codes of a lane
//     are one 16..128-bit store; a second kernel walks the codes back from (Lq, Lt), one thread per pair.
#include <algorithm>
#include <vector>

#include "mdf_common.cuh"

namespace mdf {

namespace {

constexpr int NW_NEG = -1000000000;
constexpr int NW_WARPS = 4;                 // warps (= pairs) per block

struct NwPair {
    int64_t q_off, t_off;                   // into the packed residue codes
    int lq, lt;
    int64_t dir_off;                        // uint32 words: panel-major, then row, lane, word
    int64_t bnd_off;                        // int32: 2 * (lq + 1) boundary values (H, F of the panel's last column)
    int64_t ops_off;                        // output bytes
    int out;                                // index in the caller's order
};

template <int W>
__global__ void __launch_bounds__(NW_WARPS * 32) nw_fill_kernel(int n, const NwPair *__restrict__ pairs, const uint8_t *__restrict__ codes,
                                                                const int8_t *__restrict__ matrix /* [32][32] */, int A, int open, int ext,
                                                                uint32_t *__restrict__ dir, int32_t *__restrict__ bnd,
                                                                int32_t *__restrict__ scores)
{
    constexpr int WL = W >= 8 ? W / 8 : 1;             // uint32 words of direction codes per lane and row
    extern __shared__ __align__(16) int8_t smem[];
    int8_t *mat = smem;                                // [32][32]
    int8_t *prof = smem + 1024 + (threadIdx.x >> 5) * (A * 32 * W);       // this warp: [r < A][lane][W]
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) mat[i] = matrix[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int pi = blockIdx.x * NW_WARPS + (threadIdx.x >> 5);
    if (pi >= n) return;
    const NwPair pr = pairs[pi];
    const uint8_t *q = codes + pr.q_off, *t = codes + pr.t_off;
    const int lq = pr.lq, lt = pr.lt;
    if (lq == 0 || lt == 0) {
        if (lane == 0) scores[pr.out] = (lq + lt) ? -(open + (lq + lt - 1) * ext) : 0;
        return;
    }
    int32_t *bH = bnd + pr.bnd_off, *bF = bH + (lq + 1);
    const int panels = (lt + 32 * W - 1) / (32 * W);
    for (int panel = 0; panel < panels; ++panel) {
        const int c0 = panel * 32 * W + lane * W;      // columns c0 + 1 .. c0 + W (1-based) belong to this lane
        // ---- target profile of the strip
        __syncwarp();
        for (int k = 0; k < W; ++k) {
            const int j = c0 + k;                      // 0-based target index
            const int tc = j < lt ? t[j] : 0;
            for (int r = 0; r < A; ++r) prof[(r * 32 + lane) * W + k] = mat[r * 32 + tc];
        }
        __syncwarp();
        // ---- row 0
        int H[W], E[W];
#pragma unroll
        for (int k = 0; k < W; ++k) { H[k] = -(open + (c0 + k) * ext); E[k] = NW_NEG; }
        int h_last_cur = H[W - 1], h_last_prev = H[W - 1], f_last_cur = NW_NEG;
        uint32_t *dpan = dir + pr.dir_off + (int64_t)panel * lq * 32 * WL;
        const bool lane_live = c0 < lt;
        for (int s = 0; s < lq + 31; ++s) {
            const int i = s - lane;                    // 0-based query row
            int hl = __shfl_up_sync(0xffffffffu, h_last_cur, 1);
            int fl = __shfl_up_sync(0xffffffffu, f_last_cur, 1);
            int hd = __shfl_up_sync(0xffffffffu, h_last_prev, 1);
            const bool active = i >= 0 && i < lq;
            if (lane == 0 && active) {
                if (panel == 0) {
                    hl = -(open + i * ext);            // H[i + 1][0]
                    fl = NW_NEG;
                    hd = i == 0 ? 0 : -(open + (i - 1) * ext);
                } else {
                    hl = bH[i + 1]; fl = bF[i + 1]; hd = bH[i];
                }
            }
            if (active && lane_live) {
                const int qc = q[i];
                const int8_t *prow = prof + (qc * 32 + lane) * W;
                uint32_t scw[W / 4];                    // W signed bytes of the profile row, kept in registers
                if constexpr (W == 4) { scw[0] = *reinterpret_cast<const uint32_t *>(prow); }
                else if constexpr (W == 8) { const uint2 v = *reinterpret_cast<const uint2 *>(prow); scw[0] = v.x; scw[1] = v.y; }
                else {
#pragma unroll
                    for (int v = 0; v < W / 16; ++v) {
                        const uint4 x = reinterpret_cast<const uint4 *>(prow)[v];
                        scw[4 * v] = x.x; scw[4 * v + 1] = x.y; scw[4 * v + 2] = x.z; scw[4 * v + 3] = x.w;
                    }
                }
                uint32_t words[WL];
#pragma unroll
                for (int w = 0; w < WL; ++w) words[w] = 0u;
                int h_left = hl, f_left = fl, h_diag = hd;
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const int e1 = H[k] - open, e2 = E[k] - ext;
                    const int f1 = h_left - open, f2 = f_left - ext;
                    const int e = max(e1, e2), f = max(f1, f2);
                    int h = h_diag + (int)(int8_t)(scw[k >> 2] >> (8 * (k & 3)));
                    uint32_t code = 0u;
                    if (e > h) { h = e; code = 1u; }
                    if (f > h) { h = f; code = 2u; }
                    code |= (e2 > e1 ? 4u : 0u) | (f2 > f1 ? 8u : 0u);
                    words[k >> 3] |= code << (4 * (k & 7));
                    h_diag = H[k];
                    H[k] = h; E[k] = e;
                    h_left = h; f_left = f;
                }
                h_last_prev = h_last_cur;
                h_last_cur = h_left; f_last_cur = f_left;
                uint32_t *dst = dpan + ((int64_t)i * 32 + lane) * WL;
                if constexpr (WL == 1) dst[0] = words[0];
                else if constexpr (WL == 2) *reinterpret_cast<uint2 *>(dst) = make_uint2(words[0], words[1]);
                else *reinterpret_cast<uint4 *>(dst) = make_uint4(words[0], words[1], words[2], words[3]);
                if (lane == 31 && panel + 1 < panels) { bH[i + 1] = h_left; bF[i + 1] = f_left; }
            }
        }
        // H[Lq][Lt]: the strip that holds column Lt still has the last row in registers (picked outside the row loop so that H
        // is never indexed dynamically inside it)
        if (lt > c0 && lt <= c0 + W) {
            int hv = 0;
#pragma unroll
            for (int k = 0; k < W; ++k) hv = (k == lt - 1 - c0) ? H[k] : hv;
            scores[pr.out] = hv;
        }
        if (panel + 1 < panels && lane == 31) bH[0] = -(open + (c0 + W - 1) * ext);      // H[0][last column of the panel]
        __syncwarp();
    }
}

template <int W>
__global__ void nw_trace_kernel(int n, const NwPair *__restrict__ pairs, const uint8_t *__restrict__ codes, const uint32_t *__restrict__ dir,
                                char *__restrict__ ops, int *__restrict__ ops_len)
{
    constexpr int WL = W >= 8 ? W / 8 : 1;
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= n) return;
    const NwPair pr = pairs[pi];
    const uint8_t *q = codes + pr.q_off, *t = codes + pr.t_off;
    char *out = ops + pr.ops_off;
    int i = pr.lq, j = pr.lt, cnt = 0, state = 0;
    const uint32_t *d = dir + pr.dir_off;
    while (i > 0 || j > 0) {
        if (i == 0) { out[cnt++] = 'I'; --j; continue; }
        if (j == 0) { out[cnt++] = 'D'; --i; continue; }
        const int col = j - 1, panel = col / (32 * W), lane = (col / W) & 31, k = col % W;
        const uint32_t word = d[(((int64_t)panel * pr.lq + (i - 1)) * 32 + lane) * WL + (k >> 3)];
        const uint32_t code = (word >> (4 * (k & 7))) & 15u;
        if (state == 0) {
            const uint32_t src = code & 3u;
            if (src == 0u) { out[cnt++] = q[i - 1] == t[j - 1] ? 'M' : 'X'; --i; --j; }
            else state = (int)src;
        } else if (state == 1) {
            out[cnt++] = 'D'; --i;
            state = (code & 4u) ? 1 : 0;
        } else {
            out[cnt++] = 'I'; --j;
            state = (code & 8u) ? 2 : 0;
        }
    }
    for (int a = 0, b = cnt - 1; a < b; ++a, --b) { const char c = out[a]; out[a] = out[b]; out[b] = c; }
    ops_len[pr.out] = cnt;
}

template <int W>
int launch_class(mdf_ctx *ctx, int n, const NwPair *d_pairs, const uint8_t *d_codes, const int8_t *d_matrix, int A, int open, int ext,
                 uint32_t *d_dir, int32_t *d_bnd, int32_t *d_scores, char *d_ops, int *d_len, bool full)
{
    if (n <= 0) return MDF_OK;
    const size_t smem = 1024 + (size_t)NW_WARPS * A * 32 * W;
    auto fill = nw_fill_kernel<W>;
    MDF_CUDA(cudaFuncSetAttribute(fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fill<<<cdiv(n, NW_WARPS), NW_WARPS * 32, smem, ctx->stream>>>(n, d_pairs, d_codes, d_matrix, A, open, ext, d_dir, d_bnd, d_scores);
    MDF_LAUNCH_CHECK(ctx);
    if (full) {
        nw_trace_kernel<W><<<cdiv(n, 64), 64, 0, ctx->stream>>>(n, d_pairs, d_codes, d_dir, d_ops, d_len);
        MDF_LAUNCH_CHECK(ctx);
    }
    return MDF_OK;
}

}  // namespace
}  // namespace mdf

using namespace mdf;

// n (query, target) pairs -> optimal global alignment scores and, with `ops`, the alignment strings over M / X / I / D.
//   query[p] / target[p]: q_len[p] / t_len[p] residues (letters of `alphabet`, anything else -> MDF_EINVAL);
//   matrix: A x A int8 row-major substitution scores in the order of `alphabet` (A <= 32);
//   gap of length k costs gap_open + (k - 1) * gap_extend;
//   ops: flat output, pair p at byte ops_off[p] with room for q_len[p] + t_len[p] columns; ops_len[p] = columns written.
//   ops == NULL: scores only (best_hit_database, alignment.py:163-194).
extern "C" int mdf_nw_align(mdf_ctx *ctx, int n, const char *const *query, const int *q_len, const char *const *target, const int *t_len,
                            const int8_t *matrix, const char *alphabet, int A, int gap_open, int gap_extend, int32_t *scores, char *ops,
                            const int64_t *ops_off, int *ops_len)
{
    MDF_REQUIRE(ctx && n >= 0 && scores && matrix && alphabet && A > 0 && A <= 32, "mdf_nw_align: bad arguments");
    MDF_REQUIRE(n == 0 || (query && q_len && target && t_len), "mdf_nw_align: sequences missing");
    MDF_REQUIRE(!ops || (ops_off && ops_len), "mdf_nw_align: ops_off / ops_len missing");
    MDF_REQUIRE(gap_open >= 0 && gap_extend >= 0 && gap_open < (1 << 20) && gap_extend < (1 << 20), "mdf_nw_align: gap penalties out of range");
    if (n == 0) return MDF_OK;
    MDF_CUDA(cudaSetDevice(ctx->device));
    uint8_t lut[256];
    memset(lut, 255, sizeof lut);
    for (int a = 0; a < A; ++a) lut[(unsigned char)alphabet[a]] = (uint8_t)a;
    // ---- pack + encode on the host, classify by target length
    int64_t total = 0;
    for (int p = 0; p < n; ++p) {
        MDF_REQUIRE(q_len[p] >= 0 && t_len[p] >= 0 && q_len[p] < (1 << 20) && t_len[p] < (1 << 20), "mdf_nw_align: pair %d has an invalid length", p);
        MDF_REQUIRE((query[p] || !q_len[p]) && (target[p] || !t_len[p]), "mdf_nw_align: pair %d has a NULL sequence", p);
        total += q_len[p] + t_len[p];
    }
    std::vector<uint8_t> h_codes((size_t)total + 16);
    std::vector<NwPair> cls[4];
    int64_t at = 0, dir_words = 0, bnd_ints = 0;
    const bool full = ops != nullptr;
    for (int p = 0; p < n; ++p) {
        NwPair pr;
        pr.lq = q_len[p]; pr.lt = t_len[p]; pr.out = p;
        pr.q_off = at;
        for (int i = 0; i < pr.lq; ++i) {
            const uint8_t c = lut[(unsigned char)query[p][i]];
            MDF_REQUIRE(c != 255, "mdf_nw_align: query %d holds '%c', which is not in the scoring matrix alphabet", p, query[p][i]);
            h_codes[(size_t)at++] = c;
        }
        pr.t_off = at;
        for (int i = 0; i < pr.lt; ++i) {
            const uint8_t c = lut[(unsigned char)target[p][i]];
            MDF_REQUIRE(c != 255, "mdf_nw_align: target %d holds '%c', which is not in the scoring matrix alphabet", p, target[p][i]);
            h_codes[(size_t)at++] = c;
        }
        const int k = pr.lt <= 128 ? 0 : pr.lt <= 256 ? 1 : pr.lt <= 512 ? 2 : 3;
        const int W = 4 << k, WL = W >= 8 ? W / 8 : 1;
        const int panels = std::max(1, (pr.lt + 32 * W - 1) / (32 * W));
        pr.dir_off = dir_words;
        // score-only calls still write the codes (one code path); their storage is recycled pair after pair is NOT possible
        // with concurrent warps, so it is sized for real
        dir_words += (int64_t)panels * pr.lq * 32 * WL;
        pr.bnd_off = bnd_ints;
        bnd_ints += 2 * ((int64_t)pr.lq + 1);
        pr.ops_off = full ? ops_off[p] : 0;
        cls[k].push_back(pr);
    }
    int64_t ops_bytes = 0;
    if (full)
        for (int p = 0; p < n; ++p) ops_bytes = std::max(ops_bytes, ops_off[p] + q_len[p] + t_len[p]);
    // longest first inside a class: the tail of the launch is made of short pairs
    for (auto &c : cls) std::stable_sort(c.begin(), c.end(), [](const NwPair &a, const NwPair &b) { return (int64_t)a.lq * a.lt > (int64_t)b.lq * b.lt; });
    std::vector<NwPair> all;
    int cls_at[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < 4; ++k) { all.insert(all.end(), cls[k].begin(), cls[k].end()); cls_at[k + 1] = (int)all.size(); }
    ArenaScope scope(ctx);
    MDF_TRY(ctx->reserve((size_t)total + 16 + all.size() * sizeof(NwPair) + 1024 + (size_t)dir_words * 4 + (size_t)bnd_ints * 4 + (size_t)n * 8 +
                         (size_t)ops_bytes + (1 << 16)));
    uint8_t *d_codes; NwPair *d_pairs; int8_t *d_matrix; uint32_t *d_dir; int32_t *d_bnd, *d_scores; char *d_ops = nullptr; int *d_len = nullptr;
    MDF_TRY(ctx->alloc_n(&d_codes, (size_t)total + 16));
    MDF_TRY(ctx->alloc_n(&d_pairs, all.size()));
    MDF_TRY(ctx->alloc_n(&d_matrix, 1024));
    MDF_TRY(ctx->alloc_n(&d_dir, (size_t)dir_words + 4));
    MDF_TRY(ctx->alloc_n(&d_bnd, (size_t)bnd_ints + 4));
    MDF_TRY(ctx->alloc_n(&d_scores, (size_t)n));
    if (full) { MDF_TRY(ctx->alloc_n(&d_ops, (size_t)ops_bytes + 4)); MDF_TRY(ctx->alloc_n(&d_len, (size_t)n)); }
    int8_t h_matrix[1024];
    memset(h_matrix, 0, sizeof h_matrix);
    for (int r = 0; r < A; ++r) for (int c = 0; c < A; ++c) h_matrix[r * 32 + c] = matrix[r * A + c];
    cudaStream_t s = ctx->stream;
    MDF_CUDA(cudaMemcpyAsync(d_codes, h_codes.data(), (size_t)total, cudaMemcpyHostToDevice, s));
    MDF_CUDA(cudaMemcpyAsync(d_pairs, all.data(), all.size() * sizeof(NwPair), cudaMemcpyHostToDevice, s));
    MDF_CUDA(cudaMemcpyAsync(d_matrix, h_matrix, 1024, cudaMemcpyHostToDevice, s));
    {
        ProfScope ps(ctx, "nw_align", 0.0);
        MDF_TRY(launch_class<4>(ctx, cls_at[1] - cls_at[0], d_pairs + cls_at[0], d_codes, d_matrix, A, gap_open, gap_extend, d_dir, d_bnd, d_scores, d_ops, d_len, full));
        MDF_TRY(launch_class<8>(ctx, cls_at[2] - cls_at[1], d_pairs + cls_at[1], d_codes, d_matrix, A, gap_open, gap_extend, d_dir, d_bnd, d_scores, d_ops, d_len, full));
        MDF_TRY(launch_class<16>(ctx, cls_at[3] - cls_at[2], d_pairs + cls_at[2], d_codes, d_matrix, A, gap_open, gap_extend, d_dir, d_bnd, d_scores, d_ops, d_len, full));
        MDF_TRY(launch_class<32>(ctx, cls_at[4] - cls_at[3], d_pairs + cls_at[3], d_codes, d_matrix, A, gap_open, gap_extend, d_dir, d_bnd, d_scores, d_ops, d_len, full));
    }
    MDF_CUDA(cudaMemcpyAsync(scores, d_scores, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (full) {
        MDF_CUDA(cudaMemcpyAsync(ops_len, d_len, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        if (ops_bytes) MDF_CUDA(cudaMemcpyAsync(ops, d_ops, (size_t)ops_bytes, cudaMemcpyDeviceToHost, s));
    }
    MDF_CUDA(cudaStreamSynchronize(s));
    return MDF_OK;
}
