// tcgen05 GEMM kernel (see gemm_tc.cuh).
#include <algorithm>

#include "gemm_tc.cuh"

namespace mdf {
namespace tc {

constexpr int GEMM_THREADS = 320;   // warp 0 producer, warp 1 MMA, warps 2-9 epilogue
// Adjacency instance (EPI_IMG_ROWSCALE, BN = 256): its K is one protein long (a handful of k-blocks), so the ELU
// epilogue, not the tensor pipe, sets the pace - it gets 16 epilogue warps (4 per scheduler, better latency hiding) and
// four more warps that expand the A tiles from the bit-packed contact map.
__host__ __device__ constexpr int gemm_epi_warps(int epi, int bn) { return (epi == EPI_IMG_ROWSCALE && bn == 256) ? 16 : 8; }
__host__ __device__ constexpr int gemm_threads(int epi, int bn, bool expand) { return (2 + gemm_epi_warps(epi, bn) + (expand ? 4 : 0)) * 32; }
// Warp roles, lowest warp id first: epilogue warps, (expander warps), producer, MMA issuer.  The SM's warp arbiter prefers
// the highest warp id among the eligible warps of a scheduler, so the single-thread roles that feed the tensor pipe sit
// at the top: with the issuer as warp 1 it starved behind the epilogue warps of its scheduler and the tensor pipe idled
// 70 % of the adjacency GEMM.

struct __align__(8) GemmBarriers {
    uint64_t full[8], empty[8], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float act_f(float v, int act, float alpha)
{
    if (act == 1) return fmaxf(v, 0.0f);
    if (act == 2) return v > 0.0f ? v : alpha * (__expf(v) - 1.0f);
    return v;
}

__device__ __forceinline__ int lane_id() { return (int)(threadIdx.x & 31); }

// column sums of a 32 x 32 block held one row per lane: butterfly transpose-reduce, lane l returns the total of column l
__device__ __forceinline__ float warp_column_sums(float (&c)[32], int lane)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; ++j) {
            const float send = upper ? c[j] : c[j + off];
            const float keep = upper ? c[j + off] : c[j];
            c[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return c[0];
}

// One 32-column chunk of an output row: fused scale / bias / activation / gather, then fp16 image or fp32 store.
template <int EPI>
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmArgs &g, uint32_t (&r)[32], int64_t m, int n0, float rs,
                                                    const float *grow, uint8_t *row_ptr, float *pool_row = nullptr,
                                                    int corr_mode = 0, float corr_c = 0.0f)
{
    if (EPI == EPI_F32_BIAS) {
        if (m < g.m_valid) {
            float *dst = g.out_f32 + (size_t)m * g.ldc + n0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (n0 + 4 * q < g.n_valid) {
                    float4 v;
                    const float4 b = g.bias ? __ldg(reinterpret_cast<const float4 *>(g.bias + n0 + 4 * q)) : make_float4(0, 0, 0, 0);
                    v.x = __uint_as_float(r[4 * q + 0]) + b.x; v.y = __uint_as_float(r[4 * q + 1]) + b.y;
                    v.z = __uint_as_float(r[4 * q + 2]) + b.z; v.w = __uint_as_float(r[4 * q + 3]) + b.w;
                    if (g.act == 1) { v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f); }
                    *reinterpret_cast<float4 *>(dst + 4 * q) = v;
                }
            }
        }
    } else {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (EPI == EPI_IMG_COLSCALE) {
            // per-protein correction of the weight rounding (tc_engine.cu, mean-corrected single term).  The caller looked the
            // chunk's columns up one chunk ahead: mode 1 = all 32 columns belong to one protein (corr_c is its value for this
            // row), mode 2 = the chunk crosses a protein boundary and every column is looked up here
            if (corr_mode == 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += corr_c;
            } else if (corr_mode == 2) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {                           // (fully unrolled: v[] must stay in registers)
                    const int pj = __ldg(g.col_group + min(n0 + j, g.col_valid - 1));
                    v[j] += __ldg(g.corr + (size_t)pj * g.corr_ld + m) * g.corr_scale;
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                // pair kernel: `grow` points into the shared-memory copy of this tile's 256 column scales (staged one tile ahead:
                // read from global memory here, the load's L2 latency sat in front of every 32-column chunk)
                const float4 cs = grow ? *reinterpret_cast<const float4 *>(grow + n0 + 4 * q)
                                       : __ldg(reinterpret_cast<const float4 *>(g.colscale + n0 + 4 * q));
                v[4 * q + 0] = cs.x != 0.0f ? v[4 * q + 0] * cs.x : 0.0f;
                v[4 * q + 1] = cs.y != 0.0f ? v[4 * q + 1] * cs.y : 0.0f;
                v[4 * q + 2] = cs.z != 0.0f ? v[4 * q + 2] * cs.z : 0.0f;
                v[4 * q + 3] = cs.w != 0.0f ? v[4 * q + 3] * cs.w : 0.0f;
            }
        } else if (EPI == EPI_IMG_ROWSCALE && (g.ablate & 1)) {
        } else if (EPI == EPI_IMG_ROWSCALE) {
            if (g.bias) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(g.bias + n0 + 4 * q));
                    v[4 * q + 0] = fmaf(v[4 * q + 0], rs, b.x); v[4 * q + 1] = fmaf(v[4 * q + 1], rs, b.y);
                    v[4 * q + 2] = fmaf(v[4 * q + 2], rs, b.z); v[4 * q + 3] = fmaf(v[4 * q + 3], rs, b.w);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= rs;
            }
            if (g.act == 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
            } else if (g.act == 2) {
                const float alpha = g.alpha;
#pragma unroll
                for (int j = 0; j < 32; ++j) {                                   // branch-free ELU, one MUFU + 5 ALU ops
                    const float e = fmaf(alpha, ex2_ftz(fminf(v[j], 0.0f) * 1.4426950408889634f), -alpha);
                    v[j] = v[j] > 0.0f ? v[j] : e;
                }
            }
        } else if (EPI == EPI_IMG_EMBED) {
            if (g.embed_staged) {
                // pair kernel: `grow` points into the shared-memory slice of (W_aa + b) for this tile's columns (see there)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 a = *reinterpret_cast<const float4 *>(grow + n0 + 4 * q);
                    v[4 * q + 0] = fmaxf(v[4 * q + 0] + a.x, 0.0f);
                    v[4 * q + 1] = fmaxf(v[4 * q + 1] + a.y, 0.0f);
                    v[4 * q + 2] = fmaxf(v[4 * q + 2] + a.z, 0.0f);
                    v[4 * q + 3] = fmaxf(v[4 * q + 3] + a.w, 0.0f);
                }
            } else
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 b = __ldg(reinterpret_cast<const float4 *>(g.bias + n0 + 4 * q));
                const float4 a = __ldg(reinterpret_cast<const float4 *>(grow + n0 + 4 * q));
                v[4 * q + 0] = fmaxf(v[4 * q + 0] + b.x + a.x, 0.0f);
                v[4 * q + 1] = fmaxf(v[4 * q + 1] + b.y + a.y, 0.0f);
                v[4 * q + 2] = fmaxf(v[4 * q + 2] + b.z + a.z, 0.0f);
                v[4 * q + 3] = fmaxf(v[4 * q + 3] + b.w + a.w, 0.0f);
            }
        }
        if (EPI == EPI_IMG_ROWSCALE && pool_row && !(g.ablate & 2)) {
            // fused sum-pool: column sums over the 32 rows of this warp (pad rows masked), butterfly transpose-reduce so
            // that lane l ends with the total of column n0 + l (31 shuffles), then one fp32 atomic per lane
            float c[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) c[j] = rs != 0.0f ? v[j] : 0.0f;
            const float tot = warp_column_sums(c, threadIdx.x & 31);
            if (!(g.ablate & 64) || tot == 1234.5f) atomicAdd(pool_row + n0 + lane_id(), tot);
        }
        // The adjacency product is bound by HBM bytes (Y^T in, X out), not by the tensor pipe: it never stores pad rows (nothing
        // reads them as values: rows are independent in X.W and its column scale zeroes their Y^T columns), and stores no image
        // at all when only the fused sum-pool consumes the layer (out_img == nullptr: the last GraphConv layer)
        if (EPI == EPI_IMG_ROWSCALE && (g.out_img == nullptr || (rs == 0.0f && g.skip_pad_rows) || ((g.ablate & 4) && v[0] != 1234.5f))) return;
        uint8_t *dst = row_ptr + (size_t)(n0 >> 6) * TILE_BYTES + (((n0 & 63) >> 3) * 2048);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 pk;
            pk.x = pack_half2(v[8 * q + 0], v[8 * q + 1]); pk.y = pack_half2(v[8 * q + 2], v[8 * q + 3]);
            pk.z = pack_half2(v[8 * q + 4], v[8 * q + 5]); pk.w = pack_half2(v[8 * q + 6], v[8 * q + 7]);
            *reinterpret_cast<uint4 *>(dst + q * 2048) = pk;
        }
    }
}

template <int EPI, int BN>
__global__ void __launch_bounds__(gemm_threads(EPI, BN, true), 1)
gemm_tc_kernel(GemmArgs g, int a_terms, int b_terms, int stages)
{
    constexpr int EW = gemm_epi_warps(EPI, BN);    // epilogue warps: 4 TMEM lane quarters x EW/4 column ranges
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int NSUB = BN / 128;                 // B sub-tiles (one 128x128x16 MMA each)
    __shared__ GemmBarriers bars;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_bytes = a_terms * TILE_BYTES;
    const int stage_bytes = a_bytes + b_terms * NSUB * TILE_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = blockDim.x >> 5;
    const int w_prod = n_warps - 2, w_mma = n_warps - 1;
    // One CTA walks a whole group of tiles that share an operand panel back to back (all feature tiles of one
    // residue block when m_fastest, all column tiles of one row block otherwise), so the shared panel is re-read
    // from L2 by the CTA that just touched it instead of by neighbours that may have drifted out of phase.
    const int inner = g.m_fastest ? g.m_tiles : g.n_tiles, outer = g.m_fastest ? g.n_tiles : g.m_tiles;

    if (threadIdx.x == 0) {
        // adjacency mode: a stage is full when the B copies landed AND the four expander warps finished the A tile
        for (int s = 0; s < stages; ++s) { mbar_init(&bars.full[s], g.adj_packed ? 5 : 1); mbar_init(&bars.empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bars.tmem_full[s], 1); mbar_init(&bars.tmem_empty[s], EW); }
        fence_mbar_init();
    }
    if (warp == w_mma) tmem_alloc<2 * BN>(&bars.tmem_base);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = bars.tmem_base;

    if (warp == w_prod) {
        // ===================== producer: bulk-TMA tile images into the stage ring
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int grp = blockIdx.x; grp < outer; grp += gridDim.x)
            for (int in = 0; in < inner; ++in) {
                const int mt = g.m_fastest ? in : grp, nt = g.m_fastest ? grp : in;
                int a_tile0, b_kb0, nkb;
                if (g.tile_info) { const int4 ti = g.tile_info[mt]; a_tile0 = ti.x; b_kb0 = ti.y; nkb = ti.z; }
                else { a_tile0 = mt * g.KB_A; b_kb0 = 0; nkb = g.nkb; }
                if (g.adj_kb_cnt) nkb = g.adj_kb_cnt[mt];
                for (int j = 0; j < nkb; ++j) {
                    const int kb = g.adj_kb_idx ? (int)g.adj_kb_idx[a_tile0 + j] : j;
                    mbar_wait(&bars.empty[st], ph ^ 1);
                    mbar_arrive_expect_tx(&bars.full[st], (uint32_t)(g.adj_packed ? stage_bytes - a_bytes : stage_bytes));
                    uint8_t *dst = smem + (size_t)st * stage_bytes;
                    if (g.adj_packed) {
                        // A tile comes from the expander warps
                    } else {
                        for (int ta = 0; ta < a_terms; ++ta)
                            bulk_g2s(dst + ta * TILE_BYTES,
                                     reinterpret_cast<const uint8_t *>(g.A[ta]) + (size_t)(a_tile0 + kb) * TILE_BYTES,
                                     TILE_BYTES, &bars.full[st]);
                    }
                    for (int tb = 0; tb < b_terms; ++tb)
                        for (int j = 0; j < NSUB; ++j)
                            bulk_g2s(dst + a_bytes + (tb * NSUB + j) * TILE_BYTES,
                                     reinterpret_cast<const uint8_t *>(g.B[tb]) + ((size_t)(nt * NSUB + j) * g.KB_B + b_kb0 + kb) * TILE_BYTES,
                                     TILE_BYTES, &bars.full[st]);
                    if (++st == stages) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == w_mma) {
        // ===================== MMA issuer: the whole warp walks the loop converged, one elected lane issues
        {
            constexpr uint32_t idesc = umma_idesc_f16(128, 128);
            int st = 0; uint32_t ph = 0;
            int acc = 0; uint32_t acc_ph = 0;
            const bool tr = g.trace && blockIdx.x == 0;
            const uint32_t full_bar0 = pin_u32(smem_u32(&bars.full[0])), empty_bar0 = pin_u32(smem_u32(&bars.empty[0]));
            const uint32_t tfull_bar0 = pin_u32(smem_u32(&bars.tmem_full[0]));
            const uint64_t desc0 = pin_u64(umma_smem_desc(smem_u32(smem), TILE_LBO, TILE_SBO));
            long long t_info = 0, t_acc = 0, t_full = 0, n_tiles_done = 0;
            const long long t_begin = tr ? clock64() : 0;
            for (int grp = blockIdx.x; grp < outer; grp += gridDim.x)
            for (int in = 0; in < inner; ++in) {
                const int mt = g.m_fastest ? in : grp;
                long long c0 = tr ? clock64() : 0;
                const int nkb = g.adj_kb_cnt ? g.adj_kb_cnt[mt] : g.tile_info ? g.tile_info[mt].z : g.nkb;
                if (tr) { const long long c1 = clock64() + (nkb & 0); t_info += c1 - c0; c0 = c1; }
                mbar_wait(&bars.tmem_empty[acc], acc_ph ^ 1);
                if (tr) { const long long c1 = clock64(); t_acc += c1 - c0; ++n_tiles_done; }
                tcgen05_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(acc * BN);
                if (EPI == EPI_IMG_ROWSCALE && BN == 256 && a_terms == 1 && b_terms == 1) {
                    // adjacency product: one wait + ONE elected asm block per k-block (eight MMAs + commits; bases pinned, descriptors
                    // = base + stage offset).  With an elect per MMA the issuing warp needed ~1 k cycles per k-block against 512 of
                    // tensor work (same finding as in the fused LSTM kernel).
                    for (int kb = 0; kb < nkb; ++kb) {
                        c0 = tr ? clock64() : 0;
                        mbar_wait1_asm(full_bar0 + 8u * st, ph);
                        if (tr) t_full += clock64() - c0;
                        const uint64_t ad = desc0 + (uint64_t)(st * (stage_bytes >> 4));
                        if (g.ablate & 16) {
                            umma_commit_elect(&bars.empty[st]);
                            if (kb == nkb - 1) umma_commit_elect(&bars.tmem_full[acc]);
                        } else
                        umma_f16_kblock_2n128_elect(d0, ad, ad + (uint64_t)(TILE_BYTES >> 4), ad + (uint64_t)(2 * TILE_BYTES >> 4), idesc, kb ? 1u : 0u,
                                                    empty_bar0 + 8u * st, kb == nkb - 1 ? tfull_bar0 + 8u * acc : 0u);
                        if (++st == stages) { st = 0; ph ^= 1; }
                    }
                    if (nkb == 0) umma_commit_elect(&bars.tmem_full[acc]);
                    __syncwarp();
                    if (++acc == 2) { acc = 0; acc_ph ^= 1; }
                    continue;
                }
                for (int kb = 0; kb < nkb; ++kb) {
                    c0 = tr ? clock64() : 0;
                    mbar_wait(&bars.full[st], ph);
                    if (tr) t_full += clock64() - c0;
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)st * stage_bytes);
                    const uint32_t sb = sa + a_bytes;
                    {
                        // every (A term, B term) pair contributes; at most one side has two terms
                        for (int ta = 0; ta < a_terms; ++ta)
                            for (int tb = (a_terms == 3 && ta == 2) ? 1 : 0;
                                 tb < ((a_terms == 3) ? (ta == 2 ? 2 : 1) : b_terms - (ta == 1 && b_terms == 2 ? 1 : 0)); ++tb)
                            // a_terms 2 x b_terms 2: hi.hi, hi.lo, lo.hi (lo.lo skipped); a_terms 3: (hi, hi), (lo, hi), (hi*2^-11, lo*2^11)
#pragma unroll
                                for (int ks = 0; ks < TILE_K / 16; ++ks) {
                                    const uint64_t ad = umma_smem_desc(sa + ta * TILE_BYTES + ks * 2 * TILE_LBO, TILE_LBO, TILE_SBO);
#pragma unroll
                                    for (int j = 0; j < NSUB; ++j) {
                                        const uint64_t bd = umma_smem_desc(sb + (tb * NSUB + j) * TILE_BYTES + ks * 2 * TILE_LBO,
                                                                           TILE_LBO, TILE_SBO);
                                        umma_f16_elect(d0 + j * 128, ad, bd, idesc, (kb | ks | ta | tb) != 0);
                                    }
                                }
                    }
                    umma_commit_elect(&bars.empty[st]);               // frees the smem stage when the MMAs retire
                    if (kb == nkb - 1) umma_commit_elect(&bars.tmem_full[acc]);
                    __syncwarp();
                    if (++st == stages) { st = 0; ph ^= 1; }
                }
                if (nkb == 0) umma_commit_elect(&bars.tmem_full[acc]);
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
            }
            if (tr && lane == 0) { g.trace[0] = clock64() - t_begin; g.trace[1] = t_info; g.trace[2] = t_acc; g.trace[3] = t_full; g.trace[4] = n_tiles_done; }
        }
    } else if (warp >= EW) {
        // ===================== A-tile expanders (adjacency mode, 128 threads): thread = one row of the 128 x 64 tile
        // (measured: a second group of four warps alternating stages is slower - the kernel is issue-bound, not
        // expander-bound)
        const int et = threadIdx.x - EW * 32;
        int st = 0; uint32_t ph = 0;
        for (int grp = blockIdx.x; grp < outer; grp += gridDim.x)
        for (int in = 0; in < inner; ++in) {
            const int mt = g.m_fastest ? in : grp;
            const int4 ti = g.tile_info[mt];
            const int nkb = g.adj_kb_cnt ? g.adj_kb_cnt[mt] : ti.z, p = ti.w;
            const int L = (int)(g.adj_seq_off[p + 1] - g.adj_seq_off[p]);
            const int rw = packed_row_words(L);
            const int i = (mt - (int)(g.adj_seg_off[p] >> 7)) * 128 + et;                 // row of the protein's map
            const uint32_t *row = g.adj_packed + g.adj_packed_off[p] + (size_t)i * rw;
            if (g.adj_kb_idx) {
                // block-sparse walk: the two words of listed k-block j, loaded one list entry ahead of the stage that uses them
                const unsigned short *list = g.adj_kb_idx + ti.x;
                // compact axis: listed block j = compact residues 64 (ti.y + j) .. + 63 = bits j0 .. j0 + 63 of this protein's row
                const int jbase = g.adj_compact ? (int)(64ll * ti.y - g.adj_seq_off[p]) : 0;
                auto window = [&](int j) -> uint2 {
                    if (g.adj_compact) return adj_row_window(row, rw, jbase + 64 * j);
                    return __ldg(reinterpret_cast<const uint2 *>(row + 2 * j));
                };
                uint2 nx = make_uint2(0u, 0u);
                if (nkb > 0 && i < L) nx = window((int)list[0]);
                for (int j = 0; j < nkb; ++j) {
                    const uint2 cw = nx;
                    nx = make_uint2(0u, 0u);
                    if (j + 1 < nkb && i < L) nx = window((int)list[j + 1]);
                    const uint32_t w[2] = {cw.x, cw.y};
                    if (lane == 0) mbar_wait(&bars.empty[st], ph ^ 1);
                    __syncwarp();
                    const uint32_t dst = smem_u32(smem + (size_t)st * stage_bytes) + (uint32_t)((et >> 3) * 128 + (et & 7) * 16);
                    const uint32_t one = 0x3C00u;
                    if (!(g.ablate & 8))
#pragma unroll
                    for (int k8 = 0; k8 < 8; ++k8) {
                        const uint32_t b8 = (w[k8 >> 2] >> (8 * (k8 & 3))) & 0xFFu;
                        uint4 pk;
                        pk.x = ((b8 & 1u) ? one : 0u) | ((b8 & 2u) ? one << 16 : 0u);
                        pk.y = ((b8 & 4u) ? one : 0u) | ((b8 & 8u) ? one << 16 : 0u);
                        pk.z = ((b8 & 16u) ? one : 0u) | ((b8 & 32u) ? one << 16 : 0u);
                        pk.w = ((b8 & 64u) ? one : 0u) | ((b8 & 128u) ? one << 16 : 0u);
                        st_shared_v4(dst + k8 * 2048, pk);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.full[st]);
                    if (++st == stages) { st = 0; ph ^= 1; }
                }
                continue;
            }
            // the packed row is read four words (two k-blocks) at a time, one load AHEAD of the stage that consumes it: a
            // load issued right before its stage put ~1 k cycles of L2 latency on every stage and paced the whole kernel
            uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
            if (i < L && 0 < rw) nxt = __ldg(reinterpret_cast<const uint4 *>(row));
            uint4 cur = nxt;
            for (int kb = 0; kb < nkb; ++kb) {
                {
                    if ((kb & 1) == 0) {
                        cur = nxt;
                        nxt = make_uint4(0u, 0u, 0u, 0u);
                        if (i < L && 2 * (kb + 2) < rw) nxt = __ldg(reinterpret_cast<const uint4 *>(row + 2 * (kb + 2)));
                    }
                    const uint32_t w[2] = {(kb & 1) ? cur.z : cur.x, (kb & 1) ? cur.w : cur.y};
                    if (lane == 0) mbar_wait(&bars.empty[st], ph ^ 1);
                    __syncwarp();
                    const uint32_t dst = smem_u32(smem + (size_t)st * stage_bytes) + (uint32_t)((et >> 3) * 128 + (et & 7) * 16);
                    const uint32_t one = 0x3C00u;                                            // fp16 1.0
#pragma unroll
                    for (int k8 = 0; k8 < 8; ++k8) {
                        const uint32_t b8 = (w[k8 >> 2] >> (8 * (k8 & 3))) & 0xFFu;
                        uint4 pk;
                        pk.x = ((b8 & 1u) ? one : 0u) | ((b8 & 2u) ? one << 16 : 0u);
                        pk.y = ((b8 & 4u) ? one : 0u) | ((b8 & 8u) ? one << 16 : 0u);
                        pk.z = ((b8 & 16u) ? one : 0u) | ((b8 & 32u) ? one << 16 : 0u);
                        pk.w = ((b8 & 64u) ? one : 0u) | ((b8 & 128u) ? one << 16 : 0u);
                        st_shared_v4(dst + k8 * 2048, pk);
                    }
                    fence_proxy_async_smem();                    // generic-proxy tile -> tensor-core (async proxy) reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.full[st]);
                }
                if (++st == stages) { st = 0; ph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> global (8 warps: 4 lane blocks x 2 column halves)
        const int lb = (warp & 3) * 32;          // this warp's TMEM lane block = output rows
        const int ch = warp >> 2;                // column range handled by this warp
        int acc = 0; uint32_t acc_ph = 0;
        for (int grp = blockIdx.x; grp < outer; grp += gridDim.x)
        for (int in = 0; in < inner; ++in) {
            const int mt = g.m_fastest ? in : grp, nt = g.m_fastest ? grp : in;
            const int nkb = g.adj_kb_cnt ? g.adj_kb_cnt[mt] : g.tile_info ? g.tile_info[mt].z : g.nkb;
            mbar_wait(&bars.tmem_full[acc], acc_ph);
            tcgen05_fence_after();
            int64_t m = (int64_t)mt * 128 + lb + lane;
            bool row_valid = true;
            if (EPI == EPI_IMG_ROWSCALE && g.adj_compact) {
                // row i of protein p lives at compact row seq_off[p] + i; rows past L are padding of the tile only
                const int p = g.tile_info[mt].w;
                const int64_t s0 = g.adj_seq_off[p];
                const int i = (mt - (int)(g.adj_seg_off[p] >> 7)) * 128 + lb + lane;
                row_valid = i < (int)(g.adj_seq_off[p + 1] - s0);
                m = row_valid ? s0 + i : s0;
            }
            const uint32_t trow = tmem_base + ((uint32_t)lb << 16) + (uint32_t)(acc * BN);
            float rs = 1.0f;
            const float *grow = nullptr;
            if (EPI == EPI_IMG_ROWSCALE) rs = row_valid ? g.rowscale[m] : 0.0f;
            if (EPI == EPI_IMG_EMBED) grow = g.gtab + (size_t)g.gidx[m] * g.ldg;
            float *pool_row = nullptr;
            if (EPI == EPI_IMG_ROWSCALE && g.pool && g.tile_info) pool_row = g.pool + (size_t)g.tile_info[mt].w * g.pool_ld + g.pool_off;
            // byte address of (row m, k = 0) in the output image; a 32-column chunk stays inside one k-block
            uint8_t *row_ptr = reinterpret_cast<uint8_t *>(g.out_img) + (size_t)(m >> 7) * g.KB_out * TILE_BYTES +
                               (size_t)((((int)m & 127) >> 3) * 128 + ((int)m & 7) * 16);
#pragma unroll 1
            for (int c0 = ch * (BN * 4 / EW); c0 < (ch + 1) * (BN * 4 / EW); c0 += 32) {
                uint32_t r[32];
                if (nkb > 0 && !(g.ablate & 32)) {
                    tmem_ld_32x32b_x32(trow + c0, r);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                gemm_epilogue_chunk<EPI>(g, r, m, nt * BN + c0, rs, grow, row_ptr, pool_row);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == w_mma) tmem_dealloc<2 * BN>(tmem_base);
}


// ------------------------------------------------------------------------------------------- CTA-pair GEMM (cta_group::2)
// 256 x 256 output tiles on two SMs: each CTA stages its own 128 rows of A and its own 128 rows of B (per term),
// the leader issues M = 256, N = 256 MMAs that read both CTAs' shared memory, and each CTA's TMEM receives its
// 128 rows x 256 columns.  Per SM that is half the L2 -> SM operand traffic per FLOP of the 128 x 128/256 single-CTA
// tiles - the weight GEMMs (K = 512 / 1024, two fp16 terms on the weight side) are bound by that traffic, not by
// the tensor pipe.  Operand tiles arrive by tensor-map TMA (cta_group::2 form) so that both CTAs' copies complete
// on the leader's barrier.
struct PairGemmArgs {
    GemmArgs g;                        // m_tiles / n_tiles count 256-row / 256-column tiles here
    alignas(64) CUtensorMap tmA[2];    // flat [bytes/512][256] u16 maps over the A / B term images
    alignas(64) CUtensorMap tmB[2];
    int a_terms, b_terms, stages;
};

#ifndef MDF_PAIR_EPI_WARPS
#define MDF_PAIR_EPI_WARPS 8
#endif
constexpr int EMB_ROWS = 26, EMB_LD = 256 + 4;            // staged embedding-table slice: 26 residue types x 256 columns, rows padded by 16 B
constexpr int PAIR_EW = MDF_PAIR_EPI_WARPS;               // epilogue warps per CTA (8 or 16)
constexpr int PAIR_THREADS = (PAIR_EW + 2) * 32;
constexpr int CS_SLICE = 256 + 16;                        // EPI_IMG_COLSCALE: 256 column scales + 8 chunk groups (+ pad) per staged tile slice

template <int EPI>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ PairGemmArgs pa)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int BN = 256;
    __shared__ GemmBarriers bars;
    const GemmArgs &g = pa.g;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_terms = pa.a_terms, b_terms = pa.b_terms, stages = pa.stages;
    const int a_bytes = a_terms * TILE_BYTES;
    const int stage_bytes = a_bytes + b_terms * TILE_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int w_prod = PAIR_EW, w_mma = PAIR_EW + 1;     // warps 0..PAIR_EW-1 epilogue, then producer, MMA issuer
    const int rank = (int)cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int inner = g.m_fastest ? g.m_tiles : g.n_tiles, outer = g.m_fastest ? g.n_tiles : g.m_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bars.tmem_full[s], 1); mbar_init(&bars.tmem_empty[s], 2 * PAIR_EW); }
        fence_mbar_init();
    }
    if (warp == w_mma) tmem_alloc_pair<2 * BN>(&bars.tmem_base);
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = bars.tmem_base;

    if (warp == w_prod) {
        // ===================== producer (both CTAs): my 128 rows of every A / B term tile; bytes land on the leader's barrier.
        // (Measured alternative: the leader issuing the peer's copies too is 10-20 % slower - one SM's TMA engine then
        // carries the traffic of two.)
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int grp = pair; grp < outer; grp += n_pairs)
            for (int in = 0; in < inner; ++in) {
                const int mt = g.m_fastest ? in : grp, nt = g.m_fastest ? grp : in;
                for (int kb = 0; kb < g.nkb; ++kb) {
                    mbar_wait(&bars.empty[st], ph ^ 1);
                    if (leader) mbar_arrive_expect_tx(&bars.full[st], (uint32_t)(2 * stage_bytes));
                    uint8_t *dst = smem + (size_t)st * stage_bytes;
                    for (int ta = 0; ta < a_terms; ++ta)
                        tma_tile_g2s_pair(dst + ta * TILE_BYTES, &pa.tmA[ta], ((mt * 2 + rank) * g.KB_A + kb) * (TILE_BYTES / 512), &bars.full[st]);
                    for (int tb = 0; tb < b_terms; ++tb)
                        tma_tile_g2s_pair(dst + a_bytes + tb * TILE_BYTES, &pa.tmB[tb], ((nt * 2 + rank) * g.KB_B + kb) * (TILE_BYTES / 512),
                                          &bars.full[st]);
                    if (++st == stages) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == w_mma) {
        // ===================== MMA issuer (leader CTA): converged warp, elected lane issues
        if (leader) {
            constexpr uint32_t idesc = umma_idesc_f16(256, 256);
            const uint32_t full_bar0 = pin_u32(smem_u32(&bars.full[0])), empty_bar0 = pin_u32(smem_u32(&bars.empty[0]));
            const uint32_t tfull_bar0 = pin_u32(smem_u32(&bars.tmem_full[0]));
            const uint64_t desc0 = pin_u64(umma_smem_desc(smem_u32(smem), TILE_LBO, TILE_SBO));
            int st = 0; uint32_t ph = 0;
            int acc = 0; uint32_t acc_ph = 0;
            for (int grp = pair; grp < outer; grp += n_pairs)
            for (int in = 0; in < inner; ++in) {
                mbar_wait(&bars.tmem_empty[acc], acc_ph ^ 1);
                tcgen05_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < g.nkb; ++kb) {
                    // one wait + one elected asm block per (A term, B term) pair: four K = 16 MMAs, the last pair also carries the
                    // commits (stage free in both CTAs; accumulator full after the last k-block)
                    mbar_wait1_asm(full_bar0 + 8u * st, ph);
                    const uint64_t sdesc = desc0 + (uint64_t)(st * (stage_bytes >> 4));
                    const int npairs = a_terms * b_terms;
                    for (int tp = 0; tp < npairs; ++tp) {
                        const int ta = tp / b_terms, tb = tp % b_terms;
                        const bool last = tp == npairs - 1;
                        umma_f16_pair_kblock_elect(d0, sdesc + (uint64_t)(ta * (TILE_BYTES >> 4)), sdesc + (uint64_t)((a_bytes + tb * TILE_BYTES) >> 4), idesc,
                                                   (kb | tp) ? 1u : 0u, 0u, 0u, last ? empty_bar0 + 8u * st : 0u, (uint16_t)3,
                                                   (last && kb == g.nkb - 1) ? tfull_bar0 + 8u * acc : 0u, (uint16_t)3);
                    }
                    if (++st == stages) { st = 0; ph ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (both CTAs): my 128 rows x 256 columns
        const int lb = (warp & 3) * 32;
        const int ch = warp >> 2;
        int acc = 0; uint32_t acc_ph = 0;
        // EPI_IMG_COLSCALE: the 256 column scales of a tile live in shared memory, double-buffered - thread t fetches column t of
        // the NEXT tile while the current one is processed and parks it after the last chunk; one named barrier per tile
        // (+ the tile's eight chunk groups of the mean correction behind the 256 scales: CS_SLICE words per buffer)
        float *cs_tab = reinterpret_cast<float *>(smem + (size_t)stages * stage_bytes);
        auto tile_nt = [&](int grp_, int in_) { return g.m_fastest ? grp_ : in_; };
        int tile_no = 0;
        const bool cg_thread = EPI == EPI_IMG_COLSCALE && g.chunk_group && threadIdx.x < BN / 32;
        if (EPI == EPI_IMG_COLSCALE && pair < outer) {
            if (threadIdx.x < BN) cs_tab[threadIdx.x] = __ldg(g.colscale + (size_t)tile_nt(pair, 0) * BN + threadIdx.x);
            if (cg_thread) reinterpret_cast<int *>(cs_tab)[BN + threadIdx.x] = __ldg(g.chunk_group + (size_t)tile_nt(pair, 0) * (BN / 32) + threadIdx.x);
            asm volatile("bar.sync 2, %0;" ::"n"(PAIR_EW * 32) : "memory");
        }
        for (int grp = pair; grp < outer; grp += n_pairs)
        for (int in = 0; in < inner; ++in) {
            const int mt = g.m_fastest ? in : grp, nt = g.m_fastest ? grp : in;
            const int64_t m = ((int64_t)mt * 2 + rank) * 128 + lb + lane;
            float rs = 1.0f;
            const float *grow = nullptr;
            float cs_next = 0.0f;
            int cg_next = -2;
            bool have_next = false;
            const int *cg_cur = reinterpret_cast<const int *>(cs_tab) + (tile_no & 1) * CS_SLICE + BN;
            if (EPI == EPI_IMG_COLSCALE) {
                const int in2 = in + 1 < inner ? in + 1 : 0, grp2 = in + 1 < inner ? grp : grp + n_pairs;
                have_next = grp2 < outer;
                if (have_next && threadIdx.x < BN) cs_next = __ldg(g.colscale + (size_t)tile_nt(grp2, in2) * BN + threadIdx.x);
                if (have_next && cg_thread) cg_next = __ldg(g.chunk_group + (size_t)tile_nt(grp2, in2) * (BN / 32) + threadIdx.x);
                grow = cs_tab + (tile_no & 1) * CS_SLICE - nt * BN;                            // the chunk adds n0 = nt * BN + c0
            }
            if (EPI == EPI_IMG_EMBED) {
                if (g.embed_staged) {
                    // The one-hot embedding is a gather of W_aa[idx[m]] per ROW, i.e. per lane: from global memory every LDG.128
                    // touches 32 different lines and the LSU, not the tensor pipe, set the pace (tensor 74 % active).  The slice
                    // of the 26-row table for this tile's 256 columns, bias folded in, is staged in shared memory instead (rows
                    // padded by 16 B so that different residues hit different banks) while the tile's MMAs are still running.
                    float *tab = reinterpret_cast<float *>(smem + (size_t)stages * stage_bytes);
                    asm volatile("bar.sync 2, %0;" ::"n"(PAIR_EW * 32) : "memory");       // every warp is done with the previous slice
                    for (int e = threadIdx.x; e < EMB_ROWS * (BN / 4); e += PAIR_EW * 32) {
                        const int row = e / (BN / 4), c4 = (e % (BN / 4)) * 4;
                        float4 w = __ldg(reinterpret_cast<const float4 *>(g.gtab + (size_t)row * g.ldg + nt * BN + c4));
                        if (g.bias) {
                            const float4 bb = __ldg(reinterpret_cast<const float4 *>(g.bias + nt * BN + c4));
                            w.x += bb.x; w.y += bb.y; w.z += bb.z; w.w += bb.w;
                        }
                        *reinterpret_cast<float4 *>(tab + row * EMB_LD + c4) = w;
                    }
                    asm volatile("bar.sync 2, %0;" ::"n"(PAIR_EW * 32) : "memory");
                    grow = tab + (size_t)g.gidx[m] * EMB_LD - nt * BN;                     // the chunk adds n0 = nt * BN + c0
                } else {
                    grow = g.gtab + (size_t)g.gidx[m] * g.ldg;
                }
            }
            const uint32_t trow = tmem_base + ((uint32_t)lb << 16) + (uint32_t)(acc * BN);
            if (EPI == EPI_IMG_ROWSCALE) rs = g.rowscale[m];
            uint8_t *row_ptr = reinterpret_cast<uint8_t *>(g.out_img) + (size_t)(m >> 7) * g.KB_out * TILE_BYTES +
                               (size_t)((((int)m & 127) >> 3) * 128 + ((int)m & 7) * 16);
            // mean-corrected single term: the correction of a chunk (two dependent loads) is fetched one chunk ahead, the first
            // one before the accumulator is awaited - fetched inside the chunk, the load chain paced the whole kernel
            int corr_mode_n = 0;
            float corr_c_n = 0.0f;
            auto corr_fetch = [&](int n0) {
                corr_mode_n = 0; corr_c_n = 0.0f;
                if (EPI == EPI_IMG_COLSCALE && g.corr) {
                    const int cg = cg_cur[(n0 - nt * BN) >> 5];          // staged with the column scales: no global load in the chain
                    if (cg >= 0) { corr_mode_n = 1; corr_c_n = __ldg(g.corr + (size_t)cg * g.corr_ld + m) * g.corr_scale; }
                    else if (cg == -1) corr_mode_n = 2;
                }
            };
            constexpr int c_begin_stride = BN * 4 / PAIR_EW;
            corr_fetch(nt * BN + ch * c_begin_stride);
            mbar_wait(&bars.tmem_full[acc], acc_ph);
            tcgen05_fence_after();
#pragma unroll 1
            for (int c0 = ch * c_begin_stride; c0 < (ch + 1) * c_begin_stride; c0 += 32) {
                uint32_t r[32];
                const int corr_mode = corr_mode_n;
                const float corr_c = corr_c_n;
                if (c0 + 32 < (ch + 1) * c_begin_stride) corr_fetch(nt * BN + c0 + 32);
                tmem_ld_32x32b_x32(trow + c0, r);
                tmem_ld_wait();
                gemm_epilogue_chunk<EPI>(g, r, m, nt * BN + c0, rs, grow, row_ptr, nullptr, corr_mode, corr_c);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(&bars.tmem_empty[acc]); else mbar_arrive_remote(&bars.tmem_empty[acc], 0);
            }
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
            if (EPI == EPI_IMG_COLSCALE) {
                ++tile_no;
                if (have_next && threadIdx.x < BN) cs_tab[(tile_no & 1) * CS_SLICE + threadIdx.x] = cs_next;
                if (have_next && cg_thread) reinterpret_cast<int *>(cs_tab)[(tile_no & 1) * CS_SLICE + BN + threadIdx.x] = cg_next;
                asm volatile("bar.sync 2, %0;" ::"n"(PAIR_EW * 32) : "memory");
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == w_mma) tmem_dealloc_pair<2 * BN>(tmem_base);
}

int make_tile_map(CUtensorMap *map, const void *base, size_t bytes)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        MDF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is unavailable"); return MDF_ECUDA; }
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t dims[2] = {256, (cuuint64_t)(bytes / 512)};
    const cuuint64_t strides[1] = {512};
    const cuuint32_t box[2] = {256, 32};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return MDF_ECUDA; }
    return MDF_OK;
}

template <int EPI>
static int launch_pair(mdf_ctx *ctx, int a_terms, int b_terms, const GemmArgs &args, const size_t a_bytes[2], const size_t b_bytes[2])
{
    PairGemmArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.g = args;
    pa.a_terms = a_terms; pa.b_terms = b_terms;
    const int stage_bytes = (a_terms + b_terms) * TILE_BYTES;
    pa.stages = std::min(8, (int)((200 * 1024) / stage_bytes));
    for (int t = 0; t < a_terms; ++t) MDF_TRY(make_tile_map(&pa.tmA[t], args.A[t], a_bytes[t]));
    for (int t = 0; t < b_terms; ++t) MDF_TRY(make_tile_map(&pa.tmB[t], args.B[t], b_bytes[t]));
    size_t smem = (size_t)pa.stages * stage_bytes + 1024;
    if (EPI == EPI_IMG_COLSCALE) smem += 2 * CS_SLICE * sizeof(float);      // double-buffered column-scale slice
    if (EPI == EPI_IMG_EMBED && pa.g.embed_staged) {
        const size_t tab = (size_t)EMB_ROWS * EMB_LD * sizeof(float);
        while (pa.stages > 2 && smem + tab > 226 * 1024) { --pa.stages; smem -= stage_bytes; }
        smem += tab;
    }
    auto kern = gemm_pair_kernel<EPI>;
    MDF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int outer = args.m_fastest ? args.n_tiles : args.m_tiles;
    if (outer <= 0 || args.m_tiles * args.n_tiles <= 0) return MDF_OK;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * std::min(outer, ctx->sm_count / 2));
    cfg.blockDim = dim3(PAIR_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    MDF_CUDA(cudaLaunchKernelEx(&cfg, kern, pa));
    ctx->launches++;
    return MDF_OK;
}

// m_tiles / n_tiles in `args` count 256-row / 256-column tiles; a_bytes / b_bytes = sizes of the term images
int launch_gemm_pair(mdf_ctx *ctx, int epi, int a_terms, int b_terms, const GemmArgs &args, const size_t a_bytes[2], const size_t b_bytes[2])
{
    if (a_terms < 1 || a_terms > 2 || b_terms < 1 || b_terms > 2 || (a_terms == 2 && b_terms == 2) || args.tile_info) {
        set_error("gemm_pair: unsupported configuration");
        return MDF_EUNSUPPORTED;
    }
    if (epi == EPI_IMG_EMBED) return launch_pair<EPI_IMG_EMBED>(ctx, a_terms, b_terms, args, a_bytes, b_bytes);
    if (epi == EPI_IMG_COLSCALE) return launch_pair<EPI_IMG_COLSCALE>(ctx, a_terms, b_terms, args, a_bytes, b_bytes);
    set_error("gemm_pair: no kernel for epilogue %d", epi);
    return MDF_EUNSUPPORTED;
}

template <int EPI, int BN>
static int launch_one(mdf_ctx *ctx, int a_terms, int b_terms, const GemmArgs &args)
{
    const int stage_bytes = (a_terms + b_terms * (BN / 128)) * TILE_BYTES;
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) { set_error("gemm_tc: stage of %d bytes does not fit twice in shared memory", stage_bytes); return MDF_EUNSUPPORTED; }
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    auto kern = gemm_tc_kernel<EPI, BN>;
    MDF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int total = args.m_tiles * args.n_tiles;
    if (total <= 0) return MDF_OK;
    const int outer = args.m_fastest ? args.n_tiles : args.m_tiles;
    const int grid = outer < ctx->sm_count ? outer : ctx->sm_count;
    kern<<<grid, gemm_threads(EPI, BN, args.adj_packed != nullptr), smem, ctx->stream>>>(args, a_terms, b_terms, stages);
    MDF_LAUNCH_CHECK(ctx);
    return MDF_OK;
}

int launch_gemm_tc(mdf_ctx *ctx, int epi, int bn, int a_terms, int b_terms, const GemmArgs &args)
{
    if (a_terms < 1 || a_terms > 3 || b_terms < 1 || b_terms > 2 || (a_terms == 3 && b_terms != 2)) {
        set_error("gemm_tc: unsupported term split %d x %d", a_terms, b_terms);
        return MDF_EUNSUPPORTED;
    }
#define MDF_GEMM_CASE(E, N) if (epi == E && bn == N) return launch_one<E, N>(ctx, a_terms, b_terms, args)
    MDF_GEMM_CASE(EPI_F32_BIAS, 128);
    MDF_GEMM_CASE(EPI_F32_BIAS, 256);
    MDF_GEMM_CASE(EPI_IMG_COLSCALE, 128);
    MDF_GEMM_CASE(EPI_IMG_COLSCALE, 256);
    MDF_GEMM_CASE(EPI_IMG_ROWSCALE, 128);
    MDF_GEMM_CASE(EPI_IMG_ROWSCALE, 256);
    MDF_GEMM_CASE(EPI_IMG_EMBED, 128);
    MDF_GEMM_CASE(EPI_IMG_EMBED, 256);
#undef MDF_GEMM_CASE
    set_error("gemm_tc: no kernel for epilogue %d / BN %d", epi, bn);
    return MDF_EUNSUPPORTED;
}

}  // namespace tc
}  // namespace mdf
