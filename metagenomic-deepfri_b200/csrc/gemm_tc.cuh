// tcgen05 GEMM over fp16 operand tile images (see tc_ptx.cuh) with fused epilogues.
//   D[M x N] (fp32, TMEM) = sum_terms A_t[M x K] . B_t[N x K]^T
// Persistent, warp-specialised: warp 0 = bulk-TMA producer, warp 1 = MMA issuer (one elected lane),
// warps 2..5 = epilogue (TMEM -> registers -> global).  Accumulators are double-buffered in TMEM so
// the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include <cuda.h>

#include "mdf_common.cuh"
#include "tc_ptx.cuh"

namespace mdf {
namespace tc {

enum Epi : int {
    EPI_F32_BIAS = 0,   // out_f32[m * ldc + n] = acc + bias[n]
    EPI_IMG_COLSCALE,   // out image (rows = m, k = n): colscale[n] != 0 ? (acc [+ corr]) * colscale[n] : 0
    EPI_IMG_ROWSCALE,   // out image: act(acc * rowscale[m] + bias[n])
    EPI_IMG_EMBED,      // out image: relu(acc + bias[n] + gtab[gidx[m] * ldg + n])
};

struct GemmArgs {
    const __half *A[3] = {nullptr, nullptr, nullptr};
    const __half *B[2] = {nullptr, nullptr};
    int KB_A = 0, KB_B = 0;          // k-blocks per row tile of the A / B images
    int m_tiles = 0, n_tiles = 0;    // output tiles (128 rows x BN columns)
    int nkb = 0;                     // k-blocks per output tile (regular case)
    int m_fastest = 0;               // tile order: consecutive CTAs walk M first (B tile shared in L2) instead of N first
    // grouped case (adjacency product): per m-tile {A tile index of its first k-block, first k-block
    // on the B side, number of k-blocks, unused}
    const int4 *tile_info = nullptr;
    // grouped case with on-the-fly A tiles: instead of reading an expanded fp16 A_hat image (2 B per matrix entry, read
    // once per column tile) four extra warps build every 128 x 64 A tile in shared memory from the bit-packed contact
    // map (1 bit per entry).  adj_packed != nullptr selects it; A[] is then unused.
    // sum-pool readout fused into the EPI_IMG_ROWSCALE epilogue (grouped case): pool[protein * pool_ld + pool_off + n] +=
    // fp32 activations of the valid rows (rowscale != 0) of the tile; protein = tile_info[mt].w
    float *pool = nullptr;
    int pool_ld = 0, pool_off = 0;
    long long *trace = nullptr;      // optional [8] cycle counters of CTA 0 (MDF_GEMM_TRACE=1, adjacency GEMM)
    // block-sparse adjacency (tc_engine.cu adj_tile_scan_kernel): the k-blocks of an m-tile whose 128 x 64 A tile holds at
    // least one contact, as a compact list adj_kb_idx[tile_info[mt].x + j], j < adj_kb_cnt[mt].  All-zero tiles contribute
    // nothing to A_hat . Y, so the producer, the expanders and the MMA issuer walk this list instead of 0..nkb-1.
    int adj_compact = 0;             // adjacency product on the compact residue axis (tc_engine.cu build_meta)
    const unsigned short *adj_kb_idx = nullptr;
    const int *adj_kb_cnt = nullptr;
    const uint32_t *adj_packed = nullptr;
    const int64_t *adj_packed_off = nullptr, *adj_seq_off = nullptr, *adj_seg_off = nullptr;
    // epilogue
    float *out_f32 = nullptr;
    int ldc = 0;
    __half *out_img = nullptr;       // EPI_IMG_ROWSCALE: nullptr = no image (only the fused sum-pool consumes the tile)
    int skip_pad_rows = 0;           // EPI_IMG_ROWSCALE: rows with rowscale == 0 (padding) are not stored
    int KB_out = 0;
    const float *bias = nullptr;
    const float *rowscale = nullptr;
    const float *colscale = nullptr;
    // EPI_IMG_COLSCALE with ONE weight term (tc_engine.cu, "mean-corrected single term"): acc + corr_scale * corr[group(n) * corr_ld + m]
    // before the column scale; col_group[n] = protein of image column n (n < col_valid; columns past it are padding)
    const float *corr = nullptr;
    const int *col_group = nullptr;
    const int *chunk_group = nullptr;   // per 32 image columns: their protein if all 32 share one, -1 if the chunk straddles proteins, -2 = padding
    int corr_ld = 0, col_valid = 0;
    float corr_scale = 1.0f;
    const float *gtab = nullptr;
    const uint8_t *gidx = nullptr;
    int ldg = 0;
    int embed_staged = 0;            // pair kernel, EPI_IMG_EMBED: gather from a shared-memory slice of (gtab + bias) instead of global memory
    int act = 0;
    float alpha = 1.0f;
    int m_valid = 0, n_valid = 0;    // bounds for the fp32 epilogue
    int ablate = 0;                  // MDF_ADJ_ABLATE (timing experiments, wrong results): 1 no epilogue math, 2 no pool, 4 no image stores,
                                     //   8 expanders skip the expansion, 16 no MMAs, 32 no TMEM loads
};

// bits j0 .. j0 + 63 of a bit-packed contact-map row (rw words, padding bits zero); j0 may be negative (> -64): the bits
// before the row start read as zero.  Used by the compact-axis adjacency product, where a protein's columns start at an
// arbitrary bit offset inside a 64-residue k-block.
__device__ __forceinline__ uint2 adj_row_window(const uint32_t *__restrict__ row, int rw, int j0)
{
    if (j0 >= 0) {
        const int wi = j0 >> 5, sh = j0 & 31;
        const uint32_t a0 = wi < rw ? __ldg(row + wi) : 0u, a1 = wi + 1 < rw ? __ldg(row + wi + 1) : 0u, a2 = wi + 2 < rw ? __ldg(row + wi + 2) : 0u;
        return make_uint2(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh));
    }
    const unsigned long long v = (((unsigned long long)__ldg(row + 1) << 32) | (unsigned long long)__ldg(row)) << (-j0);
    return make_uint2((uint32_t)v, (uint32_t)(v >> 32));
}

int launch_gemm_tc(mdf_ctx *ctx, int epi, int bn, int a_terms, int b_terms, const GemmArgs &args);

// CTA-pair (cta_group::2) variant for the regular weight GEMMs: 256 x 256 tiles, m_tiles / n_tiles in `args` count
// 256-row / 256-column tiles, a_bytes / b_bytes are the byte sizes of the term images (for the tensor maps).
int launch_gemm_pair(mdf_ctx *ctx, int epi, int a_terms, int b_terms, const GemmArgs &args, const size_t a_bytes[2], const size_t b_bytes[2]);


// flat [bytes/512][256] u16 tensor map whose [32 x 256] boxes are the 16 KiB operand tiles of an image
int make_tile_map(CUtensorMap *map, const void *base, size_t bytes);

}  // namespace tc
}  // namespace mdf
