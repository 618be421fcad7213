// Contact-map kernels (K1/K2): pairwise-distance threshold -> bit-packed maps, alignment
// transfer, and the dense/sparse variants behind the reference's low-level functions.
//
// Reference semantics (bit-exact):
//   mDeepFRI/contact_map_utils.pyx:17-37   pairwise_sqeuclidean  (unfused fp32, k order)
//   mDeepFRI/bio_utils.py:214-223          threshold (strict <, float32(thr**2)), argwhere
//   mDeepFRI/contact_map_utils.pyx:44-117  align_contact_map     (column walk, diag, generated
//                                                                 contacts, one-directional transfer)
//
// The fused path never materialises the target map: the alignment is scanned once into a
// query-frame coordinate table (target residue aligned to each query residue, NaN when there
// is none) and the distance test is evaluated directly on query index pairs.  Because the
// target->query map is injective on aligned residues, out[qi][qj] = 1 for a target contact
// (ti,tj) <=> dist(coords[q2t[qi]], coords[q2t[qj]]) < thr2, which is what the gather computes.
#include <stdlib.h>

#include "mdf_common.cuh"
#include "cmap_kernels.cuh"

namespace mdf {

__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    // ((0 + dx*dx) + dy*dy) + dz*dz, every operation rounded to fp32, no FMA contraction
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// -------------------------------------------------------------------------------------------
// K2: alignment scan + transfer into the query frame.  One block per protein.
//   column c:  q gap           -> target index advances (a '-/-' column counts here too)
//              q res, t gap    -> query residue has no target partner (flag bit0 = "gap": it
//                                 spawns generated contacts), query index advances
//              q res, t res    -> query residue maps to the current target index, both advance
// qc[seq_off[p] + qi] = (x, y, z, flag) of the mapped target residue; xyz = NaN when unmapped
// or when the target index lies beyond the structure's coordinate rows.
// -------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// block-wide exclusive scan of a packed pair of 16-bit counters; returns exclusive prefix and
// the block total through `total`
__device__ __forceinline__ int block_excl_scan(int v, int *total, int *smem /*>= 33 ints*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = warp_incl_scan(v, lane);
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (SCAN_THREADS / 32) ? smem[lane] : 0;
        int wi = warp_incl_scan(w, lane);
        smem[lane] = wi - w;
        if (lane == 31) smem[32] = wi;
    }
    __syncthreads();
    int excl = incl - v + smem[warp];
    *total = smem[32];
    __syncthreads();
    return excl;
}

__global__ void __launch_bounds__(SCAN_THREADS)
aln_transfer_kernel(int n, const char *__restrict__ q_aln, const char *__restrict__ t_aln,
                    const int64_t *__restrict__ aln_off, const int64_t *__restrict__ seq_off,
                    const float *__restrict__ coords, const int64_t *__restrict__ coord_off,
                    float4 *__restrict__ qc, int *__restrict__ err)
{
    __shared__ int sm[40];
    const int p = blockIdx.x;
    if (p >= n) return;
    const int64_t a0 = aln_off[p];
    const int La = (int)(aln_off[p + 1] - a0);
    const int64_t s0 = seq_off[p];
    const int Lq = (int)(seq_off[p + 1] - s0);
    const int64_t c0 = coord_off[p];
    const int nc = (int)(coord_off[p + 1] - c0);
    const float qnan = __int_as_float(0x7fc00000);
    int qbase = 0, tbase = 0;
    for (int base = 0; base < La; base += SCAN_THREADS) {
        const int c = base + threadIdx.x;
        int isq = 0, ist = 0, tres = 0;
        if (c < La) {
            isq = q_aln[a0 + c] != '-';
            tres = t_aln[a0 + c] != '-';
            ist = isq ? tres : 1;
        }
        int tot;
        int ex = block_excl_scan(isq | (ist << 16), &tot, sm);
        if (isq) {
            const int qi = qbase + (ex & 0xffff);
            const int ti = tbase + (ex >> 16);
            if (qi < Lq) {
                float4 v = make_float4(qnan, qnan, qnan, __int_as_float(tres ? 0 : 1));
                if (tres && ti < nc) {
                    const float *x = coords + (c0 + ti) * 3;
                    v.x = x[0]; v.y = x[1]; v.z = x[2];
                }
                qc[s0 + qi] = v;
            }
        }
        qbase += tot & 0xffff;
        tbase += tot >> 16;
    }
    if (threadIdx.x == 0 && qbase != Lq) atomicCAS(err, 0, MDF_DERR_LQ_MISMATCH);
}

// identity "alignment": the query frame is the structure itself (standalone contact maps)
__global__ void coords_to_frame_kernel(int64_t total, const float *__restrict__ coords, float4 *__restrict__ qc)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < total) qc[i] = make_float4(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2], 0.0f);
}

// -------------------------------------------------------------------------------------------
// K1: pairwise distance threshold on the query frame -> bit-packed rows.
// Block = (protein, 32-row block).  Row coordinates are staged in shared memory and broadcast;
// every lane keeps 4 columns (one per 32-bit word of a 128-column tile) in registers, so a row
// costs one LDS.128 + 4 x (8 FP32 ops + compare + ballot).  Diagonal and generated contacts are
// OR-ed in by the storing lane; padding columns compare against NaN and stay 0.
// -------------------------------------------------------------------------------------------
constexpr int PAIR_WARPS = 4;

__global__ void __launch_bounds__(PAIR_WARPS * 32)
cmap_pair_kernel(const int2 *__restrict__ work, const float4 *__restrict__ qc,
                 const int64_t *__restrict__ seq_off, float thr2, int gen, int diag_val,
                 uint32_t *__restrict__ packed, const int64_t *__restrict__ packed_off)
{
    __shared__ float4 rows[32];
    const int p = work[blockIdx.x].x, rb = work[blockIdx.x].y;
    const int64_t s0 = seq_off[p];
    const int L = (int)(seq_off[p + 1] - s0);
    const int rw = packed_row_words(L);
    const float4 *__restrict__ q = qc + s0;
    const float qnan = __int_as_float(0x7fc00000);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 32) {
        const int i = rb * 32 + threadIdx.x;
        rows[threadIdx.x] = i < L ? q[i] : make_float4(qnan, qnan, qnan, 0.f);
    }
    __syncthreads();
    uint32_t *__restrict__ out = packed + packed_off[p];
    const int ntile = (L + 127) >> 7;
    for (int item = warp; item < 4 * ntile; item += PAIR_WARPS) {
        const int tile = item >> 2, rs = item & 3;
        const int jlo = tile << 7;
        float cx[4], cy[4], cz[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = jlo + 32 * k + lane;
            float4 v = j < L ? q[j] : make_float4(qnan, qnan, qnan, 0.f);
            cx[k] = v.x; cy[k] = v.y; cz[k] = v.z;
        }
        // 8 rows x 128 columns per item.  The ballots give every lane the four words of a row; lane r keeps row r and
        // lanes 0..7 store their rows together afterwards (the diagonal / generated-contact band is OR-ed in by
        // cmap_band_kernel: handled here it cost a divergent lane-0 path on every row, more warp instructions than the
        // distance arithmetic itself)
        const int i0 = rb * 32 + rs * 8;
        uint32_t keep[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float4 a = rows[rs * 8 + r];                     // rows past L hold NaN -> all-zero words, never stored
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t w = __ballot_sync(0xffffffffu, sqdist3(a.x, a.y, a.z, cx[k], cy[k], cz[k]) < thr2);
                if (lane == r) keep[k] = w;
            }
        }
        if (lane < 8 && i0 + lane < L)
            *reinterpret_cast<uint4 *>(out + (size_t)(i0 + lane) * rw + tile * 4) = make_uint4(keep[0], keep[1], keep[2], keep[3]);
    }
}

// Per-lane constants of the 32 x 32 bit transpose across a warp: stage j (16, 8, 4, 2, 1) swaps the j x j sub-blocks across
// lanes l and l ^ j; a lane with bit j set takes (o >> j) & m, the other one (o << j) & ~m - both are a rotation of o (by 32 - j
// or j) under the complementary mask, so with the lane's keep mask and rotation held in registers a stage is one shuffle, one
// funnel shift and one LOP3.
struct Transpose32 {
    uint32_t keep[5], rot[5];
    __device__ __forceinline__ explicit Transpose32(int lane)
    {
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int j = 16 >> s;
            const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
            const bool up = (lane & j) != 0;
            // opaque moves: the optimiser would otherwise re-derive the masks from the lane bit with two extra LOP3 per stage
            asm volatile("mov.b32 %0, %1;" : "=r"(keep[s]) : "r"(up ? ~m : m));
            asm volatile("mov.b32 %0, %1;" : "=r"(rot[s]) : "r"(up ? 32 - j : j));
        }
    }
    __device__ __forceinline__ uint32_t operator()(uint32_t x) const
    {
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const uint32_t o = __shfl_xor_sync(0xffffffffu, x, 16 >> s);
            const uint32_t r = __funnelshift_l(o, o, rot[s]);
            x = (x & keep[s]) | (r & ~keep[s]);
        }
        return x;
    }
};

// a volatile shared-memory row load: the optimiser must not hoist the 32 (loop-invariant) rows out of the unit loop into
// 96 registers per lane
__device__ __forceinline__ float4 lds_row(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// Diagonal and generated contacts (contact_map_utils.pyx:82-97) of one 32 x 32 block in column orientation (lane = column c,
// bit r = row i0 + r; dg = (group of c) - (group of the rows) >= 0):  out[i][i] = diag_val ? 1 : computed;  out[i][c] = 1 for
// 0 < |c - i| <= gen when residue i or residue c was generated (a query residue facing a target gap).
__device__ __forceinline__ uint32_t band_bits(int dg, int lane, int gen, int diag_val, bool col_ok, bool col_gen,
                                              uint32_t row_gen /* bit r = row r generated */, uint32_t row_ok /* bit r = row < L */)
{
    const int ctr = 32 * dg + lane;                                          // c - i0: the row bit that faces column c
    const int lo = max(ctr - gen, 0), hi = min(ctr + gen, 31);
    uint32_t m = 0u;
    if (col_ok && lo <= hi) {
        m = (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
        const uint32_t dbit = dg == 0 ? 1u << lane : 0u;
        m &= ~dbit;
        m = col_gen ? (m & row_ok) : (m & row_gen);
        if (diag_val) m |= dbit & row_ok;
    }
    return m;
}

// Triangular form (the default).  One warp per (protein, 32-row block rb): the warp walks the 32-column groups rb .. nb-1 of its
// rows (exactly the upper triangle at 32 x 32 granularity; the diagonal block in full) in units of four groups that start AT its
// own group, the last unit group by group, and writes every block twice: as row words (word cb of its 32 rows, after a 32 x 32 bit
// transpose across the warp) and, right of the diagonal, transposed (word rb of the group's 32 rows).  Every word of the map is
// written exactly once: words >= group(row) by the row's own warp, words < group(row) by the transposed stores of warp (word),
// row padding words (nb .. rw-1) by the row's own warp.  Diagonal and generated contacts are OR-ed into the blocks next to the
// diagonal before both stores (no second pass over the map).  Single-warp CTAs: no warp of a CTA waits for another one's tail.
__global__ void __launch_bounds__(32, 32)
cmap_pair_tri_kernel(const int2 *__restrict__ work, const float4 *__restrict__ qc,
                     const int64_t *__restrict__ seq_off, float thr2, int gen, int diag_val,
                     uint32_t *__restrict__ packed, const int64_t *__restrict__ packed_off)
{
    __shared__ float4 rows[32];
    const int p = work[blockIdx.x].x, rb = work[blockIdx.x].y;
    const int64_t s0 = seq_off[p];
    const int L = (int)(seq_off[p + 1] - s0);
    const int rw = packed_row_words(L);
    const float4 *__restrict__ q = qc + s0;
    const float qnan = __int_as_float(0x7fc00000);
    const int lane = threadIdx.x;
    const int i = rb * 32 + lane;
    const bool rowok = i < L;
    const float4 mine = rowok ? q[i] : make_float4(qnan, qnan, qnan, 0.f);
    rows[lane] = mine;
    const uint32_t row_ok = __ballot_sync(0xffffffffu, rowok);
    const uint32_t row_gen = __ballot_sync(0xffffffffu, rowok && (__float_as_int(mine.w) & 1));
    __syncwarp();
    const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows);
    uint32_t *__restrict__ out = packed + packed_off[p];
    uint32_t *__restrict__ rowp = out + (size_t)i * rw;                      // lane r stores the words of row rb*32 + r
    const int nb = (L + 31) >> 5;
    const int band_groups = (31 + gen) >> 5;                                  // groups rb .. rb + band_groups hold diagonal / generated contacts
    const Transpose32 transpose(lane);
    // The row loop is unrolled by 8 only (a 5 KB loop body per unit instead of 22 KB of straight-line code: with 32 warps per SM
    // at different places of a fully unrolled unit, a quarter of all issue slots were lost to instruction-cache misses); each
    // 8-row chunk collects its bits with constant masks and is funnel-shifted into the word from the top.
    int cb = rb;
    for (; cb + 4 <= nb; cb += 4) {
        float cx[4], cy[4], cz[4];
        uint32_t tw[4] = {0u, 0u, 0u, 0u};                                    // bit r = contact(row rb*32 + r, column (cb+k)*32 + lane)
        uint32_t cgen = 0u;                                                   // bit k = column of group cb+k is a generated residue
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = (cb + k) * 32 + lane;
            const float4 v = c < L ? q[c] : make_float4(qnan, qnan, qnan, 0.f);
            cx[k] = v.x; cy[k] = v.y; cz[k] = v.z;
            cgen |= (uint32_t)(__float_as_int(v.w) & 1) << k;
        }
#pragma unroll 1
        for (int o = 0; o < 4; ++o) {
            float f8[4] = {0.f, 0.f, 0.f, 0.f};                              // the chunk's bits as an exact float sum (< 256)
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 a = lds_row(rows_s + 128 * o + 16 * r);          // rows past L hold NaN -> no contact
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (sqdist3(a.x, a.y, a.z, cx[k], cy[k], cz[k]) < thr2) f8[k] = __fadd_rn(f8[k], (float)(1 << r));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) tw[k] = __funnelshift_r(tw[k], __float2uint_rz(f8[k]), 8);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = (cb + k) * 32 + lane;
            if (cb + k - rb <= band_groups) tw[k] |= band_bits(cb + k - rb, lane, gen, diag_val, c < L, (cgen >> k) & 1u, row_gen, row_ok);
            if (cb + k > rb && c < L) out[(size_t)c * rw + rb] = tw[k];       // transposed copy: word rb of the group's rows
            const uint32_t rowword = transpose(tw[k]);
            if (rowok) rowp[cb + k] = rowword;
        }
    }
#pragma unroll 1
    for (; cb < nb; ++cb) {
        const int c = cb * 32 + lane;
        const float4 v = c < L ? q[c] : make_float4(qnan, qnan, qnan, 0.f);
        uint32_t t1 = 0u;
#pragma unroll 1
        for (int o = 0; o < 2; ++o) {
            uint32_t t16 = 0u;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const float4 a = lds_row(rows_s + 256 * o + 16 * r);
                if (sqdist3(a.x, a.y, a.z, v.x, v.y, v.z) < thr2) t16 |= 1u << r;
            }
            t1 = __funnelshift_r(t1, t16, 16);
        }
        if (cb - rb <= band_groups) t1 |= band_bits(cb - rb, lane, gen, diag_val, c < L, __float_as_int(v.w) & 1, row_gen, row_ok);
        if (cb > rb && c < L) out[(size_t)c * rw + rb] = t1;
        const uint32_t rowword = transpose(t1);
        if (rowok) rowp[cb] = rowword;
    }
    if (rowok)                                                                // row padding words behind the last group
        for (int w = nb; w < rw; ++w) rowp[w] = 0u;
}

// Diagonal and generated contacts (contact_map_utils.pyx:82-97): out[i][i] = diag_val ? 1 : computed; out[i][i +- d] = 1 for
// d = 1..gen when residue i or residue i +- d was generated (a query residue facing a target gap).  One thread per row, at
// most 2 gen + 1 bits: OR-ed into the words cmap_pair_kernel wrote.
__global__ void cmap_band_kernel(int n, const float4 *__restrict__ qc, const int64_t *__restrict__ seq_off, int gen, int diag_val,
                                 uint32_t *__restrict__ packed, const int64_t *__restrict__ packed_off)
{
    for (int p = blockIdx.x; p < n; p += gridDim.x) {
        const int64_t s0 = seq_off[p];
        const int L = (int)(seq_off[p + 1] - s0);
        const int rw = packed_row_words(L);
        const float4 *__restrict__ q = qc + s0;
        uint32_t *__restrict__ out = packed + packed_off[p];
        for (int i = threadIdx.x; i < L; i += blockDim.x) {
            const int gi = __float_as_int(q[i].w) & 1;
            uint32_t *row = out + (size_t)i * rw;
            for (int dj = -gen; dj <= gen; ++dj) {
                const int j = i + dj;
                if (j < 0 || j >= L) continue;
                const bool set = dj == 0 ? (diag_val != 0) : (gi || (__float_as_int(q[j].w) & 1));
                if (set) row[j >> 5] |= 1u << (j & 31);
            }
        }
    }
}


// packed -> dense int32 [L, L] (the reference's layout; bio_utils.py:220 / contact_map_utils.pyx:82).  HBM-bound by its 4 bytes
// per cell of output: one warp per row, 16-byte stores from the first 16-byte boundary of the row on (rows of an L x L int32
// block start at arbitrary 4-byte offsets), four bits per store cut out of two packed words with one funnel shift.
__global__ void __launch_bounds__(256) unpack_dense_kernel(const int2 *__restrict__ work, const int64_t *__restrict__ seq_off,
                                                           const uint32_t *__restrict__ packed, const int64_t *__restrict__ packed_off,
                                                           int32_t *__restrict__ dense, const int64_t *__restrict__ dense_off)
{
    const int p = work[blockIdx.x].x, rb = work[blockIdx.x].y;
    const int L = (int)(seq_off[p + 1] - seq_off[p]);
    const int rw = packed_row_words(L);
    const uint32_t *src = packed + packed_off[p];
    int32_t *dst = dense + dense_off[p];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < 32; r += 8) {
        const int i = rb * 32 + r;
        if (i >= L) break;
        const uint32_t *row = src + (size_t)i * rw;
        int32_t *out = dst + (size_t)i * L;
        int head = (int)(((16u - (unsigned)((uintptr_t)out & 15u)) & 15u) >> 2);
        if (head > L) head = L;
        if (lane < head) out[lane] = (int32_t)((row[0] >> lane) & 1u);
        const int nvec = (L - head) >> 2;
        int4 *ov = reinterpret_cast<int4 *>(out + head);
        for (int v = lane; v < nvec; v += 32) {
            const int j = head + 4 * v, w = j >> 5;
            const uint32_t bits = __funnelshift_r(row[w], row[min(w + 1, rw - 1)], j & 31);
            ov[v] = make_int4((int)(bits & 1u), (int)((bits >> 1) & 1u), (int)((bits >> 2) & 1u), (int)((bits >> 3) & 1u));
        }
        for (int j = head + 4 * nvec + lane; j < L; j += 32) out[j] = (int32_t)((row[j >> 5] >> (j & 31)) & 1u);
    }
}

// -------------------------------------------------------------------------------------------
// contact_map_utils.pyx:17-37, any m.  D[i][i] = 0 (never computed by the reference).
// -------------------------------------------------------------------------------------------
__global__ void pairwise_sq_kernel(const float *__restrict__ X, int n, int m, float *__restrict__ D)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n) return;
    float d = 0.0f;
    if (i != j) {
        const float *a = X + (size_t)min(i, j) * m, *b = X + (size_t)max(i, j) * m;
        for (int k = 0; k < m; ++k) {
            float diff = __fsub_rn(a[k], b[k]);
            d = __fadd_rn(d, __fmul_rn(diff, diff));
        }
    }
    D[(size_t)i * n + j] = d;
}

// -------------------------------------------------------------------------------------------
// sparse (argwhere) output: row popcounts -> exclusive scan -> ordered emit
// -------------------------------------------------------------------------------------------
__global__ void row_popcount_kernel(const uint32_t *__restrict__ packed, int L, int rw, int64_t *__restrict__ counts)
{
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= L) return;
    int c = 0;
    for (int w = lane; w < rw; w += 32) c += __popc(packed[(size_t)i * rw + w]);
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) counts[i] = c;
}

// single block: exclusive scan of `counts[0..n)` in place, total to counts[n]
__global__ void excl_scan64_kernel(int64_t *counts, int n)
{
    __shared__ int64_t carry;
    __shared__ int64_t wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        int64_t v = i < n ? counts[i] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int64_t woff = 0;
        for (int w = 0; w < warp; ++w) woff += wsum[w];
        int64_t tot = 0;
        for (int w = 0; w < nw; ++w) tot += wsum[w];
        if (i < n) counts[i] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[n] = carry;
}

__global__ void emit_pairs_kernel(const uint32_t *__restrict__ packed, int L, int rw,
                                  const int64_t *__restrict__ row_start, int32_t *__restrict__ pairs)
{
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= L) return;
    int64_t base = row_start[i];
    for (int w0 = 0; w0 < rw; w0 += 32) {
        const int w = w0 + lane;
        uint32_t bits = w < rw ? packed[(size_t)i * rw + w] : 0u;
        int c = __popc(bits), incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int64_t pos = base + incl - c;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            pairs[2 * pos] = i;
            pairs[2 * pos + 1] = (w << 5) + b;
            ++pos;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// -------------------------------------------------------------------------------------------
// contact_map_utils.pyx:44-117 with an explicit sparse target map (drop-in align_contact_map).
// Step 1: scan the alignment -> t2q[] (target index -> query index or -1), gapq[] flags.
// Step 2: diagonal + generated contacts.  Step 3: scatter target contacts (idempotent stores,
// same benign race as the reference's prange, :105-115).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS)
aln_t2q_kernel(const char *__restrict__ q_aln, const char *__restrict__ t_aln, int La,
               int *__restrict__ t2q, int *__restrict__ gapq, int *__restrict__ totals /*[2] = Lq, nt*/)
{
    __shared__ int sm[40];
    int qbase = 0, tbase = 0;
    for (int base = 0; base < La; base += SCAN_THREADS) {
        const int c = base + threadIdx.x;
        int isq = 0, ist = 0, tres = 0;
        if (c < La) {
            isq = q_aln[c] != '-';
            tres = t_aln[c] != '-';
            ist = isq ? tres : 1;
        }
        int tot;
        int ex = block_excl_scan(isq | (ist << 16), &tot, sm);
        const int qi = qbase + (ex & 0xffff), ti = tbase + (ex >> 16);
        if (c < La) {
            if (ist) t2q[ti] = isq ? qi : -1;
            if (isq) gapq[qi] = tres ? 0 : 1;
        }
        qbase += tot & 0xffff;
        tbase += tot >> 16;
    }
    if (threadIdx.x == 0) { totals[0] = qbase; totals[1] = tbase; }
}

__global__ void align_diag_gen_kernel(int Lq, int gen, const int *__restrict__ gapq, int32_t *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Lq) return;
    out[(size_t)i * Lq + i] = 1;
    if (gapq[i]) {
        for (int j = 1; j <= gen; ++j) {
            if (i + j < Lq) { out[(size_t)(i + j) * Lq + i] = 1; out[(size_t)i * Lq + i + j] = 1; }
            if (i - j >= 0) { out[(size_t)(i - j) * Lq + i] = 1; out[(size_t)i * Lq + i - j] = 1; }
        }
    }
}

__global__ void align_scatter_kernel(const int32_t *__restrict__ sparse, int64_t nnz, const int *__restrict__ t2q,
                                     int nt, int Lq, int32_t *__restrict__ out)
{
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nnz) return;
    const int ti = sparse[2 * r], tj = sparse[2 * r + 1];
    if ((unsigned)ti < (unsigned)nt && (unsigned)tj < (unsigned)nt) {   // negatives fail, like size_t compare
        const int a = t2q[ti], b = t2q[tj];
        if (a != -1 && b != -1) out[(size_t)a * Lq + b] = 1;
    }
}

// ------------------------------------------------------------------------------------------- host launchers
int launch_aln_transfer(mdf_ctx *ctx, int n, const char *q_aln, const char *t_aln, const int64_t *aln_off,
                        const int64_t *seq_off, const float *coords, const int64_t *coord_off, float4 *qc)
{
    if (n <= 0) return MDF_OK;
    aln_transfer_kernel<<<n, SCAN_THREADS, 0, ctx->stream>>>(n, q_aln, t_aln, aln_off, seq_off, coords,
                                                              coord_off, qc, ctx->d_err);
    MDF_LAUNCH_CHECK(ctx);
    return MDF_OK;
}

int launch_coords_to_frame(mdf_ctx *ctx, int64_t total, const float *coords, float4 *qc)
{
    if (total <= 0) return MDF_OK;
    coords_to_frame_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, ctx->stream>>>(total, coords, qc);
    MDF_LAUNCH_CHECK(ctx);
    return MDF_OK;
}

int launch_cmap_pair(mdf_ctx *ctx, int n, int nwork, const int2 *work, const float4 *qc, const int64_t *seq_off,
                     float thr2, int gen, int diag_val, uint32_t *packed, const int64_t *packed_off)
{
    if (nwork <= 0) return MDF_OK;
    static const bool full_square = getenv("MDF_CMAP_SYM") && atoi(getenv("MDF_CMAP_SYM")) == 0;   // A/B switch: evaluate both triangles
    if (!full_square) {
        cmap_pair_tri_kernel<<<nwork, 32, 0, ctx->stream>>>(work, qc, seq_off, thr2, gen < 0 ? 0 : gen, diag_val, packed, packed_off);
        MDF_LAUNCH_CHECK(ctx);
        return MDF_OK;
    }
    cmap_pair_kernel<<<nwork, PAIR_WARPS * 32, 0, ctx->stream>>>(work, qc, seq_off, thr2, gen < 0 ? 0 : gen, diag_val, packed, packed_off);
    MDF_LAUNCH_CHECK(ctx);
    if (n > 0 && (diag_val || gen > 0)) {
        cmap_band_kernel<<<std::min(n, 16 * ctx->sm_count), 128, 0, ctx->stream>>>(n, qc, seq_off, gen < 0 ? 0 : gen, diag_val, packed, packed_off);
        MDF_LAUNCH_CHECK(ctx);
    }
    return MDF_OK;
}

int launch_unpack_dense(mdf_ctx *ctx, int nwork, const int2 *work, const int64_t *seq_off, const uint32_t *packed,
                        const int64_t *packed_off, int32_t *dense, const int64_t *dense_off)
{
    if (nwork <= 0) return MDF_OK;
    unpack_dense_kernel<<<nwork, 256, 0, ctx->stream>>>(work, seq_off, packed, packed_off, dense, dense_off);
    MDF_LAUNCH_CHECK(ctx);
    return MDF_OK;
}

int launch_pairwise_sq(mdf_ctx *ctx, const float *X, int n, int m, float *D)
{
    if (n <= 0) return MDF_OK;
    dim3 grid(cdiv(n, 256), n);
    pairwise_sq_kernel<<<grid, 256, 0, ctx->stream>>>(X, n, m, D);
    MDF_LAUNCH_CHECK(ctx);
    return MDF_OK;
}

int launch_sparse_count(mdf_ctx *ctx, const uint32_t *packed, int L, int64_t *counts /*[L+1]*/)
{
    const int rw = packed_row_words(L);
    if (L > 0) {
        row_popcount_kernel<<<cdiv(L, 8), 256, 0, ctx->stream>>>(packed, L, rw, counts);
        MDF_LAUNCH_CHECK(ctx);
    }
    excl_scan64_kernel<<<1, 1024, 0, ctx->stream>>>(counts, L);
    MDF_LAUNCH_CHECK(ctx);
    return MDF_OK;
}

int launch_sparse_emit(mdf_ctx *ctx, const uint32_t *packed, int L, const int64_t *row_start, int32_t *pairs)
{
    if (L <= 0) return MDF_OK;
    emit_pairs_kernel<<<cdiv(L, 8), 256, 0, ctx->stream>>>(packed, L, packed_row_words(L), row_start, pairs);
    MDF_LAUNCH_CHECK(ctx);
    return MDF_OK;
}

int launch_aln_t2q(mdf_ctx *ctx, const char *q_aln, const char *t_aln, int La, int *t2q, int *gapq, int *totals)
{
    aln_t2q_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(q_aln, t_aln, La, t2q, gapq, totals);
    MDF_LAUNCH_CHECK(ctx);
    return MDF_OK;
}

int launch_align_scatter(mdf_ctx *ctx, int Lq, int gen, const int *gapq, const int32_t *sparse, int64_t nnz,
                         const int *t2q, int nt, int32_t *out)
{
    if (Lq <= 0) return MDF_OK;
    align_diag_gen_kernel<<<cdiv(Lq, 256), 256, 0, ctx->stream>>>(Lq, gen < 0 ? 0 : gen, gapq, out);
    MDF_LAUNCH_CHECK(ctx);
    if (nnz > 0) {
        align_scatter_kernel<<<(unsigned)cdiv64(nnz, 256), 256, 0, ctx->stream>>>(sparse, nnz, t2q, nt, Lq, out);
        MDF_LAUNCH_CHECK(ctx);
    }
    return MDF_OK;
}

}  // namespace mdf
