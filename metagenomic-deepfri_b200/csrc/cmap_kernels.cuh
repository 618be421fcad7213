// Internal launchers of cmap_kernels.cu (all on ctx->stream, device pointers only).
#pragma once
#include "mdf_common.cuh"

namespace mdf {

int launch_aln_transfer(mdf_ctx *ctx, int n, const char *q_aln, const char *t_aln, const int64_t *aln_off,
                        const int64_t *seq_off, const float *coords, const int64_t *coord_off, float4 *qc);
int launch_coords_to_frame(mdf_ctx *ctx, int64_t total, const float *coords, float4 *qc);
int launch_cmap_pair(mdf_ctx *ctx, int n, int nwork, const int2 *work, const float4 *qc, const int64_t *seq_off,
                     float thr2, int gen, int diag_val, uint32_t *packed, const int64_t *packed_off);
int launch_unpack_dense(mdf_ctx *ctx, int nwork, const int2 *work, const int64_t *seq_off, const uint32_t *packed,
                        const int64_t *packed_off, int32_t *dense, const int64_t *dense_off);
int launch_pairwise_sq(mdf_ctx *ctx, const float *X, int n, int m, float *D);
int launch_sparse_count(mdf_ctx *ctx, const uint32_t *packed, int L, int64_t *counts);
int launch_sparse_emit(mdf_ctx *ctx, const uint32_t *packed, int L, const int64_t *row_start, int32_t *pairs);
int launch_aln_t2q(mdf_ctx *ctx, const char *q_aln, const char *t_aln, int La, int *t2q, int *gapq, int *totals);
int launch_align_scatter(mdf_ctx *ctx, int Lq, int gen, const int *gapq, const int32_t *sparse, int64_t nnz,
                         const int *t2q, int nt, int32_t *out);

}  // namespace mdf
