// Persistent tcgen05 LSTM recurrence (lstm_tc.cu).
#pragma once
#include "mdf_common.cuh"

namespace mdf {
namespace tc {

size_t lstm_tc_smem_bytes(int H);
size_t lstm_tc_scratch_bytes(const mdf_ctx *ctx, int H);
// Rimg: [H/16][2][64 x H] fp16 operand images ((hi, lo) or, with `alternate`, (R_a, R_b)); tab: [H/16][26][16][4] fp32 (layer 1) or nullptr;
// pre: [Tp][4H] fp32 in [unit][gate] order (layers >= 2) or nullptr; scratch: lstm_tc_scratch_bytes.
int launch_lstm_tc(mdf_ctx *ctx, int H, int n, const __half *Rimg, const float *tab, const float *pre,
                   const uint8_t *idx_pad, const int *order, const int64_t *seq_off, const int64_t *seg_off,
                   __half *Himg, void *scratch, int alternate);

// streamed-weight kernel for large batches (lstm_stream.cu): Ra/Rb = [4H rows (cta, gate, unit) x H] images,
// tab = [26][H][4] fp32 (layer 1) or nullptr
bool lstm_stream_supported(int H);
size_t lstm_stream_scratch_bytes(const mdf_ctx *ctx, int H);
int launch_lstm_stream(mdf_ctx *ctx, int H, int n, const __half *Ra, const __half *Rb, const float *tab, const float *pre,
                       const uint8_t *idx_pad, const int *order, const int64_t *seq_off, const int64_t *seg_off,
                       __half *Himg, void *scratch);

// fused two-layer wavefront kernel (lstm_fused.cu): W = [R1, W2, R2][phase][4H x H] contiguous images with rows ordered
// (slice, gate, unit) for 64-unit slices, `phases` time-dither roundings each; tab = [26][H][4] fp32, b2 = [H][4] fp32
// ([unit][gate] order)
bool lstm_fused_supported(int H, int n_lstm);
size_t lstm_fused_scratch_bytes(const mdf_ctx *ctx, int H);
int launch_lstm_fused(mdf_ctx *ctx, int H, int n, const __half *W, int phases, const float *tab, const float *b2,
                      const uint8_t *idx_pad, const int *order, const int64_t *seq_off, const int64_t *seg_off,
                      __half *H1img, __half *H2img, void *scratch, const int *h_order, const int64_t *h_seq_off);

}  // namespace tc
}  // namespace mdf
