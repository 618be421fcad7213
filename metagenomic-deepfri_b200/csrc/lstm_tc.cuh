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

}  // namespace tc
}  // namespace mdf
