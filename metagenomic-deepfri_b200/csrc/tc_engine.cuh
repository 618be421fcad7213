// Tensor-core (tcgen05) engine of the GCN half of the path (engine 1).
#pragma once
#include <functional>

#include "gcn.cuh"

namespace mdf {

int tc_model_init(mdf_model *m, const mdf_model_desc *d);   // builds fp16 hi/lo operand images
void tc_model_free(mdf_model *m);
bool tc_available(const mdf_model *m);
size_t tc_workspace_bytes(const mdf_model *m, int n, const int64_t *seq_off);
void tc_batch_free(mdf_batch *b);
// `before_graphconv` (optional) runs on the host right before the first kernel that needs the contact maps / degrees is enqueued:
// the path uses it to build the maps AFTER the LSTM language model and the embedding have been enqueued (they only need sequences)
int tc_forward(mdf_model *m, mdf_batch *b, int upto, const std::function<int()> *before_graphconv = nullptr);

// pieces of the head shared with the CNN branch (cnn_tc.cu)
int tc_dense_split(mdf_ctx *ctx, int n, const float *src, int K, const __half *const W[2], int N, int ldc, const float *bias,
                   int act, float *out);
int tc_softmax0_strided(mdf_ctx *ctx, int n, int C, int ld, const float *logits, float *scores);
void tc_build_split_weight_images(const float *src_kn, int rows, int K, std::vector<__half> &hi, std::vector<__half> &lo);

}  // namespace mdf
