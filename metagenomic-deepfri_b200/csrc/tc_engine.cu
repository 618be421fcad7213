// Tensor-core engine — placeholder until gemm_tc.cu / lstm_tc.cu land.
#include "tc_engine.cuh"

namespace mdf {

int tc_model_init(mdf_model *, const mdf_model_desc *) { return MDF_OK; }
void tc_model_free(mdf_model *) {}
bool tc_available(const mdf_model *) { return false; }
size_t tc_workspace_bytes(const mdf_model *, int, int64_t) { return 0; }
int tc_forward(mdf_model *, mdf_batch *, int)
{
    set_error("tensor-core engine not built");
    return MDF_EUNSUPPORTED;
}

}  // namespace mdf
