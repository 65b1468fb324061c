// gemm_tc.cu — tensor-core (tcgen05, 3xTF32) variants of the dense feature transforms.
// Entry points return -1 when a shape is not supported so gemm.cu falls back to the FP32 FMA kernel.
#include "common.cuh"

namespace gnn {

int gemm_tc_nt(gnn_ctx *, int64_t, int32_t, int32_t, const float *, int64_t, const float *, int64_t, float *, int64_t,
               const float *, int) {
    return -1;
}
int gemm_tc_nn(gnn_ctx *, int64_t, int32_t, int32_t, const float *, int64_t, const float *, int64_t, float *, int64_t,
               const float *, int64_t) {
    return -1;
}
int gemm_tc_tn(gnn_ctx *, int64_t, int32_t, int32_t, const float *, int64_t, const float *, int64_t, float *,
               int64_t) {
    return -1;
}

} // namespace gnn
