// gemm_sm100.cu -- tcgen05 / TMEM GEMM path (GM_MATH_BF16X3, GM_MATH_BF16).
// Placeholder until the tensor-core kernel lands: fails loudly, never substitutes.
#include "common.cuh"
#include "linear_simt.cuh"

namespace gm {
int64_t linear_tc_workspace_bytes(int64_t, int, int, int) { return 0; }
int linear_tc(const LinearArgs&, int math, void*, int64_t, cudaStream_t) {
    set_error("math mode %d (tcgen05) is not built in this revision", math);
    return GM_ERR_INVALID;
}
}  // namespace gm
