// gemm_dispatch.cu -- routes a fused linear layer to the arithmetic the caller asked for.
#include "common.cuh"
#include "linear_simt.cuh"

namespace gm {

int linear_tc(const LinearArgs& a, int math, void* ws, int64_t ws_bytes, cudaStream_t s);  // gemm_sm100.cu
int64_t linear_tc_workspace_bytes(int64_t M, int N, int K, int math);

int linear_dispatch(const LinearArgs& a, int math, void* ws, int64_t ws_bytes, cudaStream_t s) {
    if (math == GM_MATH_FP32) return launch_linear_simt(a, s);
    if (math == GM_MATH_BF16X3 || math == GM_MATH_BF16) return linear_tc(a, math, ws, ws_bytes, s);
    set_error("unknown math mode %d", math);
    return GM_ERR_INVALID;
}

}  // namespace gm

#include "gemm_sm100.cuh"

extern "C" {

int64_t gm_packed_activation_bytes(int64_t rows, int32_t width) {
    return gm::tc_pk_bytes(rows, (int)gm::round_up(width, gm::TC_BK));
}

int64_t gm_linear_workspace_bytes(int64_t M, int32_t N, int32_t K, int32_t math) {
    if (math == GM_MATH_FP32) return 0;
    return gm::linear_tc_workspace_bytes(M, N, K, math);
}

int gm_linear(const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc, int64_t M,
              int32_t N, int32_t K, int32_t activation, int32_t math, void* workspace, int64_t workspace_bytes,
              void* stream) {
    GM_CHECK_ARG(A && W && C && M >= 0 && N > 0 && K > 0, "bad linear args");
    gm::LinearArgs a{A, lda, W, K, bias, nullptr, C, ldc, M, N, K, activation, 0};
    return gm::linear_dispatch(a, math, workspace, workspace_bytes, (cudaStream_t)stream);
}
}
