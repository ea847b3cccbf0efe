// linear_simt.cuh -- fp32 CUDA-core GEMM with fused bias / accumulate / activation epilogue.
//
//   C[M,N] (ldc) (+)= act( A[M,K] (lda) * W[N,K]^T (ldw) + bias[N] )
//
// This is the GM_MATH_FP32 arithmetic: every product and sum is an fp32 FFMA, so results
// match the reference's torch fp32 Linear / LSTMCell to summation-order noise.  It is the
// accuracy baseline for the tcgen05 path (gemm_sm100.cu) and the only path used for the
// ill-conditioned LayerNormLSTM gates.
#pragma once
#include "common.cuh"

namespace gm {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 8, SG_THREADS = 256;

struct LinearArgs {
    const float* A; int64_t lda;
    const float* W; int64_t ldw;
    const float* bias;   // may be null
    const float* bias2;  // may be null (second bias vector, e.g. b_hh)
    float* C; int64_t ldc;
    int64_t M; int N, K;
    int act;         // GM_ACT_* or -1 for identity
    int accumulate;  // C += instead of C =
};

int launch_linear_simt(const LinearArgs& a, cudaStream_t s);

}  // namespace gm
