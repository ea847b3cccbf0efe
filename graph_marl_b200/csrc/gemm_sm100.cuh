// gemm_sm100.cuh -- interface of the tcgen05 fused linear / LSTM-cell kernel (gemm_sm100.cu).
#pragma once
#include "common.cuh"

namespace gm {

enum { EPI_LINEAR = 0, EPI_LSTM = 1 };

struct TcArgs {
    // activations: logical row = [A0[m, 0:K0] | A1[m, 0:K1]] (A1 optional)
    const float* A0; int64_t lda0; int K0;
    const float* A1; int64_t lda1; int K1;
    int K0p, Kp;  // filled by tc_launch: start of segment 1 / total K in the packed K space
    // optional aggregation of segment 0 over adjacency lists: row m = (graph b, node v) reads
    // sum_q A0[b*nodes + nbr[list(b)][v][q]] for q < deg (model.py:213-229)
    const int* nbr; const int* deg; int DM; const int* list_index; int nodes; int mean;
    const uint8_t* Wp;        // weights (+ tile-ordered bias behind them) packed by tc_pack_weights
    const float* bias_tile;   // filled by tc_launch
    // EPI_LINEAR
    float* C; int64_t ldc; int act; int accumulate;
    int64_t M; int N;
    // EPI_LSTM: c' = sig(f)*c_in + sig(i)*tanh(g), h' = sig(o)*tanh(c'); optional second copy
    const float* c_in; int64_t ldc_in;
    float* h_out; int64_t ldh; float* c_out; int64_t ldco;
    float* h_out2; float* c_out2; int64_t ldh2;
    int H;
    int m_tiles, n_tiles;  // filled by tc_launch
};

struct TcShape {
    int BN, K0p, Kp, n_tiles;
    int64_t w_bytes, packed_bytes;  // weight tiles; weights + bias, rounded to 256
};

TcShape tc_shape(int N, int K0, int K1, int epi, int H);
// W fp32 [N, K0+K1] (row stride ldw), or W [N,K0] next to W1 [N,K1] -> packed bf16 hi/lo tiles
// (tc_shape(...).packed_bytes, 256-byte aligned destination)
// bias (+ bias2) are summed and stored tile-ordered behind the weights (either may be NULL)
int tc_pack_weights(const float* W, int64_t ldw, const float* W1, int64_t ldw1, const float* bias, const float* bias2, int N,
                    int K0, int K1, int epi, int H, void* out, cudaStream_t s);
int tc_launch(TcArgs a, int math, int epi, cudaStream_t s);

}  // namespace gm
