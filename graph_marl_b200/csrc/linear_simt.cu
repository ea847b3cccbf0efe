// linear_simt.cu -- the fp32 CUDA-core GEMM kernel declared in linear_simt.cuh.
#include "linear_simt.cuh"

namespace gm {

__global__ void __launch_bounds__(SG_THREADS) linear_simt_kernel(LinearArgs p) {
    __shared__ __align__(16) float As[2][SG_BK][SG_BM];
    __shared__ __align__(16) float Ws[2][SG_BK][SG_BN];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.y * SG_BM;
    const int n0 = blockIdx.x * SG_BN;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 8 x 8 outputs each
    // loader mapping: 128 rows x 8 k per tile = 1024 elements, 4 per thread (one row, 4 k)
    const int lr = tid >> 1, lk = (tid & 1) * 4;
    const bool vecA = ((p.lda & 3) == 0) && (((uintptr_t)p.A & 15) == 0);
    const bool vecW = ((p.ldw & 3) == 0) && (((uintptr_t)p.W & 15) == 0);

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

    auto load_tile = [&](int k0, float (&ra)[4], float (&rw)[4]) {
        int64_t m = m0 + lr;
        int n = n0 + lr;
        int k = k0 + lk;
        if (m < p.M && vecA && k + 3 < p.K) {
            float4 v = *(const float4*)(p.A + m * p.lda + k);
            ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) ra[q] = (m < p.M && k + q < p.K) ? p.A[m * p.lda + k + q] : 0.f;
        }
        if (n < p.N && vecW && k + 3 < p.K) {
            float4 v = *(const float4*)(p.W + (int64_t)n * p.ldw + k);
            rw[0] = v.x; rw[1] = v.y; rw[2] = v.z; rw[3] = v.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) rw[q] = (n < p.N && k + q < p.K) ? p.W[(int64_t)n * p.ldw + k + q] : 0.f;
        }
    };
    auto store_tile = [&](int buf, const float (&ra)[4], const float (&rw)[4]) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            As[buf][lk + q][lr] = ra[q];
            Ws[buf][lk + q][lr] = rw[q];
        }
    };

    float ra[4], rw[4];
    load_tile(0, ra, rw);
    store_tile(0, ra, rw);
    __syncthreads();
    const int nk = (p.K + SG_BK - 1) / SG_BK;
    for (int kt = 0; kt < nk; kt++) {
        int buf = kt & 1;
        if (kt + 1 < nk) load_tile((kt + 1) * SG_BK, ra, rw);
#pragma unroll
        for (int k = 0; k < SG_BK; k++) {
            float a[8], b[8];
            float4 a0 = *(const float4*)&As[buf][k][ty * 4];
            float4 a1 = *(const float4*)&As[buf][k][64 + ty * 4];
            float4 b0 = *(const float4*)&Ws[buf][k][tx * 4];
            float4 b1 = *(const float4*)&Ws[buf][k][64 + tx * 4];
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tile(buf ^ 1, ra, rw);
            __syncthreads();
        }
    }
    // epilogue: rows {ty*4+i, 64+ty*4+i}, cols {tx*4+j, 64+tx*4+j}
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= p.M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; jh++) {
            int n = n0 + jh * 64 + tx * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float x = acc[i][jh * 4 + j];
                int nn = n + j;
                if (nn < p.N) {
                    if (p.bias) x += p.bias[nn];
                    if (p.bias2) x += p.bias2[nn];
                    if (p.accumulate) x += p.C[m * p.ldc + nn];
                    if (p.act >= 0) x = apply_act(x, p.act);
                }
                v[j] = x;
            }
            float* c = p.C + m * p.ldc + n;
            if (n + 3 < p.N && ((p.ldc & 3) == 0) && (((uintptr_t)p.C & 15) == 0)) {
                *(float4*)c = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (n + j < p.N) c[j] = v[j];
            }
        }
    }
}

int launch_linear_simt(const LinearArgs& a, cudaStream_t s) {
    if (a.M == 0) return GM_OK;
    dim3 grid(ceil_div(a.N, SG_BN), (unsigned)((a.M + SG_BM - 1) / SG_BM));
    linear_simt_kernel<<<grid, SG_THREADS, 0, s>>>(a);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

}  // namespace gm
