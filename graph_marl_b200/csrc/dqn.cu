// dqn.cu -- DQN forward + epsilon-greedy action selection, device resident.
//
// Reference: src/model.py:199-203 (DQN.forward = MLP encoder with activation on every
// layer, then Linear q_net.fc), src/policy.py:20-51 (EpsilonGreedy.__call__):
//   q[mask] = -inf; random_actions = randint(n_act, size=A); random_filter = rand(A) < eps;
//   actions = argmax(q) * ~filter + filter * random_actions       (argmax = first maximum)
// The joint observation row [agent obs | graph obs] (wrapper.py:50) is consumed as two
// segments so the concatenation is never materialised.
#include <algorithm>

#include "common.cuh"
#include "gemm_sm100.cuh"
#include "linear_simt.cuh"

namespace gm {

int linear_dispatch(const LinearArgs& a, int math, void* ws, int64_t ws_bytes, cudaStream_t s);

constexpr int MAX_ACT = 8;

// one warp per agent row: q = h Wq^T + bq, mask, argmax, epsilon mix
__global__ void dqn_head_kernel(const float* __restrict__ h, int64_t ldh, int Hd, const float* __restrict__ qw,
                                const float* __restrict__ qb, int n_act, const uint8_t* __restrict__ mask, double eps,
                                const int* __restrict__ rand_action, const double* __restrict__ rand_u, uint64_t seed,
                                uint64_t step0, const uint64_t* __restrict__ step_dev, float* __restrict__ q_out,
                                int* __restrict__ act_out, int64_t rows) {
    const uint64_t step = step0 + (step_dev ? *step_dev : 0ull);
    int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (r >= rows) return;
    float acc[MAX_ACT];
#pragma unroll
    for (int a = 0; a < MAX_ACT; a++) acc[a] = 0.f;
    const float* hr = h + r * ldh;
    for (int k = lane; k < Hd; k += 32) {
        float x = hr[k];
#pragma unroll
        for (int a = 0; a < MAX_ACT; a++)
            if (a < n_act) acc[a] = fmaf(x, qw[(size_t)a * Hd + k], acc[a]);
    }
#pragma unroll
    for (int a = 0; a < MAX_ACT; a++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[a] += __shfl_xor_sync(FULL, acc[a], o);
    if (lane == 0) {
        int best = 0;
        float bestv = 0.f;
        for (int a = 0; a < n_act; a++) {
            float q = acc[a] + qb[a];
            if (q_out) q_out[r * n_act + a] = q;  // Q-values before masking (what the model returns)
            if (mask && mask[r * n_act + a]) q = -INFINITY;
            if (a == 0 || q > bestv) { best = a; bestv = q; }
        }
        int ra; double u;
        if (rand_action) { ra = rand_action[r]; u = rand_u[r]; }
        else {
            Philox p((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ 0x5bd1e995u, seed ^ 0xA5A5A5A5DEADBEEFull);
            ra = (int)__umulhi(p.r[0], (uint32_t)n_act);
            u = u53(p.r[1], p.r[2]);
        }
        act_out[r] = (u < eps) ? ra : best;
    }
}

static bool dqn_tc_math(int math) { return math == GM_MATH_BF16X3 || math == GM_MATH_BF16; }

// one activation buffer: fp32 [rows, maxw] or the same matrix tile-packed (rows rounded up to 128)
static int64_t dqn_buf_bytes(int64_t rows, int maxw) {
    return round_up(std::max<int64_t>(rows * maxw * 4, tc_pk_bytes(rows, (int)round_up(maxw, TC_BK))), 256);
}

// the last hidden layer carries the Q head in its epilogue when it fits one 256-column tile
static int dqn_layer_epi(const gm_dqn_params* p, int l) {
    const int U = p->units[p->n_layers - 1];
    return (l == p->n_layers - 1 && U <= 256 && (U & 3) == 0 && p->n_actions <= TC_MAX_ACT) ? EPI_QHEAD : EPI_LINEAR;
}

// packed tensor-core weights: layer l at off[l]; layer 0 is packed for input rows split as
// [split | in_features - split] so the joint observation is consumed without a concat
struct DqnPack {
    int64_t off[GM_MAX_LAYERS];
    int64_t total;
};

static DqnPack dqn_pack_layout(const gm_dqn_params* p, int split) {
    DqnPack L{};
    int64_t off = 0;
    int kin = p->in_features;
    for (int l = 0; l < p->n_layers; l++) {
        L.off[l] = off;
        int k0 = (l == 0 && split > 0) ? split : kin, k1 = (l == 0 && split > 0) ? kin - split : 0;
        off += tc_shape(p->units[l], k0, k1, dqn_layer_epi(p, l), 0).packed_bytes;
        kin = p->units[l];
    }
    L.total = off;
    return L;
}

static int dqn_pack(const gm_dqn_params* p, int split, void* out, cudaStream_t s) {
    DqnPack L = dqn_pack_layout(p, split);
    int kin = p->in_features, rc;
    for (int l = 0; l < p->n_layers; l++) {
        int k0 = (l == 0 && split > 0) ? split : kin, k1 = (l == 0 && split > 0) ? kin - split : 0;
        if ((rc = tc_pack_weights(p->w[l], kin, nullptr, 0, p->b[l], nullptr, p->units[l], k0, k1, dqn_layer_epi(p, l), 0,
                                  (char*)out + L.off[l], s)))
            return rc;
        kin = p->units[l];
    }
    return GM_OK;
}

}  // namespace gm

using namespace gm;

extern "C" {

int64_t gm_dqn_workspace_bytes(const gm_dqn_params* p, int64_t rows) {
    if (!p) return 0;
    int maxw = 0;
    for (int i = 0; i < p->n_layers; i++) maxw = max(maxw, p->units[i]);
    int64_t pack = dqn_tc_math(p->math) ? 2 * round_up(dqn_pack_layout(p, 8).total, 256) + 512 : 0;
    return 2 * dqn_buf_bytes(rows, maxw) + pack;
}

int64_t gm_dqn_packed_bytes(const gm_dqn_params* p, int32_t split) { return p ? dqn_pack_layout(p, split).total : 0; }

int gm_dqn_pack_weights(const gm_dqn_params* p, int32_t split, void* packed, int64_t packed_bytes, void* stream) {
    GM_CHECK_ARG(p && packed && ((uintptr_t)packed & 255) == 0, "packed buffer must be a 256-byte aligned device pointer");
    GM_CHECK_ARG(split >= 0 && split < p->in_features, "split %d outside [0,%d)", split, p->in_features);
    GM_CHECK_ARG(packed_bytes >= dqn_pack_layout(p, split).total, "packed buffer too small");
    return dqn_pack(p, split, packed, (cudaStream_t)stream);
}

int gm_dqn_act(const gm_dqn_params* p, int64_t rows, const float* obs_a, int32_t Da, int64_t lda, const float* obs_g,
               int32_t Dg, int64_t ldg, const void* obs_g_pk, const uint8_t* action_mask, double epsilon, const int32_t* rand_action,
               const double* rand_u, uint64_t philox_seed, uint64_t philox_step, const uint64_t* philox_step_dev, float* q_out,
               int32_t* act_out,
               void* workspace, int64_t workspace_bytes, void* stream) {
    GM_CHECK_ARG(p && obs_a && act_out && workspace, "null pointer");
    GM_CHECK_ARG(Dg == 0 || obs_g || obs_g_pk, "obs_g missing");
    GM_CHECK_ARG(p->n_layers >= 1 && p->n_layers <= GM_MAX_LAYERS, "n_layers");
    GM_CHECK_ARG(p->n_actions >= 1 && p->n_actions <= MAX_ACT, "n_actions %d > %d", p->n_actions, MAX_ACT);
    GM_CHECK_ARG(Da + Dg == p->in_features, "Da+Dg=%d != in_features %d", Da + Dg, p->in_features);
    GM_CHECK_ARG((rand_action == nullptr) == (rand_u == nullptr), "rand_action and rand_u must both be set or NULL");
    GM_CHECK_ARG(workspace_bytes >= gm_dqn_workspace_bytes(p, rows), "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    int maxw = 0;
    for (int i = 0; i < p->n_layers; i++) maxw = max(maxw, p->units[i]);
    GM_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    float* buf0 = (float*)workspace;
    float* buf1 = (float*)((char*)workspace + dqn_buf_bytes(rows, maxw));
    void* lin_ws = (char*)workspace + 2 * dqn_buf_bytes(rows, maxw);
    int64_t lin_ws_bytes = workspace_bytes - 2 * dqn_buf_bytes(rows, maxw);
    int rc;
    if (dqn_tc_math(p->math)) {
        // ---- tensor-core path: every layer one tcgen05 launch, layer 0 over both input segments ----
        const int split = Dg > 0 ? Da : 0;
        const char* packed = (const char*)p->packed;
        if (packed == nullptr || p->packed_split != split) {
            char* dst = (char*)round_up((int64_t)lin_ws, 256);
            if ((rc = dqn_pack(p, split, dst, s))) return rc;
            packed = dst;
        }
        DqnPack PL = dqn_pack_layout(p, split);
        // hidden activations stay tile-packed between layers (bulk-copied by the next layer); the last
        // hidden layer is written as fp32 rows for the Q head
        const float* x = obs_a;
        const uint8_t* xpk = nullptr;
        int64_t ldx = lda;
        int kin = p->in_features;
        for (int l = 0; l < p->n_layers; l++) {
            float* y = (l & 1) ? buf1 : buf0;
            const bool out_pk = l + 1 < p->n_layers && (p->units[l] % TC_BK) == 0;
            TcArgs a{};
            if (l == 0) {
                a.A0 = obs_a; a.lda0 = lda; a.K0 = Da;
                if (Dg > 0) {
                    if (obs_g_pk && (Dg % TC_BK) == 0) a.A1pk = (const uint8_t*)obs_g_pk;  // bulk-copied, no producer work
                    else { a.A1 = obs_g; a.lda1 = ldg; }
                    a.K1 = Dg;
                }
            } else {
                if (xpk) a.A0pk = xpk; else { a.A0 = x; a.lda0 = ldx; }
                a.K0 = kin;
            }
            a.Wp = (const uint8_t*)packed + PL.off[l];
            const int epi = dqn_layer_epi(p, l);
            if (epi == EPI_QHEAD) {
                a.q_w = p->q_w; a.q_b = p->q_b; a.n_act = p->n_actions;
                a.action_mask = action_mask; a.epsilon = epsilon; a.rand_action = rand_action; a.rand_u = rand_u;
                a.philox_seed = philox_seed; a.philox_step = philox_step; a.philox_step_dev = philox_step_dev;
                a.q_out = q_out; a.act_out = act_out;
            } else if (out_pk) a.Cpk = (uint8_t*)y; else { a.C = y; a.ldc = p->units[l]; }
            a.act = p->activation;
            a.M = rows; a.N = p->units[l];
            if ((rc = tc_launch(a, p->math, epi, s))) return rc;
            if (epi == EPI_QHEAD) return GM_OK;  // Q head, mask, argmax and epsilon mix ran in the epilogue
            xpk = out_pk ? (const uint8_t*)y : nullptr;
            x = y; ldx = p->units[l]; kin = p->units[l];
        }
        int Hd = p->units[p->n_layers - 1];
        dqn_head_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, s>>>(x, Hd, Hd, p->q_w, p->q_b, p->n_actions, action_mask, epsilon,
                                                                  rand_action, rand_u, philox_seed, philox_step, philox_step_dev,
                                                                  q_out, act_out, rows);
        GM_LAUNCH_CHECK();
        return GM_OK;
    }
    // layer 0 over the two input segments: y = act(obs_a W[:, :Da]^T + obs_g W[:, Da:]^T + b)
    {
        int U = p->units[0];
        int D = p->in_features;
        if (Dg > 0) {
            LinearArgs a{obs_a, lda, p->w[0], D, p->b[0], nullptr, buf0, U, rows, U, Da, -1, 0};
            if ((rc = linear_dispatch(a, p->math, lin_ws, lin_ws_bytes, s))) return rc;
            LinearArgs b{obs_g, ldg, p->w[0] + Da, D, nullptr, nullptr, buf0, U, rows, U, Dg, p->activation, 1};
            if ((rc = linear_dispatch(b, p->math, lin_ws, lin_ws_bytes, s))) return rc;
        } else {
            LinearArgs a{obs_a, lda, p->w[0], D, p->b[0], nullptr, buf0, U, rows, U, Da, p->activation, 0};
            if ((rc = linear_dispatch(a, p->math, lin_ws, lin_ws_bytes, s))) return rc;
        }
    }
    float* x = buf0;
    for (int l = 1; l < p->n_layers; l++) {
        float* y = (x == buf0) ? buf1 : buf0;
        LinearArgs a{x, p->units[l - 1], p->w[l], p->units[l - 1], p->b[l], nullptr, y, p->units[l], rows, p->units[l],
                     p->units[l - 1], p->activation, 0};
        if ((rc = linear_dispatch(a, p->math, lin_ws, lin_ws_bytes, s))) return rc;
        x = y;
    }
    int Hd = p->units[p->n_layers - 1];
    dqn_head_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, s>>>(x, Hd, Hd, p->q_w, p->q_b, p->n_actions, action_mask, epsilon,
                                                              rand_action, rand_u, philox_seed, philox_step, philox_step_dev, q_out, act_out,
                                                              rows);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

}  // extern "C"
