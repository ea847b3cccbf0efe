// simple_env.cu -- SimpleEnvironment (3-node line, one agent) for B instances.
//
// Reference: src/env/simple_environment.py:289-315 (step: reward = score of the chosen
// neighbour of the start node, always done, packet returns to the start node),
// :217-233 (node obs = score), :235-246 (node-agent matrix), :248-287 (agent obs =
// [now] (+ flattened node adjacency + node obs for env_var != 1)).  The per-episode
// topology (_build_network, :106-187) is drawn on the host with the legacy MT19937 stream.
#include "common.cuh"

namespace gm {

__global__ void simple_step_kernel(int B, int env_var, const int* __restrict__ scores, const int* __restrict__ edges,
                                   const int* __restrict__ start_node, const int* __restrict__ start_edges,
                                   const int* __restrict__ actions, float* __restrict__ obs, float* __restrict__ node_obs,
                                   int8_t* __restrict__ node_agent, int8_t* __restrict__ node_adj,
                                   float* __restrict__ reward) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int* sc = scores + 3 * b;
    const int* ed = edges + 4 * b;
    int n0 = start_node[b];
    if (reward && actions) {
        int t = start_edges[2 * b + actions[b]];
        int dst = (ed[2 * t] == n0) ? ed[2 * t + 1] : ed[2 * t];
        reward[b] = (float)sc[dst];
    }
    int8_t adj[9];
    for (int i = 0; i < 9; i++) adj[i] = (i % 4 == 0);
    for (int k = 0; k < 2; k++) {
        int a = ed[2 * k], c = ed[2 * k + 1];
        adj[a * 3 + c] = 1; adj[c * 3 + a] = 1;
    }
    if (node_adj) for (int i = 0; i < 9; i++) node_adj[9 * b + i] = adj[i];
    if (node_obs) for (int i = 0; i < 3; i++) node_obs[3 * b + i] = (float)sc[i];
    if (node_agent) for (int i = 0; i < 3; i++) node_agent[3 * b + i] = (int8_t)(i == n0);
    if (obs) {
        int W = env_var == 1 ? 1 : 13;
        float* o = obs + (size_t)b * W;
        o[0] = (float)n0;
        if (env_var != 1) {
            for (int i = 0; i < 9; i++) o[1 + i] = (float)adj[i];
            for (int i = 0; i < 3; i++) o[10 + i] = (float)sc[i];
        }
    }
}

}  // namespace gm

extern "C" int gm_simple_step(int32_t B, int32_t env_var, const int32_t* scores, const int32_t* edges,
                              const int32_t* start_node, const int32_t* start_edges, const int32_t* actions, float* obs,
                              float* node_obs, int8_t* node_agent, int8_t* node_adj, float* reward, void* stream) {
    GM_CHECK_ARG(B > 0 && scores && edges && start_node && start_edges, "bad simple env arguments");
    GM_CHECK_ARG(env_var >= 1 && env_var <= 3, "env_var %d", env_var);
    gm::simple_step_kernel<<<gm::ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(B, env_var, scores, edges, start_node,
                                                                                start_edges, actions, obs, node_obs,
                                                                                node_agent, node_adj, reward);
    GM_LAUNCH_CHECK();
    return GM_OK;
}
