"""CPU rollout step assembled from the ORACLE pieces (C Routing env + numpy NetMon / DQN).

TEST / MEASUREMENT INFRASTRUCTURE ONLY: used by bench.py's `cpu_baseline` leg and by
`bench.py --impl reference` as the host-core baseline of the same rollout step the CUDA path
runs (src/main.py:673-737: EpsilonGreedy.__call__ -> NetMonWrapper.step -> ReplayBuffer.add).
Never imported by graph_marl_b200.

Parity status: PINNED through its parts (oracle.py / netmon_oracle.py are checked against
the reference's recorded outputs in tests/test_oracle_*.py); the vectorised aggregation used
here is checked against netmon_oracle.netmon_forward in tests/test_oracle_netmon.py.
"""
import numpy as np

from . import netmon_oracle as NO
from . import oracle as O

try:  # multi-threaded CPU tensor ops for the timed baseline (the reference itself computes with torch on CPU)
    import torch
    import torch.nn.functional as F
except Exception:  # pragma: no cover
    torch = None


def _t_mlp(x, w, prefix, on_output=True):
    n = 0
    while f"{prefix}linear_layers.{n}.weight" in w:
        n += 1
    for i in range(n):
        x = F.linear(x, w[f"{prefix}linear_layers.{i}.weight"], w[f"{prefix}linear_layers.{i}.bias"])
        if i < n - 1 or on_output:
            x = F.leaky_relu(x, 0.01)
    return x


def _t_cell(x, h, c, w, p, rnn):
    H = h.shape[-1]
    if rnn == "lstm":  # torch.nn.LSTMCell math (model.py:491,543)
        g = F.linear(x, w[p + "weight_ih"], w[p + "bias_ih"]) + F.linear(h, w[p + "weight_hh"], w[p + "bias_hh"])
        i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
        c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        return torch.sigmoid(o) * torch.tanh(c2), c2
    ig = F.layer_norm(F.linear(x, w[p + "weight_ih"]), (4 * H,), w[p + "ln_input.weight"], w[p + "ln_input.bias"], 1e-5)
    hg = F.layer_norm(F.linear(h, w[p + "weight_hh"]), (4 * H,), w[p + "ln_hidden.weight"], w[p + "ln_hidden.bias"], 1e-5)
    g = ig + hg + w[p + "bias_ih"]  # layernormlstm.py:24-42
    i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
    c2 = F.layer_norm(torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg), (H,), w[p + "ln_cell.weight"], w[p + "ln_cell.bias"], 1e-5)
    return torch.sigmoid(o) * torch.tanh(c2), c2


def netmon_step_shared_topology_torch(w, cfg, x, nbr_full, nbr_wo_self, state, agent_node):
    """netmon_step_shared_topology with torch CPU tensors (same math, all host threads).  `w` holds torch
    tensors, x / state / outputs are torch tensors, the index tables are int64 tensors."""
    B, N, _ = x.shape
    H, K, rnn = cfg["hidden"], cfg["iterations"], cfg["rnn_type"]
    if state is None:
        state = torch.zeros((B, N, 2 * H), dtype=torch.float32)
    st = state.reshape(B * N, 2, H)
    h = _t_mlp(x.reshape(B * N, -1), w, "encode.")
    h, c = _t_cell(h, st[:, 0], st[:, 1], w, "rnn_obs.", rnn)
    last = h
    for it in range(K):
        if it == K - 1:
            last = h
        hb = h.reshape(B, N, H)
        M = hb[:, nbr_full[:, 0]]
        for q in range(1, nbr_full.shape[1]):
            M = M + hb[:, nbr_full[:, q]]
        if cfg["agg_type"] == "mean":
            M = M / float(nbr_full.shape[1])
        h, c = _t_cell(M.reshape(B * N, H), h, c, w, "rnn_update.", rnn)
    new_state = torch.stack((h, c), 1).reshape(B, N, 2 * H)
    hb, lb = h.reshape(B, N, H), last.reshape(B, N, H)
    out = torch.cat([hb] + [lb[:, nbr_wo_self[:, q]] for q in range(nbr_wo_self.shape[1])], dim=-1)
    agent_out = torch.gather(out, 1, agent_node[:, :, None].expand(-1, -1, out.shape[-1]))
    return out, new_state, agent_out


def netmon_step_shared_topology(w, cfg, x, nbr_full, nbr_wo_self, state, agent_node):
    """netmon_oracle.netmon_forward for lstm/lnlstm + carryover + sum/mean + neighbour readout
    when every env shares one 3-regular topology: aggregation and readout as index gathers.
    nbr_full i64[N,4] (self + neighbours, ascending), nbr_wo_self i64[N,3]."""
    dtype = np.float32
    act = NO._act(cfg.get("activation", "leaky_relu"))
    B, N, _ = x.shape
    H, K = cfg["hidden"], cfg["iterations"]
    cell = NO.lstm_cell if cfg["rnn_type"] == "lstm" else NO.lnlstm_cell
    if state is None:
        state = np.zeros((B, N, 2 * H), dtype=dtype)
    st = state.reshape(B * N, 2, H)
    h = NO.mlp(x.reshape(B * N, -1), w, "encode.", act)
    h, c = cell(h, st[:, 0], st[:, 1], w, "rnn_obs.")
    last = h
    for it in range(K):
        if it == K - 1:
            last = h
        hb = h.reshape(B, N, H)
        M = hb[:, nbr_full[:, 0]]
        for q in range(1, nbr_full.shape[1]):  # ascending id order, same summation order as the loop form
            M = M + hb[:, nbr_full[:, q]]
        if cfg["agg_type"] == "mean":
            M = M / dtype(nbr_full.shape[1])
        h, c = cell(M.reshape(B * N, H), h, c, w, "rnn_update.")
    new_state = np.stack((h, c), 1).reshape(B, N, 2 * H)
    hb, lb = h.reshape(B, N, H), last.reshape(B, N, H)
    out = np.concatenate([hb] + [lb[:, nbr_wo_self[:, q]] for q in range(nbr_wo_self.shape[1])], axis=-1)
    agent_out = np.take_along_axis(out, agent_node[:, :, None].astype(np.int64), axis=1)
    return out, new_state, agent_out


class CpuRollout:
    """B independent envs on one shared topology, advanced by the oracle on the host cores."""

    def __init__(self, n_nodes, n_data, topo_seed, congestion, K, rnn, H, enc, dqn_units, num_envs, threads,
                 weights_netmon, weights_dqn, replay_capacity=0, seed=0, backend="torch"):
        self.backend = backend if torch is not None else "numpy"
        if self.backend == "torch":
            torch.set_num_threads(max(int(threads), 1))
        self.N, self.A, self.B = n_nodes, n_data, num_envs
        self.topo = O.generate_topology(n_nodes, seed=topo_seed)
        self.env = O.RoutingOracle(self.topo, n_data, enable_congestion=congestion, num_envs=num_envs, threads=threads)
        self.cfg = dict(hidden=H, iterations=K, rnn_type=rnn, rnn_carryover=True, agg_type="sum",
                        output_neighbor_hidden=True, output_global_hidden=False, activation="leaky_relu")
        self.w_nm = {k: np.asarray(v, np.float32) for k, v in weights_netmon.items()}
        self.w_dq = {k: np.asarray(v, np.float32) for k, v in weights_dqn.items()}
        full, wo = NO.adjacency_lists(self.topo["adj"])
        self.nbr_full = np.stack(full).astype(np.int64)
        self.nbr_wo = np.stack(wo).astype(np.int64)
        if self.backend == "torch":
            self.tw_nm = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in self.w_nm.items()}
            self.tw_dq = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in self.w_dq.items()}
            self.t_full, self.t_wo = torch.from_numpy(self.nbr_full), torch.from_numpy(self.nbr_wo)
        self.rng = np.random.default_rng(seed)
        self.state = None
        self.replay = None
        if replay_capacity:
            Dj = 6 * n_nodes + 10 + 4 * H
            self.replay = dict(obs=np.zeros((replay_capacity, n_data, Dj), np.float32),
                               next_obs=np.zeros((replay_capacity, n_data, Dj), np.float32),
                               node_obs=np.zeros((replay_capacity, n_nodes, 4 * n_nodes + 8), np.float32),
                               next_node_obs=np.zeros((replay_capacity, n_nodes, 4 * n_nodes + 8), np.float32),
                               node_state=np.zeros((replay_capacity, n_nodes, 2 * H), np.float32),
                               action=np.zeros((replay_capacity, n_data), np.int8),
                               reward=np.zeros((replay_capacity, n_data), np.float32))
            self.rindex = 0
            self.rcap = replay_capacity

    def _draws(self):
        B, A, N = self.B, self.A, self.N
        return (self.rng.integers(0, N, (B, A)).astype(np.int32), self.rng.integers(0, N, (B, A)).astype(np.int32),
                self.rng.random((B, A)))

    def _observe(self):
        o = self.env.observe(adj=True, node_agent=True)
        prev_state = self.state
        if self.backend == "torch":
            with torch.no_grad():
                _, self.state, g = netmon_step_shared_topology_torch(
                    self.tw_nm, self.cfg, torch.from_numpy(o["node_obs"]), self.t_full, self.t_wo, self.state,
                    torch.from_numpy(self.env.now.astype(np.int64)))
            self.joint = torch.cat([torch.from_numpy(o["obs"]), g], dim=-1)
        else:
            _, self.state, g = netmon_step_shared_topology(self.w_nm, self.cfg, o["node_obs"], self.nbr_full, self.nbr_wo,
                                                           self.state, self.env.now)
            self.joint = np.concatenate([o["obs"], g], axis=-1)
        self.node_obs = o["node_obs"]
        return prev_state

    def reset(self):
        self.state = None
        self.env.reset(*self._draws())
        self._observe()

    def step(self, epsilon=1.0):
        B, A = self.B, self.A
        if self.backend == "torch":
            with torch.no_grad():
                hq = _t_mlp(self.joint.reshape(B * A, -1), self.tw_dq, "encoder.")
                q = F.linear(hq, self.tw_dq["q_net.fc.weight"], self.tw_dq["q_net.fc.bias"]).reshape(B, A, -1).numpy()
        else:
            q = NO.dqn_forward(self.w_dq, self.joint)
        ra = self.rng.integers(0, 4, (B, A))
        ru = self.rng.random((B, A))
        act = NO.epsilon_greedy(q, epsilon, ra, ru)
        r = self.env.step(act.astype(np.int32), *self._draws())
        obs, node_obs = self.joint, self.node_obs
        last_state = self._observe()
        if self.replay is not None:
            idx = (self.rindex + np.arange(B)) % self.rcap
            rp = self.replay
            as_np = (lambda t: t.numpy()) if self.backend == "torch" else (lambda t: t)
            rp["obs"][idx], rp["next_obs"][idx] = as_np(obs), as_np(self.joint)
            rp["node_obs"][idx], rp["next_node_obs"][idx] = node_obs, self.node_obs
            rp["node_state"][idx] = 0 if last_state is None else as_np(last_state)
            rp["action"][idx], rp["reward"][idx] = act, r["reward"]
            self.rindex = int((self.rindex + B) % self.rcap)
        return r["reward"]
