"""CPU ORACLE (numpy) for the floating-point part of the rollout hot path:
NetMon forward, DQN forward, epsilon-greedy selection and replay index sampling.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  A gather-based restatement of the
reference's formulas -- explicit gate math and sorted neighbour lists instead of
the dense bmm / nonzero machinery -- with every function citing the reference
file:line it follows (paths relative to /root/reference/src).  The arithmetic the
reference delegates to torch (Linear, LSTMCell, GRUCell, LayerNorm; torch>=2.0
unpinned, 2.11.0 here) is restated from torch's documented formulas.

Parity status: PINNED against tests/golden/netmon.npz, dqn_policy.npz,
replay.npz, wrapper.npz (outputs of the unmodified reference on torch CPU fp32).
`dtype=np.float64` gives a higher-precision truth for tolerance studies.
"""
import numpy as np


def _act(name):
    if name == "leaky_relu":
        return lambda x: np.where(x >= 0, x, x * x.dtype.type(0.01))
    if name == "relu":
        return lambda x: np.maximum(x, 0)
    if name == "tanh":
        return np.tanh
    if name == "sigmoid":
        return _sigmoid
    if name == "elu":
        return lambda x: np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
    raise ValueError(name)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def mlp(x, w, prefix, act, activation_on_output=True):
    """model.py:32-42 (MLP.forward); parameters `{prefix}linear_layers.{i}.weight/bias`."""
    n = 0
    while f"{prefix}linear_layers.{n}.weight" in w:
        n += 1
    for i in range(n):
        x = x @ w[f"{prefix}linear_layers.{i}.weight"].T + w[f"{prefix}linear_layers.{i}.bias"]
        if i < n - 1 or activation_on_output:
            x = act(x)
    return x


def lstm_cell(x, h, c, w, p):
    """torch.nn.LSTMCell as called at model.py:491,543; gate order i,f,g,o."""
    g = x @ w[p + "weight_ih"].T + w[p + "bias_ih"] + h @ w[p + "weight_hh"].T + w[p + "bias_hh"]
    H = h.shape[-1]
    i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
    c2 = _sigmoid(f) * c + _sigmoid(i) * np.tanh(gg)
    return _sigmoid(o) * np.tanh(c2), c2


def _layer_norm(x, weight, bias, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * weight + bias


def lnlstm_cell(x, h, c, w, p):
    """layernormlstm.py:24-42."""
    ig = _layer_norm(x @ w[p + "weight_ih"].T, w[p + "ln_input.weight"], w[p + "ln_input.bias"])
    hg = _layer_norm(h @ w[p + "weight_hh"].T, w[p + "ln_hidden.weight"], w[p + "ln_hidden.bias"])
    g = ig + hg + w[p + "bias_ih"]
    H = h.shape[-1]
    i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
    c2 = _layer_norm(_sigmoid(f) * c + _sigmoid(i) * np.tanh(gg), w[p + "ln_cell.weight"],
                     w[p + "ln_cell.bias"])
    return _sigmoid(o) * np.tanh(c2), c2


def gru_cell(x, h, w, p):
    """torch.nn.GRUCell as called at model.py:494,551; gate order r,z,n."""
    gi = x @ w[p + "weight_ih"].T + w[p + "bias_ih"]
    gh = h @ w[p + "weight_hh"].T + w[p + "bias_hh"]
    H = h.shape[-1]
    r = _sigmoid(gi[:, :H] + gh[:, :H])
    z = _sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1 - z) * n + z * h


def adjacency_lists(mask):
    """Per node: sorted ids with mask != 0 (incl. self if present), model.py:213-229,
    and the sorted neighbour ids without self, model.py:597-614."""
    N = mask.shape[0]
    full = [np.nonzero(mask[v])[0] for v in range(N)]
    nbrs = [f[f != v] for v, f in enumerate(full)]
    return full, nbrs


def netmon_forward(w, cfg, x, mask, state=None, agent_node=None, node_agent=None,
                   max_degree=None, dtype=np.float32):
    """model.py:451-631.

    w: dict of numpy params with the reference's state_dict keys.
    cfg: dict(hidden, iterations, rnn_type, rnn_carryover, agg_type,
              output_neighbor_hidden, output_global_hidden, activation).
    x [B,N,Dn]; mask [B,N,N]; state [B,N,S] or None.
    Returns (node_out [B,N,O], new_state [B,N,S], agent_out [B,A,O] or None).
    """
    w = {k: np.asarray(v, dtype=dtype) for k, v in w.items()}
    act = _act(cfg.get("activation", "leaky_relu"))
    B, N, _ = x.shape
    H, K = cfg["hidden"], cfg["iterations"]
    rnn, carry = cfg["rnn_type"], cfg.get("rnn_carryover", True)
    ns = {"lstm": 2 if carry else 4, "lnlstm": 2 if carry else 4, "gru": 1 if carry else 2,
          "none": 1}[rnn]
    if state is None:
        state = np.zeros((B, N, ns * H), dtype=dtype)  # :480-484
    state = np.asarray(state, dtype=dtype).reshape(B * N, ns, H)  # :417-434
    xs = np.asarray(x, dtype=dtype).reshape(B * N, -1)
    h = mlp(xs, w, "encode.", act)  # :489
    cell = lstm_cell if rnn == "lstm" else lnlstm_cell
    if rnn in ("lstm", "lnlstm"):
        h0, c0 = cell(h, state[:, 0], state[:, 1], w, "rnn_obs.")  # :491
        h, c = h0, c0
    elif rnn == "gru":
        h0 = gru_cell(h, state[:, 0], w, "rnn_obs.")
        h = h0
    lists = [adjacency_lists(np.asarray(mask[b])) for b in range(B)]
    last = np.zeros_like(h) if K <= 0 else None  # :498-499
    h1 = c1 = None
    for it in range(K):  # :509
        if it == K - 1:
            last = h  # :510-519 (value before this iteration's aggregation)
        hb = h.reshape(B, N, H)
        M = np.zeros_like(hb)
        for b in range(B):
            full = lists[b][0]
            for v in range(N):
                s = hb[b, full[v]].sum(axis=0) if len(full[v]) else 0.0
                if cfg["agg_type"] == "mean":
                    s = s / dtype(max(len(full[v]), 1))  # :227-229
                M[b, v] = s
        M = M.reshape(B * N, H)
        if rnn in ("lstm", "lnlstm"):
            hi, ci = (state[:, 2], state[:, 3]) if (not carry and it == 0) else (h, c)  # :538-541
            h1, c1 = cell(M, hi, ci, w, "rnn_update.")
            h, c = h1, c1
        elif rnn == "gru":
            hi = state[:, 1] if (not carry and it == 0) else h
            h1 = gru_cell(M, hi, w, "rnn_update.")
            h = h1
        else:
            h = M
    if rnn in ("lstm", "lnlstm"):  # :562-576
        new_state = np.stack((h1, c1), 1) if carry else np.stack((h0, c0, h1, c1), 1)
    elif rnn == "gru":
        # no-carryover GRU: the reference stacks [2,1,R,H] and reshapes WITHOUT the
        # per-row interleave (:571,:449) -> rows are scrambled; reproduced, not fixed.
        new_state = h1[:, None] if carry else np.stack((h0, h1), 0)
    else:
        new_state = h[:, None]
    new_state = new_state.reshape(B, N, -1)
    hb = h.reshape(B, N, H)
    parts = [hb]
    if cfg.get("output_global_hidden", False):  # :624-627
        parts.append(np.repeat(hb.mean(axis=1, keepdims=True), N, axis=1))
    if cfg.get("output_neighbor_hidden", False):  # :582-622
        lb = last.reshape(B, N, H)
        if max_degree is None:
            max_degree = int(max(np.asarray(mask).sum(axis=-1).max() - 1, 0))
        nb = np.zeros((B, N, max_degree, H), dtype=dtype)
        for b in range(B):
            for v in range(N):
                for k_, u in enumerate(lists[b][1][v][:max_degree]):
                    nb[b, v, k_] = lb[b, u]
        parts.append(nb.reshape(B, N, -1))
    out = np.concatenate(parts, axis=-1)
    agent_out = None
    if agent_node is not None:  # :629-631 with a one-hot node_agent matrix == gather
        agent_out = np.stack([out[b, agent_node[b]] for b in range(B)])
    elif node_agent is not None:
        agent_out = np.einsum("bno,bna->bao", out, np.asarray(node_agent, dtype=dtype))
    return out, new_state, agent_out


def dqn_forward(w, x, activation="leaky_relu", dtype=np.float32):
    """model.py:199-203: Linear(q_net.fc)(MLP encoder with activation on output)."""
    w = {k: np.asarray(v, dtype=dtype) for k, v in w.items()}
    shp = x.shape
    h = mlp(np.asarray(x, dtype=dtype).reshape(-1, shp[-1]), w, "encoder.", _act(activation))
    q = h @ w["q_net.fc.weight"].T + w["q_net.fc.bias"]
    return q.reshape(*shp[:-1], -1)


def epsilon_greedy(q, epsilon, rand_action, rand_u, action_mask=None):
    """policy.py:42-51. rand_action = randint(n_act,size=A) drawn BEFORE rand_u = rand(A)."""
    q = np.array(q, dtype=np.float32, copy=True)
    if action_mask is not None:
        q[np.asarray(action_mask).astype(bool)] = -np.inf
    filt = np.asarray(rand_u) < epsilon
    return np.where(filt, rand_action, np.argmax(q, axis=-1)).astype(np.int64)


def epsilon_decay(eps, step, step_before_train, update_freq, decay):
    """policy.py:55-62."""
    if eps > 0 and step > step_before_train and step % update_freq == 0:
        eps *= decay
        if eps < 0.01:
            eps = 0.01
    return eps


class ReplayIndexOracle:
    """replaybuffer.py:101-130, 284-287: ring counters and index sampling."""

    def __init__(self, seed, buffer_size):
        self.buffer_size, self.count, self.index = buffer_size, 0, 0
        self.rng = np.random.default_rng(seed)

    def add(self):
        slot = self.index
        if self.count < self.buffer_size:
            self.count += 1
        self.index = (self.index + 1) % self.buffer_size
        return slot

    def sample(self, batch_size, sequence_length=0):
        if sequence_length <= 1:
            return [self.rng.choice(self.count, batch_size, replace=True)]
        start = self.index % self.count
        s = self.rng.choice(self.count - sequence_length, batch_size, replace=True)
        s = (start + s) % self.count
        return [(s + o) % self.count for o in range(sequence_length)]
