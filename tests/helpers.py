"""Shared helpers for parity tests (golden fixture access, draw tables)."""
import numpy as np


def topo_from_golden(g):
    return dict(edges=g["edges"], node_edges=g["node_edges"], apsp=g["apsp"])


def draws_for_step(g, t, A):
    """Draw slots for step t (t=-1: the reset): slot s = s-th reset in id order."""
    sel = g["draw_t"] == t
    n = int(sel.sum())
    ds = np.zeros(A, np.int32)
    dt = np.zeros(A, np.int32)
    dz = np.zeros(A, np.float64)
    ds[:n], dt[:n], dz[:n] = g["draw_start"][sel], g["draw_target"][sel], g["draw_size"][sel]
    assert np.all(np.diff(g["draw_id"][sel]) > 0)  # id order
    return ds, dt, dz, n


def cfg_from_golden(g):
    n_nodes, n_data, env_var, cong, mask, ttl, steps, topo_seed, eval_info, k = [int(x) for x in g["cfg"]]
    return dict(n_nodes=n_nodes, n_data=n_data, env_var=env_var, congestion=bool(cong),
                mask=bool(mask), ttl=ttl, steps=steps, topo_seed=topo_seed,
                eval_info=bool(eval_info), k=k)


STATE_KEYS = ["now", "target", "edge", "time", "ttl", "spw", "start", "size", "load",
              "agent_steps", "visited", "mask"]


def det_weights(shapes, seed):
    """Same deterministic parameter values as tools/gen_golden.py:det_weights.
    `shapes`: ordered dict name -> shape, in the reference module's state_dict order."""
    out = {}
    for i, (k, shp) in enumerate(shapes.items()):
        rng = np.random.default_rng(seed + i)
        bound = 1.0 / np.sqrt(shp[1]) if len(shp) >= 2 else 0.1
        arr = rng.uniform(-bound, bound, size=tuple(shp)).astype(np.float32)
        if (".ln_" in k or k.startswith("ln_")) and k.endswith("weight"):
            arr = (1.0 + arr).astype(np.float32)
        out[k] = arr
    return out


def netmon_shapes(in_features, H, enc, rnn):
    """state_dict key order of the reference NetMon (model.py:256-401): encode MLP, then
    rnn_obs, then rnn_update; LSTMCell/GRUCell: weight_ih, weight_hh, bias_ih, bias_hh;
    LayerNormLSTMCell (layernormlstm.py:15-22): weight_ih, weight_hh, bias_ih, ln_input.*,
    ln_hidden.*, ln_cell.*."""
    s = {}
    prev = in_features
    for i, u in enumerate(list(enc) + [H]):
        s[f"encode.linear_layers.{i}.weight"] = (u, prev)
        s[f"encode.linear_layers.{i}.bias"] = (u,)
        prev = u
    G = {"lstm": 4, "lnlstm": 4, "gru": 3}.get(rnn, 0)
    for cell in ("rnn_obs", "rnn_update"):
        if G == 0:
            continue
        s[f"{cell}.weight_ih"] = (G * H, H)
        s[f"{cell}.weight_hh"] = (G * H, H)
        s[f"{cell}.bias_ih"] = (G * H,)
        if rnn == "lnlstm":
            s[f"{cell}.ln_input.weight"] = (4 * H,)
            s[f"{cell}.ln_input.bias"] = (4 * H,)
            s[f"{cell}.ln_hidden.weight"] = (4 * H,)
            s[f"{cell}.ln_hidden.bias"] = (4 * H,)
            s[f"{cell}.ln_cell.weight"] = (H,)
            s[f"{cell}.ln_cell.bias"] = (H,)
        else:
            s[f"{cell}.bias_hh"] = (G * H,)
    return s


def dqn_shapes(in_features, hidden, n_act):
    s = {}
    prev = in_features
    for i, u in enumerate(hidden):
        s[f"encoder.linear_layers.{i}.weight"] = (u, prev)
        s[f"encoder.linear_layers.{i}.bias"] = (u,)
        prev = u
    s["q_net.fc.weight"] = (n_act, prev)
    s["q_net.fc.bias"] = (n_act,)
    return s


def netmon_case(g, entry):
    name, rnn, agg = str(entry).split("|")
    cfg = [int(x) for x in g[name + "_cfg"]]
    H, K, carry, nbr, glob, wseed = cfg[:6]
    enc = cfg[6:]
    return name, dict(hidden=H, iterations=K, rnn_type=rnn, rnn_carryover=bool(carry), agg_type=agg,
                      output_neighbor_hidden=bool(nbr), output_global_hidden=bool(glob),
                      activation="leaky_relu", enc=enc, wseed=wseed)
