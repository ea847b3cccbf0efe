#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by EXECUTING the unmodified
reference (/root/reference/src) in this container.

    python tools/gen_golden.py            # writes tests/golden/*.npz

The reference has no tests and no golden vectors of its own for the Routing env,
NetMon, EpsilonGreedy or ReplayBuffer, so parity is pinned to
outputs of the reference itself.  /root/reference does not travel to the GPU box;
the fixtures written here do.  Nothing in tests/, bench.py or the product imports
this script.
"""
import argparse
import hashlib
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_stubs  # noqa: E402

ref_stubs.install()

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from env.constants import EVAL_SEEDS  # noqa: E402
from env.network import Network  # noqa: E402
from env.routing import Routing  # noqa: E402
from env.simple_environment import SimpleEnvironment  # noqa: E402
from env.wrapper import NetMonWrapper  # noqa: E402
from model import DQN, NetMon  # noqa: E402
from policy import EpsilonGreedy  # noqa: E402
from replaybuffer import ReplayBuffer  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}.npz  {os.path.getsize(path)/1024:.1f} kB")


# --------------------------------------------------------------------------
# topology
# --------------------------------------------------------------------------
def topo_arrays(net):
    n = net.n_nodes
    edges = np.array([[e.start, e.end, e.length] for e in net.edges], dtype=np.int32)
    node_edges = np.array([nd.edges for nd in net.nodes], dtype=np.int32)
    node_nbrs_creation = np.array([nd.neighbors for nd in net.nodes], dtype=np.int32)
    apsp = np.array(
        [[net.shortest_paths_weights[i][j] for j in range(n)] for i in range(n)],
        dtype=np.int32,
    )
    xy = np.array([[nd.x, nd.y] for nd in net.nodes], dtype=np.float64)
    return edges, node_edges, node_nbrs_creation, apsp, xy


def gen_topology():
    out = {}
    fixed_seeds = [923430603] + list(EVAL_SEEDS[:60])
    E, NE, NB, AP, XY = [], [], [], [], []
    for s in fixed_seeds:
        np.random.seed(12345)
        net = Network(20, random_topology=False, topology_init_seed=s)
        net.reset()
        assert net.repetitions == 1
        # the global stream must be untouched by a fixed-seed reset
        e, ne, nb, ap, xy = topo_arrays(net)
        E.append(e), NE.append(ne), NB.append(nb), AP.append(ap), XY.append(xy)
    out["fixed_seeds"] = np.array(fixed_seeds, dtype=np.int64)
    out["fixed_edges"] = np.stack(E)
    out["fixed_node_edges"] = np.stack(NE)
    out["fixed_node_nbrs_creation"] = np.stack(NB)
    out["fixed_apsp"] = np.stack(AP)
    out["fixed_xy"] = np.stack(XY)
    out["fixed_adj"] = net.adj_matrix.copy()  # of the last one

    # all 1000 eval seeds: digest only (edges+apsp), keeps the fixture small
    digests = []
    for s in EVAL_SEEDS:
        net = Network(20, random_topology=False, topology_init_seed=s)
        net.reset()
        e, ne, nb, ap, xy = topo_arrays(net)
        h = hashlib.sha256()
        h.update(e.tobytes()), h.update(ne.tobytes()), h.update(ap.tobytes())
        digests.append(np.frombuffer(h.digest()[:8], dtype=np.uint64)[0])
    out["eval_seeds"] = np.array(EVAL_SEEDS, dtype=np.int64)
    out["eval_digest64"] = np.array(digests, dtype=np.uint64)

    # N = 200 synthetic large graph (BASELINE config 4)
    net = Network(200, random_topology=False, topology_init_seed=476)
    net.reset()
    e, ne, nb, ap, xy = topo_arrays(net)
    out["n200_edges"], out["n200_node_edges"], out["n200_apsp"] = e, ne, ap

    # random-topology chain: seeds drawn from the global stream, invalid ones reseeded
    np.random.seed(7)
    net = Network(20, random_topology=True, excluded_seeds=EVAL_SEEDS)
    chain_seed, chain_rep, chain_edges, chain_after = [], [], [], []
    for _ in range(12):
        net.reset()
        chain_seed.append(net.current_topology_seed)
        chain_rep.append(net.repetitions)
        chain_edges.append(topo_arrays(net)[0])
        st = np.random.get_state()
        chain_after.append(st[2])  # MT position: stream restored + advanced only by seed draw
    out["chain_seed"] = np.array(chain_seed, dtype=np.int64)
    out["chain_rep"] = np.array(chain_rep, dtype=np.int32)
    out["chain_edges"] = np.stack(chain_edges)
    out["chain_pos"] = np.array(chain_after, dtype=np.int32)
    out["chain_next_u"] = np.array([np.random.random()], dtype=np.float64)

    # finite seed pool (network.py:100-120)
    np.random.seed(99)
    net = Network(
        20, random_topology=True, n_random_seeds=10, topology_init_seed=476,
        excluded_seeds=EVAL_SEEDS,
    )
    out["pool_seeds"] = np.array(net.seeds, dtype=np.int64)
    picks = []
    for _ in range(8):
        net.reset()  # np.random.choice(seed_list) from the global stream
        picks.append(net.current_topology_seed)
    out["pool_picks"] = np.array(picks, dtype=np.int64)
    # sequential seeds as used for evaluation (main.py:553-555)
    net.seeds = list(EVAL_SEEDS)
    net.sequential_topology_seeds = True
    seq = []
    for _ in range(4):
        net.reset()
        seq.append(net.current_topology_seed)
    out["seq_picks"] = np.array(seq, dtype=np.int64)
    save("topology", **out)


# --------------------------------------------------------------------------
# legacy numpy RNG (MT19937) known answers
# --------------------------------------------------------------------------
def gen_rng():
    out = {}
    seeds = [0, 1, 42, 923430603, 2**32 - 1, 476]
    out["seeds"] = np.array(seeds, dtype=np.uint64)
    rows = []
    for s in seeds:
        np.random.seed(s)
        r = {}
        r["random10"] = np.array([np.random.random() for _ in range(10)])
        r["randint20"] = np.array([np.random.randint(20) for _ in range(16)])
        r["randint200"] = np.array([np.random.randint(200) for _ in range(16)])
        r["randint4_vec"] = np.random.randint(4, size=9)
        r["rand_vec"] = np.random.rand(9)
        r["randint_big"] = np.array([np.random.randint(2**31 - 1) for _ in range(6)])
        r["randint1"] = np.array([np.random.randint(1) for _ in range(3)])
        a = np.arange(7)
        np.random.shuffle(a)
        r["shuffle7"] = a
        lst = [5, 9]
        np.random.shuffle(lst)
        r["shuffle_list2"] = np.array(lst)
        r["choice"] = np.array([np.random.choice([11, 22, 33, 44, 55]) for _ in range(5)])
        r["choice1"] = np.array([np.random.choice([77])])
        r["after"] = np.array([np.random.random()])
        st = np.random.get_state()
        r["state_pos"] = np.array([st[2]])
        r["state_key8"] = st[1][:8].astype(np.uint32)
        rows.append(r)
    for k in rows[0]:
        out[k] = np.stack([r[k] for r in rows])
    save("rng", **out)


# --------------------------------------------------------------------------
# routing trajectories
# --------------------------------------------------------------------------
class DrawRecorder:
    """Records (step, packet id, start, target, size) for every reset_packet call."""

    def __init__(self, env):
        self.env = env
        self.t = -1
        self.rows = []
        orig = env.reset_packet

        def wrapped(packet):
            orig(packet)
            self.rows.append((self.t, packet.id, packet.start, packet.target, packet.size))

        env.reset_packet = wrapped

    def arrays(self):
        r = self.rows
        return dict(
            draw_t=np.array([x[0] for x in r], dtype=np.int32),
            draw_id=np.array([x[1] for x in r], dtype=np.int32),
            draw_start=np.array([x[2] for x in r], dtype=np.int32),
            draw_target=np.array([x[3] for x in r], dtype=np.int32),
            draw_size=np.array([x[4] for x in r], dtype=np.float64),
        )


def env_state(env):
    d = env.data
    VW = (env.network.n_nodes + 31) // 32
    vis = np.zeros((len(d), VW), dtype=np.uint32)
    for i, p in enumerate(d):
        for v in p.visited_nodes:
            vis[i, v // 32] |= np.uint32(1 << (v % 32))
    return dict(
        now=np.array([p.now for p in d], dtype=np.int32),
        target=np.array([p.target for p in d], dtype=np.int32),
        edge=np.array([p.edge for p in d], dtype=np.int32),
        time=np.array([p.time for p in d], dtype=np.int32),
        ttl=np.array([p.ttl for p in d], dtype=np.int32),
        spw=np.array([p.shortest_path_weight for p in d], dtype=np.int32),
        start=np.array([p.start for p in d], dtype=np.int32),
        size=np.array([p.size for p in d], dtype=np.float64),
        load=np.array([e.load for e in env.network.edges], dtype=np.float64),
        agent_steps=env.agent_steps.astype(np.int32).copy(),
        visited=vis,
        mask=env.action_mask.astype(np.uint8).copy(),
    )


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8).copy()


def gen_routing_case(
    name, n_nodes, topo_seed, n_data, steps, env_var=1, congestion=True, mask=False,
    ttl=0, np_seed=0, act_seed=123, dense=True, eval_info=False, dense_steps=None,
):
    np.random.seed(np_seed)
    net = Network(n_nodes, random_topology=False, topology_init_seed=topo_seed)
    env = Routing(net, n_data, env_var, enable_congestion=congestion,
                  enable_action_mask=mask, ttl=ttl)
    env.set_eval_info(eval_info)
    rec = DrawRecorder(env)
    obs, adj = env.reset()
    act_rng = np.random.RandomState(act_seed)

    keys_state = None
    S = []  # state per t (0 = after reset)
    OBS, ADJ, NOBS, NAM = [obs], [adj], [env.get_node_observation()], [env.get_node_agent_matrix()]
    HASH = [sha(obs, adj, NOBS[0], NAM[0])]
    S.append(env_state(env))
    ACT, REW, DONE, INFO, DEL, ARR, SPR = [], [], [], [], [], [], []
    EXTRA = []
    for t in range(steps):
        rec.t = t
        act = act_rng.randint(4, size=n_data).astype(np.int32)
        steps_before = env.agent_steps.copy() + 1
        spw_before = np.array([p.shortest_path_weight for p in env.data])
        obs, adj, reward, done, info = env.step(act)
        assert reward.dtype == np.float32
        nobs, nam = env.get_node_observation(), env.get_node_agent_matrix()
        ACT.append(act), REW.append(reward), DONE.append(done.astype(np.uint8))
        INFO.append([info["looped"], info["throughput"], info["dropped"], info["blocked"]])
        # per-agent reconstruction of the ragged info lists
        delays = np.where(done, steps_before, 0).astype(np.int32)
        arrived = (reward > 5).astype(np.uint8)  # +10 / +9.8 only on success
        assert arrived.sum() == info["throughput"]
        assert sorted(delays[done].tolist()) == sorted(int(x) for x in info["delays"])
        spr = np.where(arrived == 1, steps_before / np.maximum(spw_before, 1), 0.0)
        assert np.allclose(sorted(spr[arrived == 1]), sorted(info["spr"]))
        DEL.append(delays), ARR.append(arrived), SPR.append(spr)
        if eval_info:
            EXTRA.append([info["total_edge_load"], info["occupied_edges"],
                          info["packets_on_edges"], info["total_packet_size"]])
        HASH.append(sha(obs, adj, reward, done, nobs, nam))
        if dense or (dense_steps is not None and t < dense_steps):
            OBS.append(obs), ADJ.append(adj), NOBS.append(nobs), NAM.append(nam)
        S.append(env_state(env))

    edges, node_edges, _, apsp, _ = topo_arrays(net)
    out = dict(
        cfg=np.array([n_nodes, n_data, env_var, int(congestion), int(mask), ttl, steps,
                      topo_seed, int(eval_info), env.k], dtype=np.int64),
        edges=edges, node_edges=node_edges, apsp=apsp,
        actions=np.stack(ACT), reward=np.stack(REW), done=np.stack(DONE),
        info=np.array(INFO, dtype=np.float64), delays=np.stack(DEL), arrived=np.stack(ARR),
        spr=np.stack(SPR), hash=np.stack(HASH),
        obs=np.stack(OBS), adj=np.stack(ADJ), node_obs=np.stack(NOBS), node_agent=np.stack(NAM),
        final_delays=np.array(env.get_final_info({"delays": []})["delays"], dtype=np.float64),
        node_aux=env.get_node_aux(),
    )
    if eval_info:
        out["eval_extra"] = np.array(EXTRA, dtype=np.float64)
        out["sum_packets_per_node"] = env.sum_packets_per_node
        out["sum_packets_per_edge"] = env.sum_packets_per_edge
    for k in S[0]:
        out["s_" + k] = np.stack([s[k] for s in S])
    out.update(rec.arrays())
    save(name, **out)
    return out


def gen_routing():
    a = gen_routing_case("routing_A_seed923430603_cong", 20, 923430603, 20, 300)
    # SURVEY §8c anchors
    assert abs(float(a["reward"].sum()) - 257.6) < 1e-3, a["reward"].sum()
    assert a["info"][:, 1].sum() == 37 and a["info"][:, 3].sum() == 562
    gen_routing_case("routing_B_nocong", 20, 923430603, 20, 120, congestion=False)
    gen_routing_case("routing_C_mask", 20, EVAL_SEEDS[0], 20, 120, mask=True, np_seed=1)
    gen_routing_case("routing_D_ttl", 20, EVAL_SEEDS[1], 20, 120, ttl=12, np_seed=2)
    gen_routing_case("routing_E_a35_nocong_mask_ttl", 20, EVAL_SEEDS[2], 35, 120,
                     congestion=False, mask=True, ttl=9, np_seed=3)
    gen_routing_case("routing_F_n200", 200, 476, 100, 40, np_seed=4, dense=False, dense_steps=2)
    gen_routing_case("routing_G_var2", 20, 923430603, 20, 40, env_var=2, np_seed=5)
    gen_routing_case("routing_H_var3", 20, 923430603, 20, 40, env_var=3, np_seed=6)
    gen_routing_case("routing_I_evalinfo", 20, EVAL_SEEDS[3], 20, 60, eval_info=True, np_seed=7)


# --------------------------------------------------------------------------
# simple environment
# --------------------------------------------------------------------------
def gen_simple():
    out = {}
    for var in (1, 3):
        for rt in (0, 1):
            np.random.seed(10 + var + rt)
            env = SimpleEnvironment(env_var=var, random_topology=rt)
            OBS, NOBS, NAM, ADJN, REW, ACT, EDGE, SCORE, START = [], [], [], [], [], [], [], [], []
            ar = np.random.RandomState(3)
            for ep in range(12):
                obs, adj = env.reset()
                a = int(ar.randint(2))
                OBS.append(obs), NOBS.append(env.get_node_observation())
                NAM.append(env.get_node_agent_matrix()), ADJN.append(env.get_nodes_adjacency())
                EDGE.append([[e.start, e.end] for e in env.edges])
                SCORE.append([r.score for r in env.router])
                START.append([env.start_node] + list(env.router[env.start_node].edge))
                obs2, adj2, rew, done, info = env.step([a])
                assert done == [True] and np.array_equal(obs2, obs)
                REW.append(rew), ACT.append(a)
            tag = f"v{var}_rt{rt}_"
            out[tag + "obs"] = np.stack(OBS)
            out[tag + "node_obs"] = np.stack(NOBS)
            out[tag + "node_agent"] = np.stack(NAM)
            out[tag + "node_adj"] = np.stack(ADJN)
            out[tag + "reward"] = np.stack(REW).astype(np.float64)
            out[tag + "act"] = np.array(ACT, dtype=np.int32)
            out[tag + "edges"] = np.array(EDGE, dtype=np.int32)
            out[tag + "scores"] = np.array(SCORE, dtype=np.int32)
            out[tag + "start_edges"] = np.array(START, dtype=np.int32)
            out[tag + "after"] = np.array([np.random.random()])
    save("simple_env", **out)


# --------------------------------------------------------------------------
# NetMon / DQN numerics (torch CPU fp32 is what the reference runs)
# --------------------------------------------------------------------------
def det_weights(module, seed):
    """Deterministic, framework-independent parameter values (no torch RNG):
    each tensor = default_rng(seed + i).uniform(-b, b) with b = 1/sqrt(fan) like
    torch's default init scale; LayerNorm weights near 1."""
    sd = module.state_dict()
    new = {}
    for i, (k, v) in enumerate(sd.items()):
        rng = np.random.default_rng(seed + i)
        if v.dim() >= 2:
            bound = 1.0 / np.sqrt(v.shape[1])
        else:
            bound = 0.1
        arr = rng.uniform(-bound, bound, size=tuple(v.shape)).astype(np.float32)
        if ".ln_" in k or k.startswith("ln_"):
            if k.endswith("weight"):
                arr = (1.0 + arr).astype(np.float32)
        new[k] = torch.from_numpy(arr)
    module.load_state_dict(new)
    return {k: v.numpy().copy() for k, v in new.items()}


def routing_inputs(n_envs, n_nodes, n_data, seeds, steps, np_seed=0):
    """node_obs / adj / node_agent from live reference envs driven by random actions."""
    np.random.seed(np_seed)
    envs = []
    for s in seeds[:n_envs]:
        net = Network(n_nodes, random_topology=False, topology_init_seed=s)
        env = Routing(net, n_data, 1)
        env.reset()
        envs.append(env)
    ar = np.random.RandomState(77)
    X, ADJ, NAM, OBS = [], [], [], []
    for t in range(steps):
        X.append(np.stack([e.get_node_observation() for e in envs]))
        ADJ.append(np.stack([e.get_nodes_adjacency() for e in envs]))
        NAM.append(np.stack([e.get_node_agent_matrix() for e in envs]))
        OBS.append(np.stack([e._get_observation() for e in envs]))
        for e in envs:
            e.step(ar.randint(4, size=n_data))
    return np.stack(X), np.stack(ADJ), np.stack(NAM), np.stack(OBS)


def gen_netmon():
    torch.set_num_threads(1)
    act = F.leaky_relu
    seeds = [923430603] + list(EVAL_SEEDS[:7])
    X, ADJ, NAM, OBS = routing_inputs(3, 20, 20, seeds, steps=4)
    base = dict(node_obs=X, node_adj=ADJ, node_agent=NAM, agent_obs=OBS)
    cases = {
        # name: (H, enc, K, rnn, carry, agg, nbr, glob)
        "lstm_sum_k3": (32, (48, 40), 3, "lstm", True, "sum", True, False),
        "lstm_mean_k1": (32, (48, 40), 1, "lstm", True, "mean", True, False),
        "lnlstm_sum_k2": (32, (48, 40), 2, "lnlstm", True, "sum", True, False),
        "gru_sum_k2": (32, (48, 40), 2, "gru", True, "sum", True, False),
        "none_sum_k2": (32, (48, 40), 2, "none", True, "sum", True, False),
        "lstm_nocarry_k2": (32, (48, 40), 2, "lstm", False, "sum", True, False),
        "gru_nocarry_k2": (32, (48, 40), 2, "gru", False, "sum", True, False),
        "lstm_global_k2": (32, (48, 40), 2, "lstm", True, "sum", True, True),
        "lstm_nonbr_k2": (32, (48,), 2, "lstm", True, "sum", False, False),
        "lstm_sum_k3_paper": (128, (512, 256), 3, "lstm", True, "sum", True, False),
        "lnlstm_sum_k4_paper": (128, (512, 256), 4, "lnlstm", True, "sum", True, False),
    }
    out = dict(base)
    names = []
    for name, (H, enc, K, rnn, carry, agg, nbr, glob) in cases.items():
        nm = NetMon(X.shape[-1], H, enc, K, act, rnn_type=rnn, rnn_carryover=carry,
                    agg_type=agg, output_neighbor_hidden=nbr, output_global_hidden=glob)
        wseed = 1000 + 17 * len(names)
        det_weights(nm, wseed)
        nm.eval()
        nm.state = None
        outs, states = [], []
        with torch.no_grad():
            for t in range(X.shape[0]):
                o = nm(torch.from_numpy(X[t]), torch.from_numpy(ADJ[t]).float(),
                       torch.from_numpy(NAM[t]).float())
                outs.append(o.numpy().copy())
                states.append(nm.state.numpy().copy())
        out[name + "_cfg"] = np.array([H, K, int(carry), int(nbr), int(glob), wseed] + list(enc),
                                      dtype=np.int64)
        out[name + "_agent_out"] = np.stack(outs)
        out[name + "_state"] = np.stack(states)
        names.append(f"{name}|{rnn}|{agg}")
    out["case_names"] = np.array(names)

    # node-level output (no_agent_mapping) for the paper-size lstm case, single step
    # simple env degree-2 readout (Appendix D.3): max_degree from the mask
    np.random.seed(21)
    senv = SimpleEnvironment(env_var=1, random_topology=1)
    senv.reset()
    nm = NetMon(1, 8, (6,), 2, act, rnn_type="lstm", output_neighbor_hidden=True)
    det_weights(nm, 4242)
    with torch.no_grad():
        so = nm(torch.from_numpy(senv.get_node_observation()).unsqueeze(0),
                torch.from_numpy(senv.get_nodes_adjacency()).float().unsqueeze(0),
                torch.from_numpy(senv.get_node_agent_matrix()).float().unsqueeze(0),
                no_agent_mapping=True)
    out["simple_node_obs"] = senv.get_node_observation()
    out["simple_node_adj"] = senv.get_nodes_adjacency()
    out["simple_node_out"] = so.numpy()
    save("netmon", **out)


def gen_sl():
    """BASELINE config 5 (sl.py): the test split's first 32 graphs exactly as sl.py:575-592 +
    build_dataset (:230-281) draws them (EVAL_SEEDS in order, one env.reset() per sample), the
    per-node labels / targets of get_sl_sample (:174-219), and NetMonSL.forward's NetMon call
    (:153-157: eye node-agent matrix, state None at sequence start, the SAME node obs fed for
    `sequence_length` steps, :481-493) at the paper dims for K in {1,2,4}."""
    torch.set_num_threads(1)
    np.random.seed(5)
    n_graphs, seq = 32, 8
    net = Network(20, random_topology=True, sequential_topology_seeds=True, provided_seeds=EVAL_SEEDS)
    env = Routing(net, 20, 1)
    env.network.G_weight_key = "weight"
    env.reset()
    X, ADJ, LAB, TGT, TGT_ALL, SEEDS = [], [], [], [], [], []
    for s in range(n_graphs):
        if s:
            env.reset()
        n = env.get_num_nodes()
        lab = np.zeros(n)
        for v in range(n):
            path = env.network.shortest_paths[v][0]
            if len(path) > 1:
                for e_idx, e in enumerate(env.network.nodes[v].edges):
                    if env.network.edges[e].get_other_node(v) == path[1]:
                        lab[v] = e_idx + 1
                        break
                else:
                    raise AssertionError("no edge towards the next hop")
        apsp = np.array([[env.network.shortest_paths_weights[i][j] for j in range(n)] for i in range(n)])
        X.append(env.get_node_observation()), ADJ.append(env.get_nodes_adjacency())
        LAB.append(lab), TGT.append(apsp[:, 0]), TGT_ALL.append(apsp)
        SEEDS.append(env.network.current_topology_seed)
    X, ADJ = np.stack(X).astype(np.float32), np.stack(ADJ)
    out = dict(node_obs=X, node_adj=ADJ.astype(np.int8), labels=np.stack(LAB).astype(np.int8),
               targets=np.stack(TGT).astype(np.int32), targets_all=np.stack(TGT_ALL).astype(np.int32),
               seeds=np.array(SEEDS, dtype=np.int64), seq=np.array([seq]))
    eye = torch.eye(20).repeat(n_graphs, 1, 1)
    names = []
    for K in (1, 2, 4):
        nm = NetMon(X.shape[-1], 128, (512, 256), K, F.leaky_relu, rnn_type="lstm", rnn_carryover=True,
                    agg_type="sum", output_neighbor_hidden=True, output_global_hidden=False)
        wseed = 7000 + K
        det_weights(nm, wseed)
        nm.eval()
        nm.state = None
        with torch.no_grad():
            for t in range(seq):
                o = nm(torch.from_numpy(X), torch.from_numpy(ADJ).float(), eye)
                if t == 0:
                    out[f"k{K}_out_first"] = o[:4].numpy().copy()
        name = f"k{K}"
        out[name + "_cfg"] = np.array([128, K, 1, 1, 0, wseed, 512, 256], dtype=np.int64)
        out[name + "_out_last"] = o.numpy().copy()
        out[name + "_state_last"] = nm.state[:4].numpy().copy()
        names.append(f"{name}|lstm|sum")
    out["case_names"] = np.array(names)
    save("sl_netmon", **out)


def gen_dqn_policy():
    torch.set_num_threads(1)
    act = F.leaky_relu
    rng = np.random.default_rng(5)
    A, D = 20, 130 + 4 * 32
    dqn = DQN(D, (64, 48), 4, act)
    det_weights(dqn, 9000)
    dqn.eval()
    T = 12
    obs = rng.standard_normal((T, A, D)).astype(np.float32)
    obs[:, :, :130] = (rng.random((T, A, 130)) < 0.05).astype(np.float32)
    adj = np.ones((T, A, A), dtype=np.int8)
    with torch.no_grad():
        q = dqn(torch.from_numpy(obs), torch.from_numpy(adj).float()).numpy()

    env = SimpleNamespace(enable_action_mask=True,
                          action_mask=np.zeros((A, 4), dtype=bool))
    args = SimpleNamespace(epsilon=0.5, step_before_train=3, epsilon_update_freq=2,
                           epsilon_decay=0.5)
    pol = EpsilonGreedy(env, dqn, 4, args)
    np.random.seed(5)
    acts, eps, masks, rand_a, rand_u = [], [], [], [], []
    mrng = np.random.default_rng(8)
    for t in range(T):
        m = mrng.random((A, 4)) < 0.3
        m[:, 0] = False if t % 2 else m[:, 0]
        m[m.all(axis=1), 1] = False
        env.action_mask = m
        st = np.random.get_state()
        ra = np.random.randint(4, size=A)
        ru = np.random.rand(A)
        np.random.set_state(st)
        a = pol(obs[t], adj[t])
        acts.append(a), eps.append(pol._epsilon), masks.append(m.astype(np.uint8))
        rand_a.append(ra), rand_u.append(ru)
    pol.eval()
    eps_eval = pol._epsilon
    pol.train()
    eps_train = pol._epsilon  # stays 0: reference quirk (Appendix D.1)
    save("dqn_policy", obs=obs, q=q, cfg=np.array([D, 64, 48, 4, 9000], dtype=np.int64),
         actions=np.stack(acts).astype(np.int32), eps_after=np.array(eps),
         masks=np.stack(masks), rand_action=np.stack(rand_a).astype(np.int32),
         rand_u=np.stack(rand_u), eps_eval=np.array([eps_eval]),
         eps_train=np.array([0.0 if eps_train is None else eps_train]),
         eps_train_is_none=np.array([eps_train is None]))


def gen_replay():
    rng = np.random.default_rng(11)
    cap, A, D, S, N, Dn, Sn, Ax = 16, 3, 5, 0, 4, 6, 8, 4
    rb = ReplayBuffer(3, cap, A, D, S, N, Dn, Sn, Ax)
    fields = {}
    n_add = 23

    def r(*shape, dt=np.float32):
        return rng.standard_normal(shape).astype(dt)

    tr = dict(
        obs=r(n_add, A, D), action=rng.integers(0, 4, (n_add, A)).astype(np.int32),
        reward=r(n_add, A), next_obs=r(n_add, A, D),
        adj=(rng.random((n_add, A, A)) < 0.5).astype(np.int8),
        next_adj=(rng.random((n_add, A, A)) < 0.5).astype(np.int8),
        done=rng.random((n_add, A)) < 0.2, episode_done=rng.random(n_add) < 0.1,
        node_state=r(n_add, N, Sn), node_aux=r(n_add, N, Ax), node_obs=r(n_add, N, Dn),
        node_adj=(rng.random((n_add, N, N)) < 0.5).astype(np.int8),
        node_agent=(rng.random((n_add, N, A)) < 0.5).astype(np.int8),
        next_node_obs=r(n_add, N, Dn),
        next_node_adj=(rng.random((n_add, N, N)) < 0.5).astype(np.int8),
        next_node_agent=(rng.random((n_add, N, A)) < 0.5).astype(np.int8),
    )
    samples = {}
    for i in range(n_add):
        rb.add(tr["obs"][i], tr["action"][i], tr["reward"][i], tr["next_obs"][i], tr["adj"][i],
               tr["next_adj"][i], tr["done"][i], tr["episode_done"][i], 0, tr["node_state"][i],
               tr["node_aux"][i], tr["node_obs"][i], tr["node_adj"][i], tr["node_agent"][i],
               tr["next_node_obs"][i], tr["next_node_adj"][i], tr["next_node_agent"][i])
        if i == 9:  # not yet full
            b = next(rb.get_batch(4, "cpu"))
            samples["idx_partial"] = np.asarray(b.idx)
            samples["obs_partial"] = b.obs.numpy()
            seq = list(rb.get_batch(4, "cpu", sequence_length=3))
            samples["idx_seq_partial"] = np.stack([np.asarray(s.idx) for s in seq])
    b = next(rb.get_batch(6, "cpu"))
    samples["idx_full"] = np.asarray(b.idx)
    for f in b._fields:
        if f != "idx":
            samples["full_" + f] = getattr(b, f).numpy()
    seq = list(rb.get_batch(5, "cpu", sequence_length=4))
    samples["idx_seq_full"] = np.stack([np.asarray(s.idx) for s in seq])
    samples["seq_full_node_state"] = np.stack([s.node_state.numpy() for s in seq])
    samples["final_index_count"] = np.array([rb.index, rb.count])
    samples["cfg"] = np.array([3, cap, A, D, S, N, Dn, Sn, Ax, n_add], dtype=np.int64)
    save("replay", **{("tr_" + k): v for k, v in tr.items()}, **samples)


def gen_wrapper():
    """NetMonWrapper + EpsilonGreedy closed loop at small dims: pins last_netmon_state
    semantics, startup iterations and the order of global-RNG consumption."""
    torch.set_num_threads(1)
    act = F.leaky_relu
    np.random.seed(31)
    net = Network(20, random_topology=False, topology_init_seed=923430603)
    env0 = Routing(net, 20, 1)
    nm = NetMon(88, 16, (24,), 2, act, rnn_type="lstm", output_neighbor_hidden=True)
    det_weights(nm, 555)
    nm.eval()
    env = NetMonWrapper(env0, nm, 2)
    rec = DrawRecorder(env0)
    obs, adj = env.reset()
    OBS, LAST, CUR, ACT = [obs], [], [], []
    ar = np.random.RandomState(9)
    for t in range(6):
        rec.t = t
        LAST.append(env.last_netmon_state.numpy().copy())
        a = ar.randint(4, size=20).astype(np.int32)
        obs, adj, rew, done, info = env.step(a)
        OBS.append(obs), CUR.append(env.current_netmon_state.numpy().copy()), ACT.append(a)
    save("wrapper", joint_obs=np.stack(OBS), last_state=np.stack(LAST), cur_state=np.stack(CUR),
         actions=np.stack(ACT), cfg=np.array([88, 16, 24, 2, 555, 2], dtype=np.int64),
         **rec.arrays())


def gen_freeze():
    """NetMonWrapper.freeze() (wrapper.py:53-75): message passing stops, agents keep reading the frozen node
    readout at their new positions; `data` read-back views (routing.py:12-40) for heuristic policies."""
    torch.set_num_threads(1)
    np.random.seed(77)
    net = Network(20, random_topology=False, topology_init_seed=923430603)
    env0 = Routing(net, 20, 1)
    nm = NetMon(88, 16, (24,), 2, F.leaky_relu, rnn_type="lstm", output_neighbor_hidden=True)
    det_weights(nm, 777)
    nm.eval()
    env = NetMonWrapper(env0, nm, 1)
    rec = DrawRecorder(env0)
    obs, adj = env.reset()
    OBS, ACT, CUR, NAM, DATA = [obs], [], [], [], []
    ar = np.random.RandomState(19)
    for t in range(8):
        rec.t = t
        if t == 3:
            env.freeze()
        a = ar.randint(4, size=20).astype(np.int32)
        obs, adj, rew, done, info = env.step(a)
        OBS.append(obs), ACT.append(a), CUR.append(env.current_netmon_state.numpy().copy())
        NAM.append(env.get_netmon_info()[2].copy())
        DATA.append(np.array([[p.now, p.target, p.edge, p.time, p.ttl, p.shortest_path_weight, p.start]
                              for p in env0.data], dtype=np.int32))
    save("wrapper_freeze", joint_obs=np.stack(OBS), actions=np.stack(ACT), cur_state=np.stack(CUR),
         node_agent=np.stack(NAM), data=np.stack(DATA), sizes=np.array([p.size for p in env0.data]),
         freeze_at=np.array([3]), cfg=np.array([88, 16, 24, 2, 777, 1], dtype=np.int64), **rec.arrays())


def gen_replay_half():
    """ReplayBuffer(half_precision=True) (replaybuffer.py:52-54, 132-187): float fields stored as fp16 and
    returned as fp32."""
    rng = np.random.default_rng(12)
    cap, A, D, S, N, Dn, Sn, Ax = 8, 3, 5, 0, 4, 6, 8, 4
    rb = ReplayBuffer(5, cap, A, D, S, N, Dn, Sn, Ax, half_precision=True)
    n_add = 11

    def r(*shape):
        return (rng.standard_normal(shape) * 3).astype(np.float32)

    tr = dict(
        obs=r(n_add, A, D), action=rng.integers(0, 4, (n_add, A)).astype(np.int32),
        reward=r(n_add, A), next_obs=r(n_add, A, D),
        adj=(rng.random((n_add, A, A)) < 0.5).astype(np.int8),
        next_adj=(rng.random((n_add, A, A)) < 0.5).astype(np.int8),
        done=rng.random((n_add, A)) < 0.2, episode_done=rng.random(n_add) < 0.1,
        node_state=r(n_add, N, Sn), node_aux=r(n_add, N, Ax), node_obs=r(n_add, N, Dn),
        node_adj=(rng.random((n_add, N, N)) < 0.5).astype(np.int8),
        node_agent=(rng.random((n_add, N, A)) < 0.5).astype(np.int8),
        next_node_obs=r(n_add, N, Dn),
        next_node_adj=(rng.random((n_add, N, N)) < 0.5).astype(np.int8),
        next_node_agent=(rng.random((n_add, N, A)) < 0.5).astype(np.int8),
    )
    for i in range(n_add):
        rb.add(tr["obs"][i], tr["action"][i], tr["reward"][i], tr["next_obs"][i], tr["adj"][i],
               tr["next_adj"][i], tr["done"][i], tr["episode_done"][i], 0, tr["node_state"][i],
               tr["node_aux"][i], tr["node_obs"][i], tr["node_adj"][i], tr["node_agent"][i],
               tr["next_node_obs"][i], tr["next_node_adj"][i], tr["next_node_agent"][i])
    b = next(rb.get_batch(6, "cpu"))
    samples = {"idx_full": np.asarray(b.idx)}
    for f in b._fields:
        if f != "idx":
            samples["full_" + f] = getattr(b, f).numpy()
    samples["cfg"] = np.array([5, cap, A, D, S, N, Dn, Sn, Ax, n_add], dtype=np.int64)
    save("replay_half", **{("tr_" + k): v for k, v in tr.items()}, **samples)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    gens = dict(topology=gen_topology, rng=gen_rng, routing=gen_routing, simple=gen_simple,
                netmon=gen_netmon, sl=gen_sl, dqn=gen_dqn_policy, replay=gen_replay, wrapper=gen_wrapper,
                freeze=gen_freeze, replay_half=gen_replay_half)
    for k, fn in gens.items():
        if a.only and k not in a.only.split(","):
            continue
        fn()
