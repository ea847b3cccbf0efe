"""BASELINE config 5 (the supervised shortest-path task of the reference's src/sl.py): the test split's
dataset (graphs, node observations, labels, targets) and the NetMon forward that NetMonSL.forward issues
(sl.py:153-157: identity node-agent matrix, state None at sequence start, the same node observations for
`sequence_length` steps), against outputs of the unmodified reference in tests/golden/sl_netmon.npz
(tools/gen_golden.py::gen_sl).

CPU part: the numpy oracle and the host-side Network (native generator) against the fixture.
GPU part (-m gpu): the CUDA path through the reference's class surface, fp32 and bf16x3 arithmetic, the
autograd (learner) path beside the kernel path, and the node observations of the compat env.

Stated tolerances (max abs error over an 8-step recurrent sequence, outputs O(1)): oracle 5e-6;
CUDA fp32 1e-4; tensor-core bf16x3 3e-4 (the bars of tests/test_gpu_netmon.py / test_gpu_tensorcore.py).
"""
import numpy as np
import pytest

from conftest import load_golden
from helpers import det_weights, netmon_case, netmon_shapes

G = load_golden("sl_netmon")
CASES = [str(x) for x in G["case_names"]]
SEQ = int(G["seq"][0])


def _weights(cfg, in_features):
    return det_weights(netmon_shapes(in_features, cfg["hidden"], cfg["enc"], cfg["rnn_type"]), cfg["wseed"])


@pytest.mark.parametrize("entry", CASES)
def test_oracle_netmon_sequence(entry):
    from oracle import netmon_oracle as NO

    name, cfg = netmon_case(G, entry)
    X, ADJ = G["node_obs"], G["node_adj"]
    w = _weights(cfg, X.shape[-1])
    eye = np.broadcast_to(np.eye(X.shape[1], dtype=np.float32), (X.shape[0], X.shape[1], X.shape[1]))
    state = None
    for t in range(SEQ):
        _, state, agent_out = NO.netmon_forward(w, cfg, X, ADJ, state, node_agent=eye)
        if t == 0:
            assert np.abs(agent_out[:4] - G[name + "_out_first"]).max() < 5e-6
    assert agent_out.shape == G[name + "_out_last"].shape == (32, 20, 4 * cfg["hidden"])
    assert np.abs(agent_out - G[name + "_out_last"]).max() < 5e-6
    assert np.abs(state[:4] - G[name + "_state_last"]).max() < 5e-6


def _sl_labels(net):
    """get_sl_sample's per-node class label and distance targets (sl.py:174-219) from the Network API."""
    n = net.n_nodes
    lab = np.zeros(n, np.int8)
    for v in range(n):
        path = net.shortest_paths[v][0]
        if len(path) > 1:
            hops = [net.edges[e].get_other_node(v) for e in net.nodes[v].edges]
            lab[v] = hops.index(path[1]) + 1
    apsp = np.asarray(net.shortest_paths_weights)
    return lab, apsp


def test_host_network_reproduces_the_test_split():
    """sl.py:582-587: Network(20, random_topology=True, sequential_topology_seeds=True,
    provided_seeds=EVAL_SEEDS); every reset() moves on to the next evaluation graph."""
    from graph_marl_b200.env.constants import EVAL_SEEDS
    from graph_marl_b200.env.network import Network

    net = Network(20, random_topology=True, sequential_topology_seeds=True, provided_seeds=EVAL_SEEDS)
    net.G_weight_key = "weight"
    for s in range(G["node_obs"].shape[0]):
        net.reset()
        assert net.current_topology_seed == int(G["seeds"][s])
        lab, apsp = _sl_labels(net)
        assert np.array_equal(apsp, G["targets_all"][s]) and np.array_equal(apsp[:, 0], G["targets"][s])
        assert np.array_equal(lab, G["labels"][s]), s
        assert np.array_equal(np.asarray(net.get_nodes_adjacency()), G["node_adj"][s])


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
def _netmon(cfg, in_features, math):
    import torch
    import torch.nn.functional as F
    from graph_marl_b200.model import NetMon

    nm = NetMon(in_features, cfg["hidden"], cfg["enc"], cfg["iterations"], F.leaky_relu, rnn_type=cfg["rnn_type"],
                rnn_carryover=cfg["rnn_carryover"], agg_type=cfg["agg_type"],
                output_neighbor_hidden=cfg["output_neighbor_hidden"],
                output_global_hidden=cfg["output_global_hidden"], math=math)
    nm.load_state_dict({k: torch.from_numpy(v) for k, v in _weights(cfg, in_features).items()})
    return nm.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("math,tol", [("fp32", 1e-4), ("bf16x3", 3e-4)])
@pytest.mark.parametrize("entry", CASES)
def test_gpu_netmon_sequence(entry, math, tol):
    import torch
    import graph_marl_b200._lib as L

    name, cfg = netmon_case(G, entry)
    X = torch.from_numpy(G["node_obs"]).cuda()
    ADJ = torch.from_numpy(G["node_adj"]).float().cuda()
    nm = _netmon(cfg, X.shape[-1], math).eval()
    assert nm.get_out_features() == 4 * cfg["hidden"]
    eye = torch.eye(X.shape[1]).repeat(X.shape[0], 1, 1).cuda()  # sl.py:155
    n0 = L.lib().gm_kernel_launch_count()
    with torch.no_grad():
        nm.state = None
        for t in range(SEQ):
            out = nm(X, ADJ, eye)
            if t == 0:
                assert np.abs(out[:4].cpu().numpy() - G[name + "_out_first"]).max() < tol
    assert L.lib().gm_kernel_launch_count() > n0  # the CUDA path ran
    assert out.shape == G[name + "_out_last"].shape
    assert np.abs(out.cpu().numpy() - G[name + "_out_last"]).max() < tol
    assert np.abs(nm.state[:4].cpu().numpy() - G[name + "_state_last"]).max() < tol


@pytest.mark.gpu
def test_gpu_training_step_through_the_autograd_path():
    """sl.py:366-428 trains through NetMon with autograd recording: the learner path must give the reference's
    forward values, gradients for every parameter, and leave the kernel path consistent after an update."""
    import torch

    name, cfg = netmon_case(G, CASES[1])
    X = torch.from_numpy(G["node_obs"]).cuda()
    ADJ = torch.from_numpy(G["node_adj"]).float().cuda()
    nm = _netmon(cfg, X.shape[-1], "bf16x3").train()
    eye = torch.eye(X.shape[1]).repeat(X.shape[0], 1, 1).cuda()
    head = torch.nn.Linear(nm.get_out_features(), 1).cuda()
    tgt = torch.from_numpy(G["targets"].astype(np.float32)).cuda().reshape(-1, 1)
    opt = torch.optim.SGD(list(nm.parameters()) + list(head.parameters()), lr=1e-3)
    nm.state = None
    loss = 0
    for t in range(SEQ):
        out = nm(X, ADJ, eye)
        assert out.requires_grad
        if t == 0:
            assert np.abs(out[:4].detach().cpu().numpy() - G[name + "_out_first"]).max() < 1e-4
        loss = loss + torch.nn.functional.mse_loss(head(out).reshape(-1, 1), tgt)
    assert np.abs(out.detach().cpu().numpy() - G[name + "_out_last"]).max() < 1e-4
    opt.zero_grad()
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in nm.parameters())
    opt.step()
    # after the update the inference kernels must see the new weights (packed-weight cache) and agree with autograd
    nm.state = None
    a = nm(X, ADJ, eye).detach()
    nm.eval()
    with torch.no_grad():
        nm.state = None
        b = nm(X, ADJ, eye)
    assert (a - b).abs().max().item() < 1e-4
    assert (b.cpu() - torch.from_numpy(G[name + "_out_first"][:1])).abs().max().item() > 0  # weights did move


@pytest.mark.gpu
def test_gpu_compat_env_reproduces_the_dataset_observations():
    """sl.py:588-592 + build_dataset: Routing(network, 20, INDEPENDENT).reset() per sample, then
    get_node_observation() / get_nodes_adjacency() as numpy arrays of the reference's shapes."""
    from graph_marl_b200.env.constants import EVAL_SEEDS
    from graph_marl_b200.env.network import Network
    from graph_marl_b200.env.routing import Routing

    np.random.seed(5)
    net = Network(20, random_topology=True, sequential_topology_seeds=True, provided_seeds=EVAL_SEEDS)
    env = Routing(net, 20, 1)
    for s in range(G["node_obs"].shape[0]):
        env.reset()
        x = env.get_node_observation()
        assert x.dtype == np.float32 and np.array_equal(x, G["node_obs"][s]), s
        assert np.array_equal(env.get_nodes_adjacency(), G["node_adj"][s])
        assert np.array_equal(env.get_node_agent_matrix().sum(axis=0), np.ones(20))
