"""SimpleEnvironment, mirrors src/env/simple_environment.py:45-334 of the reference over the
batched CUDA kernel in csrc/simple_env.cu (gm_simple_step).

Three nodes in a line, one agent on the middle node, border scores {-1, +1}; the binary action
picks one of the two edges of the start node, the reward is the score of the node reached, the
episode is always done and the packet returns to the start node (simple_environment.py:289-315).

* num_envs == 1 (default): numpy in / numpy out with the reference's shapes and dtypes; the
  per-episode topology is drawn from the global legacy `np.random` stream with exactly the
  reference's draw order (`_build_network`, :106-187).
* num_envs  > 1: B independent instances (each with its own topology drawn from the same host
  stream at `reset()`), CUDA tensors with a leading num_envs dimension.
"""
import textwrap

import numpy as np
import torch

from .. import _lib
from .environment import EnvironmentVariant, NetworkEnv
from .routing import Discrete


class SimpleEnvironment(NetworkEnv):
    def __init__(self, env_var, random_topology, num_envs=1, device=None, batched=None):
        self.n_router = 3
        self.n_data = 1
        self.record_distance_map = False
        self.random_topology = random_topology
        self.sort_edges = True
        self.env_var = EnvironmentVariant(env_var)
        self.start_node = -1
        self.action_space = Discrete(2, start=0)  # {0, 1}
        self.adj_matrix = None
        self.num_envs = int(num_envs)
        self.batched = (self.num_envs > 1) if batched is None else bool(batched)
        self.device = torch.device(device if device is not None else "cuda")
        self._out = {}
        self._tables = None
        self._host = None

    def get_num_agents(self):
        return self.n_data

    def get_num_nodes(self):
        return self.n_router

    def __str__(self) -> str:
        return textwrap.dedent(
            f"""\
            SimpleEnvironment with parameters
            > Environment variant: {self.env_var}
            > Random topology: {self.random_topology}
            > Instances: {self.num_envs} on {self.device} (libgraphmarl_b200)\
            """
        )

    # ---- topology (host, global np.random stream; simple_environment.py:106-187) ---------------
    def _draw_network(self):
        border_scores = np.array([-1, 1])
        np.random.shuffle(border_scores)  # border scores are always shuffled
        scores = np.array([border_scores[0], 0, border_scores[1]])
        if self.random_topology:
            np.random.shuffle(scores)
        n0 = int(np.where(scores == 0)[0][0])
        n1 = (n0 + 1) % 3
        n2 = (n1 + 1) % 3
        for _ in range(3):  # router positions (:131-134), only used for rendering
            np.random.random(), np.random.random()
        edge_destinations = [n1, n2]
        if self.random_topology:
            np.random.shuffle(edge_destinations)
        edges = []
        for k in range(2):
            edge_nodes = [n0, edge_destinations[k]]
            if self.random_topology:
                np.random.shuffle(edge_nodes)
            edges.append(list(edge_nodes))
        edge_order = [0, 1]
        if self.random_topology:
            if self.sort_edges:
                edge_order = [int(x) for x in np.argsort(edge_destinations)]
            else:
                np.random.shuffle(edge_order)
        return dict(scores=scores.astype(np.int32), edges=np.array(edges, np.int32), start_node=n0,
                    start_edges=np.array(edge_order, np.int32), neighbors=(n1, n2))

    def _build_network(self):
        nets = [self._draw_network() for _ in range(self.num_envs)]
        self._host = nets
        st = lambda k, dt: torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(n[k]) for n in nets]).astype(dt))).to(self.device)
        self._tables = dict(scores=st("scores", np.int32), edges=st("edges", np.int32),
                            start_node=st("start_node", np.int32), start_edges=st("start_edges", np.int32))
        n = nets[0]
        self.start_node = n["start_node"]
        self.adj_matrix = np.eye(3, 3, dtype=np.int8)
        for nb in n["neighbors"]:
            self.adj_matrix[n["start_node"]][nb] = 1
            self.adj_matrix[nb][n["start_node"]] = 1
        # adjacency lists (ascending ids incl. self, -1 padded) for the NetMon list path
        lists = np.full((self.num_envs, 3, 3), -1, np.int32)
        deg = np.zeros((self.num_envs, 3), np.int32)
        for b, nn in enumerate(nets):
            adj = np.eye(3, dtype=bool)
            for nb in nn["neighbors"]:
                adj[nn["start_node"], nb] = adj[nb, nn["start_node"]] = True
            for v in range(3):
                ids = np.nonzero(adj[v])[0]
                lists[b, v, : len(ids)] = ids
                deg[b, v] = len(ids)
        self._lists = (torch.from_numpy(lists).to(self.device), torch.from_numpy(deg).to(self.device))

    def _launch(self, actions):
        _lib.require_device()
        B, dev = self.num_envs, self.device
        W = 1 if self.env_var == EnvironmentVariant.INDEPENDENT else 13
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
        out = dict(obs=e((B, 1, W), torch.float32), node_obs=e((B, 3, 1), torch.float32),
                   node_agent=e((B, 3, 1), torch.int8), node_adj=e((B, 3, 3), torch.int8))
        reward = e((B,), torch.float32) if actions is not None else None
        t = self._tables
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gm_simple_step(
                B, self.env_var.value, t["scores"].data_ptr(), t["edges"].data_ptr(), t["start_node"].data_ptr(),
                t["start_edges"].data_ptr(), _lib.ptr(actions), out["obs"].data_ptr(), out["node_obs"].data_ptr(),
                out["node_agent"].data_ptr(), out["node_adj"].data_ptr(), _lib.ptr(reward), _lib.current_stream()))
        out["agent_node"] = t["start_node"].reshape(B, 1)
        self._out = out
        self._keep = actions
        return out, reward

    def _ret(self, t):
        return t if self.batched else t[0].cpu().numpy()

    # ---- reference API ---------------------------------------------------------------------------
    def reset(self):
        self._build_network()
        out, _ = self._launch(None)
        return self._ret(out["obs"]), self._data_adjacency()

    def step(self, action):
        B = self.num_envs
        if not torch.is_tensor(action):
            action = torch.as_tensor(np.ascontiguousarray(np.asarray(action, dtype=np.int32)))
        act = action.to(device=self.device, dtype=torch.int32).reshape(B).contiguous()
        out, reward = self._launch(act)
        if self.batched:
            done = torch.ones((B, 1), dtype=torch.bool, device=self.device)
            return out["obs"], self._data_adjacency(), reward.reshape(B, 1), done, {}
        # reference returns np.array([score]) (int64) and the python list [True]
        return (self._ret(out["obs"]), self._data_adjacency(), np.array([int(reward[0].item())]), [True], {})

    def _data_adjacency(self):
        if self.batched:
            return torch.ones((self.num_envs, 1, 1), dtype=torch.int8, device=self.device)
        return np.eye(1, 1, dtype=np.int8)

    def render(self):
        raise NotImplementedError("rendering is out of the hot path")

    def get_nodes_adjacency(self):
        return self._out["node_adj"] if self.batched else self.adj_matrix

    def get_node_observation(self):
        return self._ret(self._out["node_obs"])

    def get_node_agent_matrix(self):
        return self._ret(self._out["node_agent"])

    # ---- hooks used by NetMonWrapper's list path -------------------------------------------------------
    def get_agent_nodes(self):
        return self._out["agent_node"]

    def get_adjacency_lists(self):
        return self._lists[0], self._lists[1], None
