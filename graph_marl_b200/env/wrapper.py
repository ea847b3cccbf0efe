"""NetMonWrapper, mirrors src/env/wrapper.py:7-109 of the reference.

The reference pulls node observations / adjacency / node-agent matrix out of the env as
numpy, uploads them, runs NetMon, maps node outputs to agents with a bmm and downloads the
result (3 H2D + 1 D2H per step).  Here the env's outputs are already on the device, the
adjacency is a list table of the topology pool, and the node->agent mapping is a gather by
the agents' node index, so the whole step stays in HBM.
"""
import os

import numpy as np
import torch


class NetMonWrapper:
    def __init__(self, env, netmon, startup_iterations, split_obs=False, graph_obs_fp32=True) -> None:
        if startup_iterations < 1:
            raise AssertionError("Number of startup iterations must be >= 1")
        self.env, self.netmon = env, netmon
        self.startup_iterations = startup_iterations
        self.split_obs = split_obs  # batched mode: return (agent_obs, graph_obs) instead of the concat
        # batched split mode on the tensor-core path: False = the graph observation is produced once, tile-packed
        # (model.PackedRows), for a DQN of the same math mode; nothing writes or reads its fp32 rows
        self.graph_obs_fp32 = graph_obs_fp32
        self.device = next(netmon.parameters()).device
        # attributes main.py / sl.py read (wrapper.py:13-19)
        self.node_obs = self.node_adj = self.node_agent_matrix = None
        self.last_netmon_state = self.current_netmon_state = None
        self.frozen = False
        self.netmon_out = None
        # batched rollouts: where the NEXT NetMon step writes its new state (a block of the replay ring's node_state
        # field, rollout.Rollout); consumed by that step
        self._state_sink = None

    def __getattr__(self, name):
        # anything the wrapper does not define is the env's (wrapper.py:25-26)
        env = self.__dict__.get("env")
        if env is None:  # not constructed yet (copy / pickle probes)
            raise AttributeError(name)
        return getattr(env, name)

    def __str__(self) -> str:
        return f"{self.env}{os.linesep}▲ environment is wrapped with NetMon (graph obs)"

    @property
    def _batched(self):
        return getattr(self.env, "batched", False)

    def _with_graph_obs(self, obs):
        """One NetMon pass on the env's current node view, appended to the agents' observations."""
        graph_obs = self._netmon_step()
        if not self._batched:
            return np.concatenate((obs, graph_obs), axis=-1)
        return (obs, graph_obs) if self.split_obs else torch.cat((obs, graph_obs), dim=-1)

    def reset(self):
        self.frozen = False
        self.last_netmon_state = self.current_netmon_state = None
        obs, adj = self.env.reset()
        sink, self._state_sink = self._state_sink, None
        for _ in range(self.startup_iterations - 1):  # warm-up passes; the last one below builds the obs
            self._netmon_step()
        self._state_sink = sink
        return self._with_graph_obs(obs), adj

    def step(self, actions):
        obs, adj, reward, done, info = self.env.step(actions)
        return self._with_graph_obs(obs), adj, reward, done, info

    def freeze(self):
        """Disable message passing for the rest of the episode (wrapper.py:53-58).  The frozen path serves
        `netmon_out` (the node readout of the latest NetMon step) at the agents' new nodes; the lean batched mode
        only reads out the agents' rows, so the node readout is rebuilt here by repeating that step from the state
        before it (same kernels, same inputs: the state it reproduces is the current one)."""
        if not self.frozen and self.netmon_out is None and self.current_netmon_state is not None:
            env = self.env
            nbr_all, deg, list_index = env.get_adjacency_lists()
            with torch.no_grad():
                keep = self.netmon.state
                self.netmon.state = self.last_netmon_state
                self.netmon_out, _ = self.netmon.forward_lists(env._out["node_obs"], nbr_all, deg, list_index,
                                                               nbr_all.shape[-1] - 1, want_node_out=True,
                                                               sparse_nnz=getattr(env, "node_obs_nnz", 0),
                                                               sparse_rows=env._out.get("node_sparse"),
                                                               static_rows=getattr(env, "node_static_rows", None))
                self.netmon.state = keep
        self.frozen = True

    def get_netmon_info(self):
        if self._batched:
            return (self.env._out["node_obs"], self.env.get_nodes_adjacency(), self.env._out["node_agent"])
        return (self.node_obs, self.node_adj, self.node_agent_matrix)

    def get(self):
        return self.env

    def _ret(self, t):
        return t if self._batched else t[0].cpu().numpy()

    def _netmon_step(self):
        env = self.env
        if hasattr(env, "get_adjacency_lists"):
            nbr_all, deg, list_index = env.get_adjacency_lists()
            agent_node = env.get_agent_nodes()
            node_obs = env._out["node_obs"]
            max_degree = nbr_all.shape[-1] - 1
        else:
            raise TypeError("NetMonWrapper needs a graph_marl_b200 environment")
        if not self._batched:
            self.node_agent_matrix = env.get_node_agent_matrix()
        if self.frozen:
            # wrapper.py:67-75: agents keep reading the frozen node outputs at their new position
            if self.netmon_out is None:
                raise RuntimeError("freeze() before the first NetMon step")
            idx = agent_node.long().unsqueeze(-1).expand(-1, -1, self.netmon_out.shape[-1])
            return self._ret(torch.gather(self.netmon_out, 1, idx))
        if not self._batched:
            self.node_obs = env.get_node_observation()
            self.node_adj = env.get_nodes_adjacency()
        with torch.no_grad():
            self.last_netmon_state = self.current_netmon_state
            self.netmon.state = self.current_netmon_state
            # batched split mode: only the agents' rows are read out (plus their tile-packed copy for
            # the tensor-core DQN); the full node readout is kept for compat / freeze()
            lean = self._batched and self.split_obs
            self.netmon_out, agent_out = self.netmon.forward_lists(
                node_obs, nbr_all, deg, list_index, max_degree, agent_node=agent_node, want_node_out=not lean,
                want_agent_pk=lean, want_agent_fp32=self.graph_obs_fp32 or not lean, state_out=self._state_sink,
                sparse_nnz=getattr(env, "node_obs_nnz", 0), sparse_rows=env._out.get("node_sparse"),
                static_rows=getattr(env, "node_static_rows", None))
            self._state_sink = None
            self.current_netmon_state = self.netmon.state
        return self._ret(agent_out)
