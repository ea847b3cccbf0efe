"""The environment contract of the rollout path: what `Routing`, `SimpleEnvironment` and `NetMonWrapper`
implement and what main.py / eval.py call (interface of src/env/environment.py:7-131; own text).

Shapes below are the single-env ones of the reference; in batched mode (`num_envs > 1`) every array is a
CUDA tensor with a leading `num_envs` dimension.
"""
from abc import ABC, abstractmethod
from enum import Enum, unique


@unique
class EnvironmentVariant(Enum):
    INDEPENDENT = 1       # agent obs carries no neighbour info
    WITH_K_NEIGHBORS = 2  # agent obs carries the obs of k neighbours
    GLOBAL = 3            # agent obs carries the topology and every node's observation


def reset_and_get_sizes(env):
    """(n_agents, obs_dim, n_nodes, node_obs_dim), read off a live reset (environment.py:16-33).
    The trailing two dims are used so single-env (2-d) and batched (3-d) observations both work."""
    agent_obs = env.reset()[0]
    node_obs = env.get_node_observation()
    return tuple(agent_obs.shape[-2:]) + tuple(node_obs.shape[-2:])


class NetworkEnv(ABC):
    """Graph/network environment (environment.py:36-131)."""

    # -- episode control ----------------------------------------------------------------------
    @abstractmethod
    def reset(self):  # -> (agent obs [A, D], agent adjacency [A, A])
        raise NotImplementedError

    @abstractmethod
    def step(self, act):  # -> (agent obs, agent adjacency, reward [A], done [A], info dict)
        raise NotImplementedError

    # -- node-level view consumed by NetMon ------------------------------------------------------
    @abstractmethod
    def get_node_observation(self):  # [N, node_obs_dim], node_obs_dim constant over an env's lifetime
        raise NotImplementedError

    @abstractmethod
    def get_nodes_adjacency(self):  # [N, N]
        raise NotImplementedError

    @abstractmethod
    def get_node_agent_matrix(self):  # [N, A], 1 where agent a sits on node n
        raise NotImplementedError

    @abstractmethod
    def get_num_agents(self):
        raise NotImplementedError

    @abstractmethod
    def get_num_nodes(self):
        raise NotImplementedError

    # -- optional hooks with defaults ------------------------------------------------------------
    def get_node_aux(self):
        """Auxiliary per-node targets [N, aux_dim]; None when the env has none."""
        return None

    def get_final_info(self, info):
        """End-of-episode additions to the step info dict (extended in place and returned)."""
        return info

    def get(self):
        """The innermost environment (wrappers forward this call)."""
        return self
