"""Environment interface, mirrors src/env/environment.py:7-131 of the reference."""
import abc
from enum import Enum
from typing import Any, Dict


class EnvironmentVariant(Enum):
    INDEPENDENT = 1       # without neighbour info in obs
    WITH_K_NEIGHBORS = 2  # with info of k neighbours in obs
    GLOBAL = 3            # with the global topology and all node observations in obs


def reset_and_get_sizes(env):
    """(n_agents, obs_dim, n_nodes, node_obs_dim) from a live reset (environment.py:16-33).
    Works for single-env (2-d) and batched (3-d, leading num_envs) observations."""
    agent_observation, _ = env.reset()
    node_obs = env.get_node_observation()
    return (agent_observation.shape[-2], agent_observation.shape[-1], node_obs.shape[-2], node_obs.shape[-1])


class NetworkEnv(abc.ABC):
    """Abstract graph/network environment (environment.py:36-131)."""

    @abc.abstractmethod
    def reset(self):
        ...

    @abc.abstractmethod
    def step(self, act):
        ...

    def get(self):
        return self

    def get_final_info(self, info: Dict[str, Any]):
        return info

    def get_node_aux(self):
        return None

    @abc.abstractmethod
    def get_node_agent_matrix(self):
        ...

    @abc.abstractmethod
    def get_nodes_adjacency(self):
        ...

    @abc.abstractmethod
    def get_node_observation(self):
        ...

    @abc.abstractmethod
    def get_num_agents(self):
        ...

    @abc.abstractmethod
    def get_num_nodes(self):
        ...
