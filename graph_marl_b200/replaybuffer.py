"""Device-resident replay ring, mirrors src/replaybuffer.py:9-287 of the reference.

Same 17 fields, dtypes and ring arithmetic (`index = (index+1) % size`, `count` saturates),
same `get_batch` index streams (`np.random.default_rng(seed).choice`, contiguous sequences
anchored at the oldest element), but the arrays live in HBM: `add()` takes one transition
(numpy or tensors, reference call shape) or a batch of B transitions (tensors with a leading
env dimension) and lands them with one fused copy kernel; `get_batch()` gathers with one
kernel that also performs the reference's dtype conversions.
"""
import ctypes as C
from collections import namedtuple
from typing import Iterator

import numpy as np
import torch

from . import _lib

TransitionBatch = namedtuple(
    "TransitionBatch",
    ["idx", "obs", "action", "reward", "next_obs", "adj", "next_adj", "done", "episode_done", "agent_state",
     "node_obs", "node_adj", "node_state", "node_aux", "node_agent_matrix", "next_node_obs", "next_node_adj",
     "next_node_agent_matrix"],
)

# add() argument order of the reference (replaybuffer.py:243-262)
ADD_ORDER = ["obs", "action", "reward", "next_obs", "adj", "next_adj", "done", "episode_done", "agent_state",
             "node_state", "node_aux", "node_obs", "node_adj", "node_agent_matrix", "next_node_obs",
             "next_node_adj", "next_node_agent_matrix"]


class ReplayBuffer(object):
    def __init__(self, seed, buffer_size, n_agents, observation_size, agent_state_size, n_nodes=0,
                 node_observation_size=0, node_state_size=0, node_aux_size=0, half_precision=False,
                 device="cuda"):
        self.buffer_size = int(buffer_size)
        self.count = 0
        self.index = 0
        self.device = torch.device(device)
        ft = torch.float16 if half_precision else torch.float32
        A, N = n_agents, n_nodes
        shapes = dict(
            obs=((A, observation_size), ft), action=((A,), torch.int8), reward=((A,), ft),
            next_obs=((A, observation_size), ft), adj=((A, A), torch.bool), next_adj=((A, A), torch.bool),
            done=((A,), torch.bool), episode_done=((), torch.bool), agent_state=((A, agent_state_size), ft),
            node_state=((N, node_state_size), ft), node_aux=((N, node_aux_size), ft),
            node_obs=((N, node_observation_size), ft), next_node_obs=((N, node_observation_size), ft),
            node_adj=((N, N), torch.bool), next_node_adj=((N, N), torch.bool),
            node_agent_matrix=((N, A), torch.bool), next_node_agent_matrix=((N, A), torch.bool))
        self._shapes = shapes
        for name, (shape, dt) in shapes.items():
            setattr(self, name, torch.zeros((self.buffer_size, *shape), dtype=dt, device=self.device))
        self._consts = {}
        self._dev_index = None  # _lib.DeviceCounter in CUDA-graph mode (rollout.Rollout)
        self._random_generator = np.random.default_rng(seed)
        # dtype conversion of _get_transition_batch (replaybuffer.py:132-187)
        self._out_dtype = {n: torch.float32 for n in shapes}
        self._out_dtype.update(action=torch.int64, done=torch.bool, episode_done=torch.bool)

    # ---- insert ------------------------------------------------------------------------------
    def _source(self, name, value, n):
        """(tensor, convert, broadcast) for one ring field without staging copies where possible:
        0/1 int8/uint8 tensors are reinterpreted as bool, int32 actions are narrowed by the insert
        kernel, python scalars / expanded (stride-0) tensors are written as one broadcast transition."""
        shape, dt = self._shapes[name]
        numel = int(np.prod(shape, dtype=np.int64))
        if isinstance(value, (bool, int, float)):
            key = (name, value)
            if key not in self._consts:  # e.g. the scalar 0 for absent states (main.py:659,695), episode_done
                self._consts[key] = torch.full((1, *shape), value, dtype=dt, device=self.device)
            return self._consts[key], 0, 1
        t = torch.as_tensor(value)
        if name == "agent_state" and t.dim() == len(shape) + 1 and n == 1 and t.shape[0] == 1:
            t = t.squeeze(0)  # replaybuffer.py:272-273
        if t.device != self.device:
            t = t.to(self.device)
        bcast = 0
        if t.dim() == len(shape) + 1 and t.shape[0] == n and n > 1 and t.stride(0) == 0:
            t, bcast = t[:1], 1  # expanded per-topology tensor (node_adj, node_aux): one copy for all envs
        rows = 1 if bcast else n
        convert = 0
        if t.dtype != dt:
            if dt == torch.bool and t.dtype in (torch.int8, torch.uint8):
                t = t.contiguous().view(torch.bool)  # env outputs are exactly 0/1
            elif dt == torch.int8 and t.dtype == torch.int32 and name == "action":
                convert = 4
            else:
                t = (t != 0) if dt == torch.bool else t.to(dt)
        if t.numel() != rows * numel:
            t = torch.broadcast_to(t, (rows, *shape))
        return t.reshape(rows, *shape).contiguous(), convert, bcast

    def add(self, obs, action, reward, next_obs, adj, next_adj, done, episode_done, agent_state, node_state,
            node_aux, node_obs, node_adj, node_agent_matrix, next_node_obs, next_node_adj,
            next_node_agent_matrix, num=1):
        """One transition (reference call, replaybuffer.py:243-287) or `num` transitions whose
        arguments carry a leading dimension of size num (batched rollout).  `obs` / `next_obs` may be
        (agent_obs, graph_obs) pairs: both parts land in the joint ring row without a concat."""
        _lib.require_device()
        vals = dict(zip(ADD_ORDER, (obs, action, reward, next_obs, adj, next_adj, done, episode_done, agent_state,
                                    node_state, node_aux, node_obs, node_adj, node_agent_matrix, next_node_obs,
                                    next_node_adj, next_node_agent_matrix)))
        n = int(num)
        assert n <= self.buffer_size
        fields = (_lib.ReplayField * _lib.GM_REPLAY_MAX_FIELDS)()
        keep = []
        k = 0
        for name in ADD_ORDER:
            ring = getattr(self, name)
            eb = ring[0].numel() * ring.element_size()
            if eb == 0:
                continue
            v = vals[name]
            if name in ("obs", "next_obs") and isinstance(v, (tuple, list)):
                # split joint observation: column blocks of the [A, D] ring element
                (A, D), dt = self._shapes[name]
                col = 0
                for part in v:
                    w = part.shape[-1]
                    t = part if part.dtype == dt else part.to(dt)
                    t = t.reshape(n * A, w).contiguous()
                    keep.append(t)
                    f = fields[k]
                    f.ring, f.src, f.elem_bytes = ring.data_ptr(), t.data_ptr(), eb
                    f.rows, f.row_bytes = A, w * ring.element_size()
                    f.ring_pitch, f.ring_offset = D * ring.element_size(), col * ring.element_size()
                    col += w
                    k += 1
                assert col == D, f"{name}: parts cover {col} of {D} columns"
                continue
            src, convert, bcast = self._source(name, v, n)
            keep.append(src)
            f = fields[k]
            f.ring, f.src, f.elem_bytes, f.convert, f.broadcast = ring.data_ptr(), src.data_ptr(), eb, convert, bcast
            k += 1
        index, index_dev = self.index, None
        if self._dev_index is not None:
            index, index_dev = self._dev_index.offset(self.index) % self.buffer_size, self._dev_index.ptr()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().gm_replay_insert(fields, k, self.buffer_size, index, index_dev, n, _lib.current_stream()))
            cur = torch.cuda.current_stream()
            if cur != torch.cuda.default_stream() and not torch.cuda.is_current_stream_capturing():
                # insert running on a side stream (overlapped with the next rollout step): keep the caching
                # allocator from recycling the sources before the copy has run
                for t in keep:
                    t.record_stream(cur)
        self._keep = keep
        self.count = min(self.buffer_size, self.count + n)
        self.index = (self.index + n) % self.buffer_size

    # ---- sample ------------------------------------------------------------------------------
    def get_batch(self, batch_size, device, sequence_length=0) -> Iterator[TransitionBatch]:
        if sequence_length <= 1:
            indices = self._random_generator.choice(self.count, batch_size, replace=True, p=None)
            yield self._get_transition_batch(indices, device)
            return
        buffer_start = self.index % self.count
        batch_sequence_start = self._random_generator.choice(self.count - sequence_length, batch_size,
                                                             replace=True, p=None)
        batch_sequence_start = (buffer_start + batch_sequence_start) % self.count
        for offset in range(sequence_length):
            indices = (batch_sequence_start + offset) % self.count
            yield self._get_transition_batch(indices, device)

    def _get_transition_batch(self, indices, device) -> TransitionBatch:
        _lib.require_device()
        n = len(indices)
        idx = torch.as_tensor(np.ascontiguousarray(indices, dtype=np.int64)).to(self.device)
        names = [f for f in TransitionBatch._fields if f != "idx"]
        fields = (_lib.ReplayField * len(names))()
        outs = {}
        k = 0
        for name in names:
            ring = getattr(self, name)
            shape, dt = self._shapes[name]
            od = self._out_dtype[name]
            out = torch.empty((n, *shape), dtype=od, device=self.device)
            outs[name] = out
            eb = ring[0].numel() * ring.element_size()
            if eb == 0:
                continue
            if od == torch.float32 and dt == torch.bool:
                conv = 1
            elif od == torch.int64:
                conv = 2
            elif od == torch.float32 and dt == torch.float16:
                conv = 3
            else:
                conv = 0
            fields[k].ring, fields[k].dst, fields[k].elem_bytes, fields[k].convert = (
                ring.data_ptr(), out.data_ptr(), eb, conv)
            k += 1
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().gm_replay_sample(fields, k, idx.data_ptr(), n, _lib.current_stream()))
        dev = torch.device(device)
        return TransitionBatch(indices, *[outs[f].to(dev, non_blocking=True) for f in names])

    def get_recent_indices(self, last_n):
        if last_n is None:
            return np.arange(self.count), np.arange(self.count)
        m = min(self.count, last_n)
        x = self.index - m + np.arange(m)
        return x, x % self.count
