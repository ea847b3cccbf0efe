"""ctypes front-end of the CPU ORACLE (oracle/gm_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke()
and the cpu_baseline / --impl reference legs of bench.py, never by the product
package (graph_marl_b200 fails loudly without its CUDA library instead).

Parity status: PINNED -- every function here is checked in tests/test_oracle_*.py
against outputs of the unmodified reference recorded in tests/golden/ by
tools/gen_golden.py.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgm_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "gm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgm_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.gmo_rng_double.restype = C.c_double
        _lib.gmo_rng_u32.restype = C.c_uint32
        _lib.gmo_rng_interval.restype = C.c_uint32
        _lib.gmo_rng_randint.restype = C.c_uint32
        _lib.gmo_topology_create.restype = C.c_int64
    return _lib


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


class MT19937:
    """np.random legacy stream."""

    def __init__(self, seed=None):
        self._buf = C.create_string_buffer(lib().gmo_rng_sizeof())
        if seed is not None:
            self.seed(seed)

    def seed(self, s):
        lib().gmo_rng_seed(self._buf, C.c_uint32(int(s) & 0xFFFFFFFF))

    def random(self):
        return lib().gmo_rng_double(self._buf)

    def randint(self, high):
        return int(lib().gmo_rng_randint(self._buf, C.c_uint32(int(high))))

    def randint_vec(self, high, n):
        return np.array([self.randint(high) for _ in range(n)], dtype=np.int64)

    def rand_vec(self, n):
        return np.array([self.random() for _ in range(n)], dtype=np.float64)

    def shuffle(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.int32)
        lib().gmo_rng_shuffle(self._buf, _p(a), C.c_int(a.size))
        return a

    def choice(self, seq):
        return seq[self.randint(len(seq))]

    @property
    def pos(self):
        return lib().gmo_rng_pos(self._buf)

    def copy(self):
        other = MT19937()
        C.memmove(other._buf, self._buf, len(self._buf))
        return other


def generate_topology(n_nodes, seed=None, global_rng=None, exclude=()):
    """network.py:215-290. seed=None draws the seed from `global_rng`."""
    N, E = n_nodes, 3 * n_nodes // 2 + 8
    edges = np.zeros((E, 3), dtype=np.int32)
    node_edges = np.zeros((N, 3), dtype=np.int32)
    nbr = np.zeros((N, 3), dtype=np.int32)
    xy = np.zeros((N, 2), dtype=np.float64)
    rep = C.c_int32(0)
    ex = np.ascontiguousarray(np.array(sorted(exclude), dtype=np.int64))
    g = global_rng._buf if global_rng is not None else None
    used = lib().gmo_topology_create(
        g, C.c_int(N), C.c_int(0 if seed is None else 1), C.c_int64(0 if seed is None else int(seed)),
        _p(ex), C.c_int(ex.size), _p(edges), _p(node_edges), _p(nbr), _p(xy), C.byref(rep))
    if used < 0:
        raise AssertionError(f"Provided seed {seed} is invalid.")
    edges = edges[: 3 * N // 2].copy()
    apsp = np.zeros((N, N), dtype=np.int32)
    lib().gmo_apsp(C.c_int(N), C.c_int(edges.shape[0]), _p(edges), _p(apsp))
    adj = np.zeros((N, N), dtype=np.int8)
    lib().gmo_node_adj(C.c_int(N), C.c_int(edges.shape[0]), _p(edges), _p(adj))
    return dict(seed=int(used), repetitions=rep.value, edges=edges, node_edges=node_edges,
                nbr_creation=nbr, xy=xy, apsp=apsp, adj=adj)


class _Env(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("N", "A", "E", "env_var", "k", "congestion", "action_mask", "ttl", "eval_info", "VW")] + \
               [(n, C.c_void_p) for n in
                ("node_edges", "edges", "apsp", "now", "target", "edge", "time", "ttl_left", "spw",
                 "start", "agent_steps", "size", "visited", "load", "mask",
                 "sum_packets_per_node", "sum_packets_per_edge")]


class _Batch(C.Structure):
    _fields_ = [("proto", _Env), ("B", C.c_int32)]


class RoutingOracle:
    """B independent Routing envs (routing.py) sharing one topology. B=1 squeezes nothing:
    all arrays carry the leading batch dimension."""

    def __init__(self, topo, n_data, env_var=1, k=3, enable_congestion=True,
                 enable_action_mask=False, ttl=0, eval_info=False, num_envs=1, threads=1):
        self.topo = topo
        self.N = N = topo["apsp"].shape[0]
        self.A = A = n_data
        self.E = E = topo["edges"].shape[0]
        self.B = B = num_envs
        self.VW = VW = (N + 31) // 32
        self.env_var, self.k = env_var, k
        self.threads = threads
        self._edges = np.ascontiguousarray(topo["edges"], dtype=np.int32)
        self._node_edges = np.ascontiguousarray(topo["node_edges"], dtype=np.int32)
        self._apsp = np.ascontiguousarray(topo["apsp"], dtype=np.int32)
        i32 = lambda *s: np.zeros(s, dtype=np.int32)
        self.now, self.target, self.edge, self.time = i32(B, A), i32(B, A), i32(B, A), i32(B, A)
        self.ttl_left, self.spw, self.start, self.agent_steps = i32(B, A), i32(B, A), i32(B, A), i32(B, A)
        self.size = np.zeros((B, A), dtype=np.float64)
        self.visited = np.zeros((B, A, VW), dtype=np.uint32)
        self.load = np.zeros((B, E), dtype=np.float64)
        self.mask = np.zeros((B, A, 4), dtype=np.uint8)
        self.sum_packets_per_node = np.zeros((B, N), dtype=np.float64)
        self.sum_packets_per_edge = np.zeros((B, E), dtype=np.float64)
        e = _Env(N, A, E, env_var, k, int(enable_congestion), int(enable_action_mask), ttl,
                 int(eval_info), VW)
        for name, arr in (("node_edges", self._node_edges), ("edges", self._edges), ("apsp", self._apsp),
                          ("now", self.now), ("target", self.target), ("edge", self.edge),
                          ("time", self.time), ("ttl_left", self.ttl_left), ("spw", self.spw),
                          ("start", self.start), ("agent_steps", self.agent_steps), ("size", self.size),
                          ("visited", self.visited), ("load", self.load), ("mask", self.mask),
                          ("sum_packets_per_node", self.sum_packets_per_node),
                          ("sum_packets_per_edge", self.sum_packets_per_edge)):
            setattr(e, name, arr.ctypes.data)
        self._batch = _Batch(e, B)
        self.obs_width = lib().gmo_routing_obs_width(C.byref(e))
        self._pool = ThreadPoolExecutor(threads) if threads > 1 else None

    def _run(self, fn, *args):
        if self._pool is None:
            fn(C.byref(self._batch), C.c_int(0), C.c_int(self.B), *args)
            return
        step = (self.B + self.threads - 1) // self.threads
        futs = [self._pool.submit(fn, C.byref(self._batch), C.c_int(lo), C.c_int(min(self.B, lo + step)), *args)
                for lo in range(0, self.B, step)]
        for f in futs:
            f.result()

    @staticmethod
    def _draws(start, target, size, B, A):
        ds = np.ascontiguousarray(np.broadcast_to(np.asarray(start, dtype=np.int32), (B, A)))
        dt = np.ascontiguousarray(np.broadcast_to(np.asarray(target, dtype=np.int32), (B, A)))
        dz = np.ascontiguousarray(np.broadcast_to(np.asarray(size, dtype=np.float64), (B, A)))
        return ds, dt, dz

    def reset(self, start, target, size):
        ds, dt, dz = self._draws(start, target, size, self.B, self.A)
        self._run(lib().gmo_batch_reset, _p(ds), _p(dt), _p(dz))

    def step(self, actions, start, target, size):
        B, A = self.B, self.A
        act = np.ascontiguousarray(np.broadcast_to(np.asarray(actions, dtype=np.int32), (B, A)))
        ds, dt, dz = self._draws(start, target, size, B, A)
        out = dict(
            reward=np.zeros((B, A), np.float32), done=np.zeros((B, A), np.uint8),
            delays=np.zeros((B, A), np.int32), arrived=np.zeros((B, A), np.uint8),
            spr=np.zeros((B, A), np.float64), looped=np.zeros((B, A), np.uint8),
            info=np.zeros((B, 4), np.int32), n_resets=np.zeros(B, np.int32))
        self._run(lib().gmo_batch_step, _p(act), _p(ds), _p(dt), _p(dz), _p(out["reward"]),
                  _p(out["done"]), _p(out["delays"]), _p(out["arrived"]), _p(out["spr"]),
                  _p(out["looped"]), _p(out["info"]), _p(out["n_resets"]))
        return out

    def step_single_evalinfo(self, actions, start, target, size):
        """B == 1 only: also returns the eval-info extras (routing.py:414-441)."""
        assert self.B == 1
        A = self.A
        act = np.ascontiguousarray(actions, dtype=np.int32).reshape(A)
        ds, dt, dz = self._draws(start, target, size, 1, A)
        out = dict(
            reward=np.zeros((1, A), np.float32), done=np.zeros((1, A), np.uint8),
            delays=np.zeros((1, A), np.int32), arrived=np.zeros((1, A), np.uint8),
            spr=np.zeros((1, A), np.float64), looped=np.zeros((1, A), np.uint8),
            info=np.zeros((1, 4), np.int32), extra=np.zeros(4, np.float64),
            packet_dist=np.zeros(A, np.int32))
        n = lib().gmo_routing_step(C.byref(self._batch.proto), _p(act), _p(ds), _p(dt), _p(dz),
                                   _p(out["reward"]), _p(out["done"]), _p(out["delays"]),
                                   _p(out["arrived"]), _p(out["spr"]), _p(out["looped"]),
                                   _p(out["info"]), _p(out["extra"]), _p(out["packet_dist"]))
        out["n_resets"] = np.array([n], np.int32)
        return out

    def observe(self, obs=True, adj=True, node_obs=True, node_agent=True):
        B, A, N = self.B, self.A, self.N
        o = np.zeros((B, A, self.obs_width), np.float32) if obs else None
        a = np.zeros((B, A, A), np.int8) if adj else None
        n = np.zeros((B, N, 4 * N + 8), np.float32) if node_obs else None
        m = np.zeros((B, N, A), np.int8) if node_agent else None
        self._run(lib().gmo_batch_observe, _p(o), _p(a), _p(n), _p(m))
        return dict(obs=o, adj=a, node_obs=n, node_agent=m)


def simple_build(rng, random_topology):
    """simple_environment.py:106-187 driven by an MT19937 stream."""
    scores = np.zeros(3, np.int32)
    edges = np.zeros((2, 2), np.int32)
    start = C.c_int32(0)
    se = np.zeros(2, np.int32)
    lib().gmo_simple_build(rng._buf, C.c_int(int(random_topology)), _p(scores), _p(edges),
                           C.byref(start), _p(se))
    return dict(scores=scores, edges=edges, start_node=start.value, start_edges=se)


def simple_step(net, act):
    return lib().gmo_simple_step(_p(net["scores"]), _p(net["edges"]), C.c_int(net["start_node"]),
                                 _p(net["start_edges"]), C.c_int(int(act)))
