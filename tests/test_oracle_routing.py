"""The C oracle's Routing env against trajectories recorded from the reference."""
import hashlib

import numpy as np
import pytest

from conftest import ROUTING_CASES, load_golden
from helpers import cfg_from_golden, draws_for_step, topo_from_golden
from oracle import oracle as O


def _sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def _check_state(env, g, t):
    for k, a in (("now", env.now), ("target", env.target), ("edge", env.edge), ("time", env.time),
                 ("ttl", env.ttl_left), ("spw", env.spw), ("start", env.start), ("size", env.size),
                 ("load", env.load), ("agent_steps", env.agent_steps), ("visited", env.visited)):
        assert np.array_equal(a[0], g["s_" + k][t]), (k, t)


@pytest.mark.parametrize("case", ROUTING_CASES)
def test_routing_trajectory(case):
    g = load_golden(case)
    c = cfg_from_golden(g)
    A = c["n_data"]
    env = O.RoutingOracle(topo_from_golden(g), A, env_var=c["env_var"], k=c["k"],
                          enable_congestion=c["congestion"], enable_action_mask=c["mask"],
                          ttl=c["ttl"], eval_info=c["eval_info"])
    ds, dt, dz, n = draws_for_step(g, -1, A)
    assert n == A
    env.reset(ds, dt, dz)
    _check_state(env, g, 0)
    o = env.observe()
    n_dense = g["obs"].shape[0]
    assert np.array_equal(o["obs"][0], g["obs"][0])
    assert np.array_equal(o["adj"][0], g["adj"][0])
    assert np.array_equal(o["node_obs"][0], g["node_obs"][0])
    assert np.array_equal(o["node_agent"][0], g["node_agent"][0])
    assert np.array_equal(_sha(o["obs"][0], o["adj"][0], o["node_obs"][0], o["node_agent"][0]), g["hash"][0])
    for t in range(c["steps"]):
        ds, dt, dz, n = draws_for_step(g, t, A)
        if c["eval_info"]:
            r = env.step_single_evalinfo(g["actions"][t], ds, dt, dz)
            assert np.allclose(r["extra"], g["eval_extra"][t], rtol=0, atol=0)
        else:
            r = env.step(g["actions"][t], ds, dt, dz)
        assert r["n_resets"][0] == n
        assert np.array_equal(r["reward"][0], g["reward"][t]) and r["reward"].dtype == np.float32
        assert np.array_equal(r["done"][0], g["done"][t])
        assert np.array_equal(r["info"][0], g["info"][t].astype(np.int32))
        assert np.array_equal(r["delays"][0], g["delays"][t])
        assert np.array_equal(r["arrived"][0], g["arrived"][t])
        assert np.array_equal(r["spr"][0], g["spr"][t])
        _check_state(env, g, t + 1)
        if c["mask"]:
            assert np.array_equal(env.mask[0], g["s_mask"][t + 1])
        o = env.observe()
        assert np.array_equal(_sha(o["obs"][0], o["adj"][0], r["reward"][0], r["done"][0].astype(bool),
                                   o["node_obs"][0], o["node_agent"][0]), g["hash"][t + 1]), t
        if t + 1 < n_dense:
            assert np.array_equal(o["obs"][0], g["obs"][t + 1])
            assert np.array_equal(o["node_obs"][0], g["node_obs"][t + 1])
    if c["eval_info"]:
        assert np.array_equal(env.sum_packets_per_node[0], g["sum_packets_per_node"])
        assert np.array_equal(env.sum_packets_per_edge[0], g["sum_packets_per_edge"])
    final = env.agent_steps[0][env.agent_steps[0] != 0]
    assert np.array_equal(final, g["final_delays"].astype(np.int32))


def test_survey_anchor():
    """sum reward 257.6, throughput 37, blocked 562; first packets."""
    g = load_golden("routing_A_seed923430603_cong")
    assert abs(float(g["reward"].sum()) - 257.6) < 1e-3
    assert g["info"][:, 1].sum() == 37 and g["info"][:, 3].sum() == 562
    assert g["s_now"][0][:4].tolist() == [12, 3, 19, 6]
    assert g["s_target"][0][:4].tolist() == [15, 3, 18, 12]
    assert abs(g["s_size"][0][0] - 0.715189) < 1e-6


def test_batched_oracle_matches_single():
    g = load_golden("routing_B_nocong")
    c = cfg_from_golden(g)
    A = c["n_data"]
    B = 5
    env = O.RoutingOracle(topo_from_golden(g), A, enable_congestion=False, num_envs=B, threads=3)
    ds, dt, dz, _ = draws_for_step(g, -1, A)
    env.reset(ds, dt, dz)
    for t in range(20):
        ds, dt, dz, _ = draws_for_step(g, t, A)
        r = env.step(g["actions"][t], ds, dt, dz)
        for b in range(B):
            assert np.array_equal(r["reward"][b], g["reward"][t])
    o = env.observe()
    for b in range(B):
        assert np.array_equal(o["obs"][b], g["obs"][20])


def test_simple_env():
    g = load_golden("simple_env")
    for var in (1, 3):
        for rt in (0, 1):
            tag = f"v{var}_rt{rt}_"
            rng = O.MT19937(10 + var + rt)
            for ep in range(len(g[tag + "act"])):
                net = O.simple_build(rng, rt)
                assert np.array_equal(net["scores"], g[tag + "scores"][ep])
                assert np.array_equal(net["edges"], g[tag + "edges"][ep])
                assert [net["start_node"]] + net["start_edges"].tolist() == g[tag + "start_edges"][ep].tolist()
                assert O.simple_step(net, g[tag + "act"][ep]) == g[tag + "reward"][ep][0]
                assert g[tag + "obs"][ep][0, 0] == net["start_node"]
            assert rng.random() == g[tag + "after"][0]
