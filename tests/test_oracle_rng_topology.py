"""The C oracle's MT19937 + topology generator against reference outputs (tests/golden)."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle as O


def test_mt19937_known_answers():
    g = load_golden("rng")
    for i, s in enumerate(g["seeds"]):
        r = O.MT19937(int(s))
        assert np.array_equal([r.random() for _ in range(10)], g["random10"][i])
        assert np.array_equal([r.randint(20) for _ in range(16)], g["randint20"][i])
        assert np.array_equal([r.randint(200) for _ in range(16)], g["randint200"][i])
        assert np.array_equal(r.randint_vec(4, 9), g["randint4_vec"][i])
        assert np.array_equal(r.rand_vec(9), g["rand_vec"][i])
        assert np.array_equal([r.randint(2**31 - 1) for _ in range(6)], g["randint_big"][i])
        assert np.array_equal([r.randint(1) for _ in range(3)], g["randint1"][i])
        assert np.array_equal(r.shuffle(np.arange(7)), g["shuffle7"][i])
        assert np.array_equal(r.shuffle(np.array([5, 9])), g["shuffle_list2"][i])
        assert np.array_equal([r.choice([11, 22, 33, 44, 55]) for _ in range(5)], g["choice"][i])
        assert r.choice([77]) == g["choice1"][i][0]
        assert r.random() == g["after"][i][0]
        assert r.pos == g["state_pos"][i][0]


def test_fixed_seed_topologies():
    g = load_golden("topology")
    for i, s in enumerate(g["fixed_seeds"]):
        t = O.generate_topology(20, seed=int(s))
        assert t["repetitions"] == 1
        assert np.array_equal(t["edges"], g["fixed_edges"][i]), s
        assert np.array_equal(t["node_edges"], g["fixed_node_edges"][i])
        assert np.array_equal(t["nbr_creation"], g["fixed_node_nbrs_creation"][i])
        assert np.array_equal(t["apsp"], g["fixed_apsp"][i])
        assert np.array_equal(t["xy"], g["fixed_xy"][i])
    assert np.array_equal(t["adj"], g["fixed_adj"])
    # SURVEY 8c anchor
    t = O.generate_topology(20, seed=923430603)
    assert t["edges"][:3].tolist() == [[0, 3, 1], [0, 6, 2], [0, 17, 2]]
    assert t["apsp"][0].tolist() == [0, 11, 11, 1, 3, 5, 2, 6, 8, 14, 23, 22, 10, 12, 17, 7, 16, 2, 16, 21]


def test_all_eval_seeds_digest():
    g = load_golden("topology")
    assert len(g["eval_seeds"]) == 1000 and g["eval_seeds"][350] == 923430603
    for s, d in zip(g["eval_seeds"], g["eval_digest64"]):
        t = O.generate_topology(20, seed=int(s))
        h = hashlib.sha256()
        h.update(t["edges"].tobytes()), h.update(t["node_edges"].tobytes()), h.update(t["apsp"].tobytes())
        assert np.frombuffer(h.digest()[:8], dtype=np.uint64)[0] == d, s


def test_n200_topology():
    g = load_golden("topology")
    t = O.generate_topology(200, seed=476)
    assert np.array_equal(t["edges"], g["n200_edges"])
    assert np.array_equal(t["node_edges"], g["n200_node_edges"])
    assert np.array_equal(t["apsp"], g["n200_apsp"])
    with pytest.raises(AssertionError):
        O.generate_topology(50, seed=476)  # invalid for N=50


def test_random_topology_chain_and_stream_restore():
    g = load_golden("topology")
    rng = O.MT19937(7)
    ex = set(int(x) for x in g["eval_seeds"])
    for i in range(len(g["chain_seed"])):
        t = O.generate_topology(20, seed=None, global_rng=rng, exclude=ex)
        assert t["seed"] == g["chain_seed"][i]
        assert t["repetitions"] == g["chain_rep"][i]
        assert np.array_equal(t["edges"], g["chain_edges"][i])
        assert rng.pos == g["chain_pos"][i]
    assert rng.random() == g["chain_next_u"][0]
    assert (g["chain_rep"] > 1).any()  # the reseed path was exercised


def test_seed_pool():
    g = load_golden("topology")
    ex = set(int(x) for x in g["eval_seeds"])
    r = O.MT19937(476)  # build_seed_list seeds the global stream with topology_init_seed
    pool = []
    while len(pool) < len(g["pool_seeds"]):
        t = O.generate_topology(20, seed=None, global_rng=r, exclude=ex)
        if t["seed"] not in pool:
            pool.append(t["seed"])
    assert pool == g["pool_seeds"].tolist()
    glob = O.MT19937(99)
    picks = [glob.choice(pool) for _ in range(len(g["pool_picks"]))]
    assert picks == g["pool_picks"].tolist()
