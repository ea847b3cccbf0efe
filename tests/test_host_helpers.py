"""CPU checks of the small host helpers that main.py uses next to the rollout classes: `Buffer` (src/buffer.py:4-63
semantics: append one element or a list, grow on overflow, `_count`, `mean(default)`), `util` (src/util.py:8-111) and
the environment contract (src/env/environment.py:7-131)."""
import numpy as np
import pytest
import torch

from graph_marl_b200.buffer import Buffer
from graph_marl_b200.env.environment import EnvironmentVariant, NetworkEnv, reset_and_get_sizes
from graph_marl_b200 import util


def test_buffer_insert_grow_mean_clear():
    b = Buffer(2, (3,), np.float32)
    assert b.mean() == 0 and b.mean(default=-1) == -1 and b._count == 0
    b.insert(np.array([1, 2, 3]))
    b.insert([np.array([4, 5, 6]), np.array([7, 8, 9])])  # list insert crosses the initial capacity
    assert b._count == 3
    np.testing.assert_array_equal(b.get(), np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], np.float32))
    assert b.get().dtype == np.float32
    assert b.mean() == pytest.approx(5.0)
    b.insert([])
    assert b._count == 3
    for i in range(40):  # repeated growth keeps what was stored
        b.insert(np.full(3, 10 + i))
    assert b._count == 43 and b.get()[2, 2] == 9 and b.get()[-1, 0] == 49
    b.clear()
    assert b._count == 0 and b.get().shape == (0, 3) and b.mean(default=7) == 7
    b.insert(np.ones(3))
    assert b._count == 1 and b.mean() == 1


def test_buffer_scalar_shape_and_tensor_inputs():
    b = Buffer(4, (1,), np.float32)  # main.py:615 style: Buffer(size, (n_data,), np.float32) with reward.mean()
    b.insert(np.float32(0.5))
    b.insert(torch.tensor(1.5))       # device / torch scalars from the batched path
    b.insert([1.0, 2.0])
    np.testing.assert_allclose(b.get()[:, 0], [0.5, 1.5, 1.0, 2.0])
    assert b.mean() == pytest.approx(1.25)


def test_util_helpers():
    assert util.dim_str_to_list("") == [] and util.dim_str_to_list("512,256") == [512, 256]
    assert util.one_hot_list(2, 4) == [0, 0, 1, 0] and util.one_hot_list(-1, 3) == [0, 0, 0]
    assert util.filter_dict(dict(a=1, b=2, c=3), ["a", "c"]) == dict(a=1, c=3)

    class Obj:
        x = 1

    o = Obj()
    util.set_attributes(o, dict(x=2, y=3))
    assert (o.x, o.y) == (2, 3)
    util.set_seed(5)
    a = (np.random.rand(), torch.rand(1).item())
    util.set_seed(5)
    assert a == (np.random.rand(), torch.rand(1).item())


def test_set_attributes_verbose_reports_changes(capsys):
    class Obj:
        keep = 1
        change = 1

    util.set_attributes(Obj(), dict(keep=1, change=2, new=3), verbose=True)
    out = capsys.readouterr().out.splitlines()
    assert out == ["> Updated: change = 2", "> Added: new = 3"]


def test_checkpoint_round_trip_and_soft_update():
    a, b = torch.nn.Linear(3, 2), torch.nn.Linear(3, 2)
    nm = torch.nn.Linear(2, 2)
    ck = util.get_state_dict(a, nm, dict(lr=1))
    assert set(ck) == {"type", "state_dict", "args", "netmon_state_dict"} and ck["type"] == "Linear"
    util.load_state_dict(ck, b, torch.nn.Linear(2, 2))
    assert torch.equal(a.weight, b.weight)
    with pytest.raises(ValueError):
        util.load_state_dict(ck, b, None)  # checkpoint has a NetMon, caller has none
    with pytest.raises(ValueError):
        util.load_state_dict(util.get_state_dict(a, None, None), b, nm)
    src, tar = torch.nn.Linear(2, 1), torch.nn.Linear(2, 1)
    w_src, w_tar = src.weight.detach().clone(), tar.weight.detach().clone()
    util.interpolate_model(src, tar, 0.25, tar)  # main.py:1019: tar <- tau * model + (1 - tau) * tar
    assert torch.allclose(tar.weight, 0.25 * w_src + 0.75 * w_tar)


def test_environment_contract():
    assert [v.value for v in EnvironmentVariant] == [1, 2, 3]
    assert EnvironmentVariant(2) is EnvironmentVariant.WITH_K_NEIGHBORS
    with pytest.raises(TypeError):
        NetworkEnv()  # abstract

    class Toy(NetworkEnv):
        def reset(self):
            return np.zeros((4, 7)), np.eye(4)

        def step(self, act):
            return None

        def get_node_observation(self):
            return np.zeros((3, 5))

        def get_nodes_adjacency(self):
            return np.eye(3)

        def get_node_agent_matrix(self):
            return np.zeros((3, 4))

        def get_num_agents(self):
            return 4

        def get_num_nodes(self):
            return 3

    t = Toy()
    assert reset_and_get_sizes(t) == (4, 7, 3, 5)
    assert t.get() is t and t.get_node_aux() is None and t.get_final_info({"a": 1}) == {"a": 1}
