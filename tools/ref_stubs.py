"""Import shim that lets the UNMODIFIED reference (/root/reference/src) import in
this container.  Only used by tools/gen_golden.py (fixture generation); nothing
in the product, the tests, smoke() or bench.py imports this file.

The reference needs three packages that are not installed here:
gymnasium (only ``Discrete(n, start).n``), matplotlib (plotting only) and
torch_geometric (non-default aggregations + cosmetic summary()).
"""
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def install():
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")
        spaces = types.ModuleType("gymnasium.spaces")

        class Discrete:
            def __init__(self, n, start=0):
                self.n = int(n)
                self.start = int(start)

            def sample(self):
                import numpy as np

                return self.start + int(np.random.randint(self.n))

        spaces.Discrete = Discrete
        gym.spaces = spaces
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt

    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tg_nn = types.ModuleType("torch_geometric.nn")
        tg_utils = types.ModuleType("torch_geometric.utils")
        tg_sum = types.ModuleType("torch_geometric.nn.summary")
        for name in ("GCNConv", "SAGEConv", "AntiSymmetricConv", "GraphSAGE"):
            setattr(tg_nn, name, type(name, (), {}))
        tg_utils.dense_to_sparse = lambda *a, **k: (_ for _ in ()).throw(
            NotImplementedError("torch_geometric stub")
        )
        tg_sum.summary = lambda *a, **k: "(summary unavailable: torch_geometric stub)"
        tg.nn = tg_nn
        tg.utils = tg_utils
        tg_nn.summary = tg_sum
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.nn"] = tg_nn
        sys.modules["torch_geometric.utils"] = tg_utils
        sys.modules["torch_geometric.nn.summary"] = tg_sum

    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
