import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


ROUTING_CASES = [
    "routing_A_seed923430603_cong", "routing_B_nocong", "routing_C_mask", "routing_D_ttl",
    "routing_E_a35_nocong_mask_ttl", "routing_F_n200", "routing_G_var2", "routing_H_var3",
    "routing_I_evalinfo",
]
