"""graph_marl_b200 -- B200-native (sm_100a) rollout hot path of jw3il/graph-marl.

Thin PyTorch/ctypes host classes with the reference's names and signatures over
libgraphmarl_b200.so (include/graphmarl_b200.h).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import GraphMarlError  # noqa: F401

__version__ = "0.1.0"
