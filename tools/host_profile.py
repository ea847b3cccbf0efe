"""cProfile of the host side of Rollout.step (python + ctypes + torch allocator), GPU kept async."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graph_marl_b200.rollout import Rollout

ro = Rollout("cfg2", num_envs=4096, math="bf16x3", seed=1)
ro.reset()
for _ in range(5):
    ro.step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    ro.step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumtime").print_stats(28)
