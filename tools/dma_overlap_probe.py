"""Do pitched device-to-device cudaMemcpy2DAsync copies (the big replay fields) overlap the tensor-core kernels?
Times NetMon forwards on one stream, the copies on another, and both together."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from cuda.bindings import runtime as rt

from graph_marl_b200.rollout import Rollout

ro = Rollout("cfg2", num_envs=4096, math="bf16x3", with_replay=False)
ro.reset()
for _ in range(3):
    ro.step()
torch.cuda.synchronize()
R = 4096 * 20
src_g = torch.randn(R, 512, device="cuda")
src_a = torch.randn(R, 130, device="cuda")
dst = torch.empty(2 * R, 642, device="cuda")
sB = torch.cuda.Stream()


def copies(stream, reps):
    h = stream.cuda_stream
    for i in range(reps):
        for half in range(2):  # obs and next_obs
            base = dst.data_ptr() + half * R * 642 * 4
            (e,) = rt.cudaMemcpy2DAsync(base, 642 * 4, src_a.data_ptr(), 130 * 4, 130 * 4, R, rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice, h)
            assert e == rt.cudaError_t.cudaSuccess, e
            (e,) = rt.cudaMemcpy2DAsync(base + 130 * 4, 642 * 4, src_g.data_ptr(), 512 * 4, 512 * 4, R, rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice, h)
            assert e == rt.cudaError_t.cudaSuccess, e


def gemms(reps):
    for _ in range(reps):
        ro.env._netmon_step()


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    sB.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


reps = 20
gemms(2); copies(sB, 1)
tA = timed(lambda: gemms(reps))
tB = timed(lambda: copies(sB, reps))
def both():
    sB.wait_stream(torch.cuda.current_stream())
    copies(sB, reps)
    gemms(reps)
    torch.cuda.current_stream().wait_stream(sB)
tAB = timed(both)
gb = 2 * R * 642 * 4 * 2 / 1e9
print(f"netmon only {tA/reps:.3f} ms/iter; copies only {tB/reps:.3f} ms/iter ({gb/(tB/reps)*1e3:.0f} GB/s r+w); together {tAB/reps:.3f} ms/iter")
