"""The learner's collectives on NCCL hardware: run under torchrun on N GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/allreduce_check.py

Every rank builds the paper-size NetMon + DQN with its own seed, broadcasts rank 0's weights, runs one
device-side forward + backward (csrc/train.cu) on its own data, then all-reduces the flat gradient through the
library's C-ABI collective (gm_allreduce_grads -> ncclAllReduce) and through torch.distributed for comparison.
Prints one JSON line: element count, equality with the sum of the per-rank gradients, device time of both forms.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F

from graph_marl_b200 import _lib
from graph_marl_b200.learner_sync import allreduce_gradients, broadcast_weights, destroy_nccl_comms, nccl_comm
from graph_marl_b200.model import DQN, NetMon


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    N = A = 20
    H = 128
    torch.manual_seed(100 + rank)
    nm = NetMon(4 * N + 8, H, (512, 256), 3, F.leaky_relu, rnn_type="lstm", output_neighbor_hidden=True).cuda()
    dq = DQN(6 * N + 10 + 4 * H, (512, 256), 4, F.leaky_relu).cuda()
    broadcast_weights([nm, dq], src=0)
    w = torch.cat([p.detach().reshape(-1) for m in (nm, dq) for p in m.parameters()])
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    same_weights = all(torch.equal(ws[0], x) for x in ws)
    torch.manual_seed(7 + rank)
    B = 32
    x = (torch.rand(B, N, 4 * N + 8, device="cuda") < 0.05).float()
    adj = (torch.eye(N, device="cuda") + torch.roll(torch.eye(N, device="cuda"), 1, 1) + torch.roll(torch.eye(N, device="cuda"), -1, 1)).clamp(max=1).repeat(B, 1, 1)
    eye = torch.eye(N, device="cuda").repeat(B, 1, 1)
    g_obs = nm(x, adj, eye)
    obs = torch.cat((torch.rand(B, A, 6 * N + 10, device="cuda"), g_obs), -1)
    dq(obs, None).square().mean().backward()
    params = [p for m in (nm, dq) for p in m.parameters()]
    local_flat = torch.cat([p.grad.reshape(-1) for p in params]).clone()
    gathered = [torch.empty_like(local_flat) for _ in range(world)]
    dist.all_gather(gathered, local_flat)
    n = allreduce_gradients(params, average=False)
    reduced = torch.cat([p.grad.reshape(-1) for p in params])
    expect = torch.stack(gathered).sum(0)
    err = float((reduced - expect).abs().max() / expect.abs().max())
    # device time of the fused all-reduce: this library's C-ABI call vs torch.distributed on the same buffer
    flat = local_flat.clone()
    comm = nccl_comm()

    def timed(fn, iters=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    us_gm = timed(lambda: _lib.check(_lib.lib().gm_allreduce_grads(comm, flat.data_ptr(), flat.numel(), 1, _lib.current_stream())))
    us_torch = timed(lambda: dist.all_reduce(flat))
    if rank == 0:
        print(json.dumps(dict(world=world, elements=n, bytes=4 * n, same_weights_after_broadcast=same_weights,
                              allreduce_rel_err_vs_sum_of_rank_gradients=err, nccl_version=_lib.lib().gm_nccl_version(),
                              gm_allreduce_grads_us=us_gm, torch_all_reduce_us=us_torch,
                              note="device time per call, max over ranks, 50 calls back to back")))
    destroy_nccl_comms()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
