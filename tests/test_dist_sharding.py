"""Multi-process (gloo, world_size 2, CPU) checks of the N>1 path of bench.py: env-instance shards
are disjoint and cover the job, per-rank draw streams are distinct, and the whole-job throughput is
(sum of units) / (max time over ranks)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graph_marl_b200.rollout import aggregate_throughput, shard_envs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_envs(total, world, rank)
    # a rank is slower the higher its index: the aggregate must use the slowest one
    ms_local = 10.0 * (rank + 1)
    value, ms_max = aggregate_throughput((hi - lo) * 5, ms_local, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi))
    if rank == 0:
        out.put((value, ms_max, gathered))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [4096, 4097])
def test_two_rank_sharding_and_max_over_ranks(total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    value, ms_max, shards = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert shards[0][0] == 0 and shards[-1][1] == total
    assert all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))  # disjoint, contiguous cover
    assert max(h - l for l, h in shards) - min(h - l for l, h in shards) <= 1
    assert ms_max == 20.0
    assert value == pytest.approx(total * 5 / 20e-3)


def test_shard_envs_properties():
    for total in (1, 7, 4096, 16384):
        for world in (1, 2, 4, 8):
            spans = [shard_envs(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert sum(h - l for l, h in spans) == total


def _learner_worker(rank, world, port, out):
    import torch.nn.functional as F

    from graph_marl_b200.learner_sync import allreduce_gradients, broadcast_weights
    from graph_marl_b200.model import DQN

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)  # different initial weights per rank
    model = DQN(12, (16, 8), 4, F.leaky_relu)
    broadcast_weights([model], src=0)
    w_after = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    torch.manual_seed(7 + rank)  # different data per rank (its own replay shard)
    x = torch.randn(5, 3, 12)
    loss = model(x, None).pow(2).mean()
    loss.backward()
    local = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    n = allreduce_gradients(model.parameters(), average=True)
    reduced = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    gathered = [None] * world
    dist.all_gather_object(gathered, (w_after, local, reduced, n))
    if rank == 0:
        # by value (numpy): a torch tensor travels as a shared-memory handle that the parent must fetch from THIS
        # process, which may have exited by then on a loaded machine
        out.put([(w.numpy(), g.numpy(), r.numpy(), k) for w, g, r, k in gathered])
    dist.barrier()
    dist.destroy_process_group()


def test_learner_weight_broadcast_and_fused_gradient_allreduce():
    """the only collectives of a multi-GPU job are the start-up weight broadcast and the learner's fused
    gradient all-reduce (gloo here, NCCL on the GPUs)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_learner_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (w0, g0, r0, n0), (w1, g1, r1, n1) = [(torch.from_numpy(w), torch.from_numpy(g), torch.from_numpy(r), k) for w, g, r, k in res]
    assert torch.equal(w0, w1)                       # identical replicas after the broadcast
    assert not torch.equal(g0, g1)                   # ranks saw different data
    assert torch.allclose(r0, (g0 + g1) / 2, atol=1e-7) and torch.equal(r0, r1)
    assert n0 == n1 == w0.numel()
