"""The only collectives of a multi-GPU run: the rollout path has none -- env instances are independent
and every rank holds a full replica of NetMon + DQN -- but when each rank trains on its own replay shard the
learner (src/main.py:1002-1004, between `loss.backward()` and `optimizer.step()`) needs the gradients summed and,
once at start-up, identical weights.

Both go through `torch.distributed` (NCCL over NVLink on the GPUs, gloo in the CPU tests) as ONE flat buffer per
call (~0.94 M fp32 elements for the paper's NetMon + DQN): collective cost on NVSwitch is launch-latency bound,
so one fused all-reduce beats one per parameter tensor.
"""
import torch
import torch.distributed as dist


def _flat_views(tensors):
    flat = torch.cat([t.reshape(-1) for t in tensors])
    return flat


def _scatter_back(flat, tensors):
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def broadcast_weights(modules, src=0, group=None):
    """Make every rank's parameters and buffers equal to rank `src`'s (one flat broadcast per dtype)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    tensors = [t for m in modules for t in list(m.parameters()) + list(m.buffers())]
    with torch.no_grad():
        for dtype in sorted({t.dtype for t in tensors}, key=str):
            part = [t.data for t in tensors if t.dtype == dtype]
            flat = _flat_views(part)
            dist.broadcast(flat, src=src, group=group)
            _scatter_back(flat, part)


def allreduce_gradients(parameters, average=True, group=None):
    """Sum (or average) the gradients of `parameters` over all ranks with one fused all-reduce.  Parameters without
    a gradient on this rank contribute zeros (every rank must pass the same parameter list).  Returns the number of
    elements reduced."""
    params = [p for p in parameters if p.requires_grad]
    if not params:
        return 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    with torch.no_grad():
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
        flat = _flat_views([g.to(torch.float32) for g in grads])
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                flat /= world
        off = 0
        for p, g in zip(params, grads):
            n = p.numel()
            new = flat[off:off + n].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = new.clone()
            else:
                p.grad.copy_(new)
            off += n
    return int(flat.numel())
