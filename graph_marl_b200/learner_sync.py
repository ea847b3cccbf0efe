"""The only collectives of a multi-GPU run: the rollout path has none -- env instances are independent
and every rank holds a full replica of NetMon + DQN -- but when each rank trains on its own replay shard the
learner (src/main.py:1002-1004, between `loss.backward()` and `optimizer.step()`) needs the gradients summed and,
once at start-up, identical weights.

Both move ONE flat buffer per call (~0.94 M fp32 elements for the paper's NetMon + DQN): collective cost on NVSwitch
is launch-latency bound, so one fused all-reduce beats one per parameter tensor.  On CUDA tensors the call is the
library's own C-ABI collective (`gm_allreduce_grads` / `gm_broadcast_weights`, csrc/collective.cpp: ncclAllReduce on a
communicator the library creates from a unique id that rank 0 hands out through the default process group); on CPU
tensors (the gloo tests) it is `torch.distributed`.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

_COMM = {}


def nccl_comm(group=None):
    """This library's NCCL communicator over the ranks of `group` (created once; the 128-byte unique id travels
    through torch.distributed's object broadcast, any backend)."""
    key = id(group)
    if key not in _COMM:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        ident = np.zeros(_lib.GM_NCCL_UNIQUE_ID_BYTES, np.uint8)
        if rank == 0:
            _lib.check(_lib.lib().gm_nccl_unique_id(_lib.ptr(ident)))
        box = [ident.tobytes()]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = np.frombuffer(box[0], np.uint8).copy()
        comm = C.c_void_p()
        _lib.check(_lib.lib().gm_nccl_comm_create(world, rank, _lib.ptr(ident), C.byref(comm)))
        _COMM[key] = comm
    return _COMM[key]


def destroy_nccl_comms():
    for comm in _COMM.values():
        _lib.lib().gm_nccl_comm_destroy(comm)
    _COMM.clear()


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
            if flat.is_cuda and dtype == torch.float32:
                with torch.cuda.device(flat.device):
                    _lib.check(_lib.lib().gm_broadcast_weights(nccl_comm(group), flat.data_ptr(), flat.numel(), src,
                                                               _lib.current_stream()))
            else:
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
        if world > 1 and flat.is_cuda:
            with torch.cuda.device(flat.device):
                _lib.check(_lib.lib().gm_allreduce_grads(nccl_comm(group), flat.data_ptr(), flat.numel(), int(average),
                                                         _lib.current_stream()))
        elif world > 1:
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
