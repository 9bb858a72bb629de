"""Data-parallel plumbing: one process per GPU, events sharded across ranks, NCCL all-reduce (mean)
of the D then the G gradient bucket over NVLink 5 / NVSwitch.

Nothing couples two events in the forward or backward pass (batch-norm statistics, the RRM, the
self-attention and every loss are per event), so the only exchange step of the path is this
gradient all-reduce; spectral-norm vectors stay identical on all ranks without communication
because the power-iteration kernel is deterministic.  The reference has no counterpart (its
`parallel` flag is never read, train.py:579-583).
"""
import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


def shard_events(n_events, rank=None, world=None):
    """[begin, end) of the events this rank owns (contiguous, balanced)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    base, rem = divmod(n_events, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def broadcast_state(net, src=0):
    """Rank `src`'s parameters and buffers (incl. u0 / sv0 / running statistics) to every rank."""
    ts = [t for t in net.state_dict().values() if t.is_floating_point()]
    flat = _flatten_dense_tensors(ts)
    dist.broadcast(flat, src)
    for t, f in zip(ts, _unflatten_dense_tensors(flat, ts)):
        t.copy_(f)


def allreduce_grads(net):
    """Mean of .grad over ranks, one flat bucket per net (17.9 MB for D, 46.8 MB for G in fp32)."""
    ps = [p for p in net.parameters() if p.grad is not None]
    if not ps:
        return
    grads = [p.grad for p in ps]
    flat = _flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(dist.get_world_size())
    for g, f in zip(grads, _unflatten_dense_tensors(flat, grads)):
        g.copy_(f)
