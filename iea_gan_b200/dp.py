"""Data-parallel plumbing: one process per GPU, events sharded across ranks, NCCL all-reduce (mean)
of the D then the G gradient bucket over NVLink 5 / NVSwitch.

Nothing couples two events in the forward or backward pass (batch-norm statistics, the RRM, the
self-attention and every loss are per event), so the only exchange step of the path is this
gradient all-reduce; spectral-norm vectors stay identical on all ranks without communication
because the power-iteration kernel is deterministic.  The reference has no counterpart (its
`parallel` flag is never read, train.py:579-583).

How it reaches an UNCHANGED train_fns.py: `attach(net)` marks the net; the engine then writes every
parameter gradient of that net straight into one flat fp32 buffer (optim.FlatGrads; `p.grad` are views)
and, when the backward pass of the net has finished, calls `GradSync.after_backward`, which queues ONE
in-place all-reduce of that buffer on autograd's end-of-backward callback.  By the time
`D_loss.backward()` / `G_loss.backward()` returns, `.grad` holds the mean over ranks -- before
utils.ortho, clip_grad_norm_ and optim.step() read it (train_fns.py:130-139, 181-192).  No pack, no
unpack, no divide pass, no change to the caller.
"""
import torch
import torch.distributed as dist


def shard_events(n_events, rank=None, world=None):
    """[begin, end) of the events this rank owns.  Shards must be equal: the all-reduce averages per-rank
    means, which is the global per-event mean only then (pad or drop the remainder upstream)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if n_events % world:
        raise ValueError("%d events do not split evenly over %d ranks: the gradient mean over ranks would "
                         "weight events unequally" % (n_events, world))
    per = n_events // world
    return rank * per, (rank + 1) * per


def _float_state(net):
    return [t for t in net.state_dict().values() if t.is_floating_point()]


def broadcast_state(net, src=0, group=None):
    """Rank `src`'s parameters and buffers (incl. u0 / sv0 / running statistics) to every rank."""
    for t in _float_state(net):
        dist.broadcast(t, src, group=group)


def sync_buffers(net, group=None):
    """Before a checkpoint / evaluation: batch-norm running statistics are per-rank (each rank folds its own
    events) -> average them; u0 / sv0 are identical on all ranks by construction -> rank 0's copy wins."""
    world = dist.get_world_size(group)
    for k, t in net.state_dict().items():
        if not t.is_floating_point() or k in dict(net.named_parameters()):
            continue
        if k.endswith(("stored_mean", "stored_var")):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            t.div_(world)
        else:
            dist.broadcast(t, 0, group=group)


def allreduce_mean_(buf, group=None):
    """In-place mean over ranks of one flat buffer (NCCL: a single AVG all-reduce; gloo has no AVG)."""
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        buf.div_(dist.get_world_size(group))
    return buf


class GradSync:
    """End-of-backward all-reduce of a net's flat gradient buffer."""

    def __init__(self, group=None):
        self.group = group
        self.pending = False
        self.enabled = True  # micro-batched steps switch it off for all but the last backward (cf. DDP.no_sync)
        self.count = 0

    def after_backward(self, net):
        """Called by the engine at the end of every backward of `net` inside one autograd pass (the
        Discriminator runs twice per D step): queue the all-reduce once, at the end of the pass."""
        if self.pending or not self.enabled:
            return
        self.pending = True

        def fire():
            self.pending = False
            self.count += 1
            allreduce_mean_(net.__dict__["_iea_flat"].buf, self.group)
        torch.autograd.Variable._execution_engine.queue_callback(fire)


def attach(net, group=None):
    """Make `net`'s gradients data-parallel: every backward ends with the mean over ranks in `.grad`."""
    sync = GradSync(group)
    net.__dict__["_iea_grad_sync"] = sync
    return sync


def detach(net):
    net.__dict__.pop("_iea_grad_sync", None)


def allreduce_grads(net, group=None):
    """Explicit form for callers that do not use attach(): mean of .grad over ranks.  Uses the flat buffer when
    the net has one, otherwise a fixed-order bucket over ALL parameters (missing gradients count as zero, so
    every rank contributes the same layout)."""
    fl = net.__dict__.get("_iea_flat")
    ps = list(net.parameters())
    if fl is not None and all(p.grad is None or p.grad.data_ptr() == fl.ptrs[id(p)] for p in ps):
        allreduce_mean_(fl.buf, group)
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
    allreduce_mean_(flat, group)
    off = 0
    for p in ps:
        n = p.numel()
        if p.grad is not None:
            p.grad.copy_(flat[off:off + n].view_as(p))
        off += n
