"""Data-parallel host logic on CPU: world_size 2, gloo backend.  The
kernels are not involved: this covers event sharding, the state broadcast, the flat gradient buffer
(optim.FlatGrads: `.grad` are views of one buffer), the end-of-backward all-reduce hook (dp.GradSync, the
mechanism that makes an unchanged train_fns.py data-parallel) and the buffer sync before checkpoints,
i.e. that N ranks x E/N events reproduce the 1-rank gradient."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Hooked(torch.autograd.Function):
    """Stand-in for the engine's autograd bridge: its backward writes nothing itself but reports the end of the
    net's backward to the attached GradSync, exactly where engine._NetFn.backward does."""

    @staticmethod
    def forward(ctx, x, net):
        ctx.net = net
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        sync = ctx.net.__dict__.get("_iea_grad_sync")
        if sync is not None:
            sync.after_backward(ctx.net)
        return g, None


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from iea_gan_b200 import dp
    from iea_gan_b200.optim import FlatGrads
    torch.manual_seed(123 + rank)  # different init per rank: broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
    net.register_buffer("u0", torch.randn(1, 16))
    net.register_buffer("stored_mean", torch.full((4,), float(rank)))
    dp.broadcast_state(net)
    sync = dp.attach(net)
    fl = FlatGrads.of(net)
    for p in net.parameters():  # what engine.Tape.galloc does on the first backward after zero_grad()
        p.grad = fl.views[id(p)]
    torch.manual_seed(7)
    x = torch.randn(6 * 40, 8)  # 6 events
    b, e = dp.shard_events(6)
    xin = x[b * 40:e * 40].clone().requires_grad_(True)
    # two uses of the net in one pass (like D(fake) + D(real)): the all-reduce must still fire exactly once
    loss = net(_Hooked.apply(xin, net)).pow(2).mean() * 0.5 + net(_Hooked.apply(xin, net)).pow(2).mean() * 0.5
    loss.backward()  # the hook fires at the end of this call: no explicit all-reduce below
    assert sync.count == 1 and not sync.pending
    assert all(p.grad.data_ptr() == fl.ptrs[id(p)] for p in net.parameters())
    net.stored_mean.fill_(float(rank))
    dp.sync_buffers(net)
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()] + [net.u0.reshape(-1), net.stored_mean])
    if rank == 0:
        torch.save({"flat": flat, "state": {k: v.clone() for k, v in net.state_dict().items()}, "x": x}, out)
    else:
        torch.save(flat, out + ".r1")
    dist.destroy_process_group()


def test_two_rank_hooked_allreduce_equals_single_rank(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0 = torch.load(out)
    r1 = torch.load(out + ".r1")
    assert torch.allclose(r0["flat"], r1), "ranks disagree after the all-reduce"
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
    net.register_buffer("u0", torch.zeros(1, 16))
    net.register_buffer("stored_mean", torch.zeros(4))
    net.load_state_dict(r0["state"])
    net(r0["x"]).pow(2).mean().backward()  # all 6 events on one rank
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()] + [net.u0.reshape(-1), torch.full((4,), 0.5)])
    assert torch.allclose(r0["flat"], ref, atol=1e-6)


def test_shard_events_partition():
    from iea_gan_b200 import dp
    for n, world in ((8, 1), (8, 2), (64, 4), (64, 8), (16, 8)):
        spans = [dp.shard_events(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert len({e - b for b, e in spans}) == 1
    with pytest.raises(ValueError):  # unequal shards would make the mean over ranks a weighted mean
        dp.shard_events(5, 0, 2)


def test_flat_grads_layout():
    from iea_gan_b200.optim import FlatGrads
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    fl = FlatGrads.of(net)
    assert fl is FlatGrads.of(net)
    for p in net.parameters():
        v = fl.views[id(p)]
        assert v.shape == p.shape and v.data_ptr() % 256 == fl.buf.data_ptr() % 256
    fl.buf.fill_(1.0)
    assert all(float(fl.views[id(p)].sum()) == p.numel() for p in net.parameters())
