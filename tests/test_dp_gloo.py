"""Data-parallel host logic on CPU: world_size 2, gloo backend.  The
kernels are not involved: this covers event sharding, the state broadcast and the flat-bucket
gradient all-reduce (mean), i.e. that N ranks x E/N events reproduce the 1-rank gradient."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from iea_gan_b200 import dp
    torch.manual_seed(123 + rank)  # different init per rank: broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
    net.register_buffer("u0", torch.randn(1, 16))
    dp.broadcast_state(net)
    torch.manual_seed(7)
    x = torch.randn(6 * 40, 8)  # 6 events
    b, e = dp.shard_events(6)
    loss = net(x[b * 40:e * 40]).pow(2).mean()
    loss.backward()
    dp.allreduce_grads(net)
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()] + [net.u0.reshape(-1)])
    if rank == 0:
        torch.save({"flat": flat, "state": {k: v.clone() for k, v in net.state_dict().items()}, "x": x}, out)
    else:
        torch.save(flat, out + ".r1")
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_rank(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0 = torch.load(out)
    r1 = torch.load(out + ".r1")
    assert torch.allclose(r0["flat"], r1), "ranks disagree after the all-reduce"
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
    net.register_buffer("u0", torch.zeros(1, 16))
    net.load_state_dict(r0["state"])
    net(r0["x"]).pow(2).mean().backward()  # all 6 events on one rank
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()] + [net.u0.reshape(-1)])
    assert torch.allclose(r0["flat"], ref, atol=1e-6)


def test_shard_events_partition():
    from iea_gan_b200 import dp
    for n in (1, 5, 8, 64):
        for world in (1, 2, 4, 8):
            spans = [dp.shard_events(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
