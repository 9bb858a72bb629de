"""N-rank == 1-rank on the REAL Generator / Discriminator (run under torchrun, one rank per GPU):
every rank runs make_train_step on its shard of the events with dp.attach()'ed nets (the end-of-backward
all-reduce hook -- nothing in the step function knows about ranks); rank 0 then repeats the step on ALL events
in one process and compares losses-independent quantities: every parameter gradient and the parameters after
both optimizer steps.  fp32 activations, small configuration (tests/golden/small_cfg.json).
Prints DP_CHECK_OK <worst rel-L2> on success."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["IEA_ACT_DTYPE"] = "fp32"


def main():
    from test_gpu_fullsize import draws_for, replay_list, FixedZ, rel
    import iea_gan_b200 as P
    from iea_gan_b200 import dp, noise
    from iea_gan_b200.train_step import make_train_step
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "small_cfg.json")))
    cfg["device"] = "cuda"
    events = 2 * world
    rows = 40 * events
    phases = draws_for(cfg, 901, rows, 64, 64)
    torch.manual_seed(902)
    x = torch.rand(rows, 1, 64, 64) * 2 - 1
    y = torch.arange(40).repeat(events)

    def step(sl, attach):
        torch.manual_seed(0)
        G, D = P.Generator(**cfg).cuda().train(), P.Discriminator(**cfg).cuda().train()
        syncs = []
        if attach:
            dp.broadcast_state(G); dp.broadcast_state(D)
            syncs = [dp.attach(G), dp.attach(D)]
        n = sl.stop - sl.start
        train = make_train_step(G, D, P.G_D(G, D), FixedZ(phases, sl), dict(cfg, batch_size=n))
        with noise.replay(replay_list(phases, sl)):
            train(x[sl].cuda(), y[sl].cuda())
        torch.cuda.synchronize()
        return G, D, syncs
    b, e = dp.shard_events(events)
    G, D, syncs = step(slice(40 * b, 40 * e), True)
    assert [s.count for s in syncs] == [1, 1], [s.count for s in syncs]  # one all-reduce per net per step
    ok = torch.ones(1, device="cuda")
    if rank == 0:
        G1, D1, _ = step(slice(0, rows), False)
        worst = 0.0
        for a, r in ((G, G1), (D, D1)):
            for (k, p), (_, q) in zip(a.named_parameters(), r.named_parameters()):
                if float(q.grad.norm()) > 1e-6:
                    worst = max(worst, rel(p.grad, q.grad))
                worst = max(worst, rel(p, q) * 10)
        if worst < 2e-3:
            print("DP_CHECK_OK %.3g (world %d)" % (worst, world), flush=True)
        else:
            print("DP_CHECK_FAIL %.3g" % worst, flush=True)
            ok.zero_()
    dist.broadcast(ok, 0)
    dist.destroy_process_group()
    sys.exit(0 if float(ok) else 1)


if __name__ == "__main__":
    main()
