"""How close is the train step to the host launch-rate limit?  Times the step at 1 event (GPU work ~1/8:
the step time is then mostly host time) and at 8 events.  usage: python tools/cpu_bound.py"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iea_gan_b200 as P
from iea_gan_b200.default_config import shipped_config
from iea_gan_b200.train_step import make_train_step, NormalNoise
cfg = shipped_config(H_base=1, device="cuda", clip_norm=1e9)
torch.manual_seed(0)
G, D = P.Generator(**cfg).cuda(), P.Discriminator(**cfg).cuda()
G.train(); D.train()
for ev in (1, 8):
    n = 40 * ev
    train = make_train_step(G, D, P.G_D(G, D), NormalNoise(n, cfg["dim_z"], "cuda"), dict(cfg, batch_size=n))
    x = torch.rand(n, 1, 256, 256, device="cuda") * 2 - 1
    y = torch.arange(40, device="cuda").repeat(ev)
    for _ in range(6):
        train(x, y)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        train(x, y)
    t_launch = (time.perf_counter() - t0) / 5   # includes the .tolist() sync at the end of each step
    torch.cuda.synchronize()
    print("events %d: %.1f ms per step (wall, host + device)" % (ev, t_launch * 1e3))
