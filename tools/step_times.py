"""Per-step device time of consecutive G+D train steps from a cold start (warm-up behaviour of the caching allocator)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iea_gan_b200 as P
from iea_gan_b200.default_config import shipped_config
from iea_gan_b200.train_step import make_train_step, NormalNoise, EMA
cfg = shipped_config(H_base=1, device="cuda", clip_norm=1e9)
torch.manual_seed(0)
G, D = P.Generator(**cfg).cuda(), P.Discriminator(**cfg).cuda()
G.train(); D.train()
ev = 8; n = 40 * ev
G_ema = P.Generator(**dict(cfg, skip_init=True, no_optim=True)).cuda()
ema = EMA(G, G_ema, cfg["ema_decay"], cfg["ema_start"])
train = make_train_step(G, D, P.G_D(G, D), NormalNoise(n, cfg["dim_z"], "cuda"), dict(cfg, batch_size=n), ema=ema)
x = torch.rand(n, 1, 256, 256, device="cuda") * 2 - 1
y = torch.arange(40, device="cuda").repeat(ev)
ts = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 36):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); train(x, y); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(" ".join("%.0f" % t for t in ts))
print("reserved GB %.1f allocated GB %.1f, cudaMalloc retries %d" % (torch.cuda.memory_reserved() / 1e9, torch.cuda.max_memory_allocated() / 1e9, torch.cuda.memory_stats().get("num_alloc_retries", 0)))
