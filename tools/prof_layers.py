"""Per-call profile of the hot path: every C-ABI call of one G+D train step (or one Generator
sampling pass) timed with CUDA events on the launching stream, grouped by (entry point, shape).
For convolutions the algorithmic bytes (input + output, layer granular) give achieved GB/s.
usage: python tools/prof_layers.py train|sample [events] [top]"""
import collections
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iea_gan_b200 as P
from iea_gan_b200 import engine as E_
from iea_gan_b200.default_config import shipped_config
from iea_gan_b200.train_step import make_train_step, NormalNoise

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
ev = int(sys.argv[2]) if len(sys.argv) > 2 else (8 if mode == "train" else 16)
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
cfg = shipped_config(H_base=1, device="cuda", clip_norm=1e9)
torch.manual_seed(0)
G = P.Generator(**cfg).cuda()
G.train()
n = 40 * ev
y = torch.arange(40, device="cuda").repeat(ev)
if mode == "train":
    D = P.Discriminator(**cfg).cuda()
    D.train()
    train = make_train_step(G, D, P.G_D(G, D), NormalNoise(n, cfg["dim_z"], "cuda"), dict(cfg, batch_size=n))
    x = torch.rand(n, 1, 256, 256, device="cuda") * 2 - 1
    step = lambda: train(x, y)
else:
    z = torch.randn(n, cfg["dim_z"], device="cuda")

    def step():
        with torch.no_grad():
            G(z, y)
for _ in range(2):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); step(); b.record()
torch.cuda.synchronize()
plain_ms = a.elapsed_time(b)
E_.PROFILE = []
a.record(); step(); b.record()
torch.cuda.synchronize()
prof_ms = a.elapsed_time(b)
t, c = collections.Counter(), collections.Counter()
for name, tag, s, e in E_.PROFILE:
    t[(name, tag)] += s.elapsed_time(e)
    c[(name, tag)] += 1
tot = sum(t.values())
print("# %s, %d events: step %.2f ms plain, %.2f ms with per-call events; C-ABI calls %d, sum of call times %.2f ms"
      % (mode, ev, plain_ms, prof_ms, len(E_.PROFILE), tot))
by = collections.Counter()
for (name, tag), v in t.items():
    by[name] += v
print("# by entry point: " + ", ".join("%s %.1f" % (k.replace("iea_", ""), v) for k, v in by.most_common(14)))
print("%9s %5s %5s %8s  %s" % ("ms", "%", "calls", "GB/s", "call"))
for (name, tag), v in t.most_common(top):
    gbs = ""
    if name == "iea_conv_fprop":
        f = tag.split()
        nn, (h, w), (ci, co) = int(f[0][1:]), map(int, f[1].split("x")), map(int, f[2].split("->"))
        mode_in = int(f[4][2:])
        px_in = nn * h * w * (0.25 if mode_in == 1 else 4 if mode_in == 2 else 1)
        gbs = "%.0f" % ((px_in * ci + nn * h * w * co) * 2 * c[(name, tag)] / (v * 1e-3) / 1e9)
    print("%9.3f %5.1f %5d %8s  %s %s" % (v, 100 * v / tot, c[(name, tag)], gbs, name.replace("iea_", ""), tag))
