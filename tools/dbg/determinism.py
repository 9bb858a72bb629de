"""debug: bf16 full-size step: run-to-run determinism (E=1 twice) and batch independence (E=2 with a duplicated event)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_fullsize import draws_for, fresh_nets, gpu_step, rel
from iea_gan_b200.default_config import shipped_config
adt = sys.argv[1] if len(sys.argv) > 1 else "bf16"
cfg = shipped_config(H_base=1, device="cuda", clip_norm=1e9)
ph1 = draws_for(cfg, 401, 40, 256, 256)
torch.manual_seed(402)
x1 = torch.rand(40, 1, 256, 256) * 2 - 1
y1 = torch.arange(40)
def dup(t):
    return torch.cat([t, t], 0)
ph2 = [(dup(z), dup(rd), {k: dup(v) for k, v in d.items()}) for z, rd, d in ph1]
def run(phases, x, y):
    G, D, _, _ = fresh_nets(cfg)
    for o in (G.optim, D.optim):
        for grp in o.param_groups:
            grp["lr"] = 0.0
    G, D, got = gpu_step(cfg, phases, x, y, adt, nets=(G, D, None, None))
    return {k: p.grad.clone() for k, p in G.named_parameters()}, {k: p.grad.clone() for k, p in D.named_parameters()}, got
a = run(ph1, x1, y1)
b = run(ph1, x1, y1)
c = run(ph2, dup(x1), dup(y1))
for name, u, v in (("run-to-run E=1", a, b), ("E=2 duplicated vs E=1", c, a)):
    for i, tag in enumerate("GD"):
        rs = sorted(((rel(u[i][k], v[i][k]), k) for k in u[i] if float(v[i][k].norm()) > 1e-5), reverse=True)
        print(name, tag, "worst", [(float("%.3g" % r), k) for r, k in rs[:5]], "median %.3g" % rs[len(rs) // 2][0])
    print(name, u[2], v[2])
