"""debug: full-size step, E events: bf16-activation gradients vs fp32-activation gradients (both on the GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_fullsize import draws_for, fresh_nets, gpu_step, rel
from iea_gan_b200.default_config import shipped_config
E_ = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = shipped_config(H_base=1, device="cuda", clip_norm=1e9)
rows = 40 * E_
phases = draws_for(cfg, 401, rows, 256, 256)
torch.manual_seed(402)
x = torch.rand(rows, 1, 256, 256) * 2 - 1
y = torch.arange(40).repeat(E_)
res = {}
for adt in ("fp32", "bf16"):
    G, D, _, _ = fresh_nets(cfg)
    for grp in D.optim.param_groups:
        grp["lr"] = 0.0
    G, D, got = gpu_step(cfg, phases, x, y, adt, nets=(G, D, None, None))
    res[adt] = ({k: p.grad.clone() for k, p in G.named_parameters()}, {k: p.grad.clone() for k, p in D.named_parameters()}, got)
    print(adt, got)
for i, tag in enumerate("GD"):
    a, b = res["fp32"][i], res["bf16"][i]
    rs = sorted(((rel(b[k], a[k]), k) for k in a if float(a[k].norm()) > 1e-5), reverse=True)
    print(tag, "E=%d" % E_, "worst", [(round(r, 3), k) for r, k in rs[:8]], "median %.3g" % rs[len(rs) // 2][0])
