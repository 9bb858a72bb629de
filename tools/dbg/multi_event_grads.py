"""debug: small config, E events, fp32 activations: per-parameter gradient error vs the CPU oracle."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
os.environ["IEA_ACT_DTYPE"] = sys.argv[2] if len(sys.argv) > 2 else "fp32"
from test_gpu_fullsize import draws_for, replay_list, FixedZ, rel
import iea_gan_b200 as P
from iea_gan_b200 import noise
from iea_gan_b200.train_step import make_train_step
from oracle import iea_oracle as O
E_ = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = json.load(open(os.path.join(ROOT, "tests/golden/small_cfg.json")))
cfg["device"] = "cuda"
rows = 40 * E_
phases = draws_for(cfg, 401, rows, 64, 64)
torch.manual_seed(402)
x = torch.rand(rows, 1, 64, 64) * 2 - 1
y = torch.arange(40).repeat(E_)
torch.manual_seed(0)
G, D = P.Generator(**cfg), P.Discriminator(**cfg)
sg = {k: v.detach().clone() for k, v in G.state_dict().items()}
sd = {k: v.detach().clone() for k, v in D.state_dict().items()}
G, D = G.cuda().train(), D.cuda().train()
for grp in D.optim.param_groups:
    grp["lr"] = 0.0
train = make_train_step(G, D, P.G_D(G, D), FixedZ(phases), dict(cfg, batch_size=rows))
with noise.replay(replay_list(phases)):
    got = train(x.cuda(), y.cuda())
nz = dict(z_d=phases[0][0], rdof_d=phases[0][1], aug_d=phases[0][2], z_g=phases[1][0], rdof_g=phases[1][1], aug_g=phases[1][2])
want = O.train_step(sg, sd, dict(cfg, device="cpu"), x, y, nz)
print("losses", got, want)
for tag, net, ref in (("G", G, sg), ("D", D, sd)):
    rs = []
    for k, p in net.named_parameters():
        if float(ref[k].grad.norm()) > 1e-7:
            rs.append((rel(p.grad, ref[k].grad), k, float(p.grad.norm()), float(ref[k].grad.norm())))
    rs.sort(reverse=True)
    print(tag, "worst:", rs[:6], "median", rs[len(rs) // 2][0])
