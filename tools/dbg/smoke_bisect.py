"""The train-step half of __graft_entry__.smoke() under a few engine switches: prints every loss next to the CPU restatement."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import iea_gan_b200 as P
from iea_gan_b200 import noise
from iea_gan_b200.train_step import make_train_step
from oracle import iea_oracle as O

with open(os.path.join(ROOT, "tests", "golden", "small_cfg.json")) as f:
    cfg = json.load(f)
cfg["device"] = "cuda"


def run(env):
    os.environ.update(env)
    try:
        torch.manual_seed(0)
        G, D = P.Generator(**cfg), P.Discriminator(**cfg)
        sg = {k: v.detach().clone() for k, v in G.state_dict().items()}
        sd = {k: v.detach().clone() for k, v in D.state_dict().items()}
        G, D = G.cuda().train(), D.cuda().train()
        gen = torch.Generator().manual_seed(5)
        x = torch.rand(40, 1, 64, 64, generator=gen) * 2 - 1
        y = torch.arange(40)
        nz, seq, zs = {}, [], []
        for ph in ("d", "g"):
            nz["z_" + ph] = torch.randn(40, cfg["dim_z"], generator=gen)
            nz["rdof_" + ph] = torch.randn(40, cfg["rdof_dim"], generator=gen)
            nz["aug_" + ph] = a = O.diffaug_draws(40, 64, 64, generator=gen)
            zs.append(nz["z_" + ph])
            seq += [nz["rdof_" + ph]] + [a[k] for k in ("brightness", "saturation", "contrast", "tx", "ty", "ox", "oy")]

        class Z:
            def sample_(self):
                return zs.pop(0).cuda()
        train = make_train_step(G, D, P.G_D(G, D), Z(), cfg)
        with noise.replay(seq):
            out = train(x.cuda(), y.cuda())
        adam = lambda st, names, lr, b1, b2: torch.optim.Adam([st[k] for k in names], lr=lr, betas=(b1, b2), weight_decay=0,
                                                              eps=cfg["adam_eps"])
        want = O.train_step(sg, sd, dict(cfg, device="cpu"), x, y, nz,
                            opt_g=adam(sg, O.param_names(sg), cfg["G_lr"], cfg["G_B1"], cfg["G_B2"]),
                            opt_d=adam(sd, O.param_names(sd), cfg["D_lr"], cfg["D_B1"], cfg["D_B2"]))
        gn = float(dict(G.named_parameters())["linear.weight"].grad.norm())
        rn = float(sg["linear.weight"].grad.norm())
        return out, want, gn, rn
    finally:
        for k in env:
            os.environ.pop(k, None)


for tag, env in (("default", {}), ("fp32", {"IEA_ACT_DTYPE": "fp32"})):
    try:
        out, want, gn, rn = run(env)
        print(tag, {k: round(float(out[k]), 4) for k in want}, "| want", {k: round(float(v), 4) for k, v in want.items()}, "| G.linear grad", round(gn, 4), round(rn, 4), flush=True)
    except Exception as e:
        print(tag, "ERROR", repr(e)[:300], flush=True)
