"""debug: G forward bf16 full size, E=2 with a duplicated event vs E=1: first layer whose output differs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
os.environ["IEA_ACT_DTYPE"] = sys.argv[1] if len(sys.argv) > 1 else "bf16"
import iea_gan_b200 as P
from iea_gan_b200 import engine as E, noise
from iea_gan_b200.default_config import shipped_config
cfg = shipped_config(H_base=1, device="cuda")
torch.manual_seed(0)
G = P.Generator(**cfg).cuda().train()
st0 = {k: v.clone() for k, v in G.state_dict().items()}
g = torch.Generator().manual_seed(5)
z, rd, y = torch.randn(40, cfg["dim_z"], generator=g).cuda(), torch.randn(40, cfg["rdof_dim"], generator=g), torch.arange(40).cuda()
def run(zz, rr, yy, grad):
    G.load_state_dict(st0)
    E.TRACE = []
    if grad:
        with noise.replay([rr]):
            img = G(zz, yy)
    else:
        with torch.no_grad(), noise.replay([rr]):
            img = G(zz, yy)
    tr, E.TRACE = E.TRACE, None
    return img.detach(), tr
for grad in (False, True):
    a, ta = run(z, rd, y, grad)
    b, tb = run(torch.cat([z, z]), torch.cat([rd, rd]), torch.cat([y, y]), grad)
    print("grad" if grad else "no_grad", "image rel diff ev0 %.3g ev1 %.3g" % (float((b[:40] - a).norm() / a.norm()), float((b[40:] - a).norm() / a.norm())))
    for (tag, ya, sa), (tag2, yb, sb) in zip(ta, tb):
        n = ya.shape[0]
        d0 = float((yb[:n].float() - ya.float()).norm() / (ya.float().norm() + 1e-30))
        d1 = float((yb[n:].float() - ya.float()).norm() / (ya.float().norm() + 1e-30))
        if d0 > 0 or d1 > 0:
            print("  %-60s ev0 %.3g ev1 %.3g" % (tag, d0, d1))
