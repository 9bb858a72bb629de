"""debug: FusedAdam vs torch.optim.Adam on random tensors; conv 64->16 1x1 @4x4 bias gradient."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from iea_gan_b200.optim import FusedAdam
torch.manual_seed(0)
shapes = [(128, 64, 3, 3), (16,), (70001,), (8192, 256), ()]
pa = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
oa = FusedAdam(pa, lr=2e-4, betas=(0.0, 0.999), eps=1e-6)
ob = torch.optim.Adam(pb, lr=2e-4, betas=(0.0, 0.999), eps=1e-6)
for it in range(5):
    for p, q in zip(pa, pb):
        g = torch.randn_like(p) * (10.0 ** (it - 2))
        p.grad = g.clone(); q.grad = g.clone()
    if it % 2:
        torch.nn.utils.clip_grad_norm_(pb, 3.0)
        oa.step(clip_norm=3.0)
    else:
        oa.step()
    ob.step()
    print(it, [float((p - q).abs().max() / (q.abs().max() + 1e-30)) for p, q in zip(pa, pb)])
sa, sb = oa.state_dict(), ob.state_dict()
print("state keys", list(sa["state"][0].keys()), float(sa["state"][0]["step"]), float(sb["state"][0]["step"]))
print("exp_avg_sq diff", float((sa["state"][3]["exp_avg_sq"] - sb["state"][3]["exp_avg_sq"]).abs().max()))
