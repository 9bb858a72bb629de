"""Profiling driver: a few Generator sampling forwards at BASELINE configs[1] size (16 events, 256x256).
Used under `ncu --metrics gpu__time_duration.sum` (launch list) and `ncu --set full -k regex:conv_tc`."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iea_gan_b200 as P
from iea_gan_b200.default_config import shipped_config

ev = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = shipped_config(H_base=1, device="cuda")
torch.manual_seed(0)
G = P.Generator(**cfg).cuda()
G.train()
z = torch.randn(40 * ev, cfg["dim_z"], device="cuda")
y = torch.arange(40, device="cuda").repeat(ev)
for _ in range(reps):
    with torch.no_grad():
        G(z, y)
torch.cuda.synchronize()
print("done")
