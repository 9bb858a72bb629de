"""The CUDA-graph form of the train step against the eager one.  (Last in collection order on purpose: a failed
stream capture leaves the CUDA context unusable for whatever runs after it in the same process.)"""
import os

import pytest
import torch

from test_gpu_fullsize import draws_for, replay_list, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_cuda_graph_step_equals_eager(small_cfg):
    """make_train_step(cuda_graph=True): three eager warm-up steps, capture, replays -- must reproduce the eager
    step function exactly (same kernels in the same order): the five losses of every step and every parameter
    and buffer after 6 steps, with the same (device-resident) noise handed to both through noise.replay."""
    import iea_gan_b200 as P
    from iea_gan_b200 import noise
    from iea_gan_b200.train_step import make_train_step, EMA
    cfg = dict(small_cfg, device="cuda")
    phases = draws_for(cfg, 601, 40, 64, 64)
    dev_ph = [(z.cuda(), rd.cuda(), {k: v.cuda() for k, v in d.items()}) for z, rd, d in phases]
    torch.manual_seed(602)
    xs = [(torch.rand(40, 1, 64, 64) * 2 - 1).cuda() for _ in range(6)]
    y = torch.arange(40, device="cuda")

    class Z:
        def __init__(self):
            self.i = 0

        def sample_(self):
            self.i += 1
            return dev_ph[(self.i - 1) % 2][0]
    out = {}
    for graph in (False, True):
        torch.manual_seed(0)
        G, D = P.Generator(**cfg).cuda().train(), P.Discriminator(**cfg).cuda().train()
        G_ema = P.Generator(**dict(cfg, skip_init=True, no_optim=True)).cuda()
        ema = EMA(G, G_ema, 0.999, 0)
        train = make_train_step(G, D, P.G_D(G, D), Z(), cfg, ema=ema, cuda_graph=graph)
        traj = []
        with noise.replay(replay_list(dev_ph) * 6):
            for x in xs:
                traj.append(train(x, y))
        torch.cuda.synchronize()
        out[graph] = (traj, {k: v.clone() for k, v in G.state_dict().items()}, {k: v.clone() for k, v in D.state_dict().items()},
                      {k: v.clone() for k, v in G_ema.state_dict().items()})
    (ta, ga, da, ea), (tb, gb, db, eb) = out[False], out[True]
    for a, b in zip(ta, tb):
        for k in a:
            assert abs(a[k] - b[k]) <= 1e-6 * max(1.0, abs(a[k])), (k, a[k], b[k])
    for sa, sb in ((ga, gb), (da, db), (ea, eb)):
        for k in sa:
            assert rel(sb[k].float(), sa[k].float()) < 1e-6, k
