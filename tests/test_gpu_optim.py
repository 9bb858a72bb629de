"""The per-step parameter update kernels against their torch originals:
FusedAdam (+ fused clip) vs torch.optim.Adam + clip_grad_norm_, FusedEMA vs the reference's apply_ema formula,
optim.ortho vs utils.ortho's formula (utils/__init__.py:843-859) incl. a tall matrix (K-split Gram path)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_fused_adam_and_clip_match_torch():
    from iea_gan_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(128, 64, 3, 3), (16,), (70001,), (8192, 256), ()]
    pa = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = FusedAdam(pa, lr=2e-4, betas=(0.0, 0.999), eps=1e-6)
    ob = torch.optim.Adam(pb, lr=2e-4, betas=(0.0, 0.999), eps=1e-6)
    for it in range(6):
        for p, q in zip(pa, pb):
            g = torch.randn_like(p) * (10.0 ** (it - 2))
            p.grad, q.grad = g.clone(), g.clone()
        if it % 2:
            n_ref = torch.nn.utils.clip_grad_norm_(pb, 3.0)
            oa.step(clip_norm=3.0)
            assert abs(float(oa.grad_norm()) - float(n_ref)) < 1e-4 * float(n_ref)
        else:
            oa.step()
        ob.step()
        if it == 3:  # a learning-rate change reaches the kernel through device memory
            for o in (oa, ob):
                o.param_groups[0]["lr"] = 5e-5
    for p, q in zip(pa, pb):
        assert rel(p, q) < 1e-6
    sa, sb = oa.state_dict()["state"], ob.state_dict()["state"]
    assert set(sa[0].keys()) == set(sb[0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    assert float(sa[0]["step"]) == float(sb[0]["step"]) == 6.0
    assert rel(sa[3]["exp_avg_sq"], sb[3]["exp_avg_sq"]) < 1e-4
    # checkpoint round trip: moments survive load_state_dict and the next step agrees again
    oa.load_state_dict(oa.state_dict())
    for p, q in zip(pa, pb):
        g = torch.randn_like(p)
        p.grad, q.grad = g.clone(), g.clone()
    oa.step(); ob.step()
    for p, q in zip(pa, pb):
        assert rel(p, q) < 1e-6


def test_fused_ema_matches_reference_formula():
    from iea_gan_b200.optim import FusedEMA
    src = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.BatchNorm1d(17)).cuda()
    dst = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.BatchNorm1d(17)).cuda()
    ema = FusedEMA(src, dst, decay=0.9, start_itr=2)
    want = {k: v.clone() for k, v in dst.state_dict().items()}
    for itr in range(5):
        with torch.no_grad():
            for p in src.parameters():
                p.add_(torch.randn_like(p))
            src[1].running_mean.add_(1.0)
        ema.update(itr)
        d = 0.0 if (itr and itr < 2) else 0.9
        for k, v in src.state_dict().items():
            if v.is_floating_point():
                want[k] = want[k] * d + v * (1 - d)
    for k, v in dst.state_dict().items():
        if v.is_floating_point():
            assert rel(v, want[k]) < 1e-6, k


def test_ortho_matches_reference_formula():
    from iea_gan_b200 import optim
    torch.manual_seed(1)
    net = torch.nn.ModuleDict({"emb": torch.nn.Embedding(40, 16), "conv": torch.nn.Conv2d(24, 40, 3), "lin": torch.nn.Linear(64, 4100),
                               "wide": torch.nn.Linear(300, 20)}).cuda()
    for p in net.parameters():
        p.grad = torch.randn_like(p)
    want = {}
    for k, p in net.named_parameters():
        want[k] = p.grad.clone()
        if p.dim() >= 2 and k != "emb.weight":
            w = p.detach().view(p.shape[0], -1).double()
            g = 2 * torch.mm(torch.mm(w, w.t()) * (1.0 - torch.eye(w.shape[0], device="cuda", dtype=torch.double)), w)
            want[k] = want[k] + 1e-2 * g.view(p.shape).float()
    optim.ortho(net, 1e-2, blacklist=[net["emb"].weight])
    for k, p in net.named_parameters():
        assert rel(p.grad, want[k]) < 1e-5, k
