"""Discriminator, DiffAugment, losses and the full G+D train step on the B200 against the
reference-run golden vectors and the CPU oracle."""
import os

import pytest
import torch

from iea_gan_b200 import noise

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def nets(cfg, seed=0):
    import iea_gan_b200 as P
    torch.manual_seed(seed)
    G = P.Generator(**cfg)
    D = P.Discriminator(**cfg)
    sg = {k: v.detach().clone() for k, v in G.state_dict().items()}
    sd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    return G.cuda(), D.cuda(), sg, sd


# fp32 activations: 3e-4 (summation order); bf16 activations: 5e-2 on embeddings / logits
@pytest.mark.parametrize("adt,tol", [("fp32", 3e-4), ("bf16", 5e-2)])
def test_discriminator_forward_vs_golden(small_cfg, golden_fwd, adt, tol):
    os.environ["IEA_ACT_DTYPE"] = adt
    try:
        cfg = dict(small_cfg, device="cuda")
        _, D, _, _ = nets(cfg)
        D.train()
        y = torch.arange(40, device="cuda")
        x = golden_fwd["x_real"].cuda()
        with torch.no_grad():
            p, e, o = D(x, y)
            assert rel(p, golden_fwd["d_proxy"]) < 1e-5
            assert rel(e, golden_fwd["d_embed"]) < tol
            assert rel(o, golden_fwd["d_out"]) < tol
            D.blocks[1][2].gamma.fill_(0.7)
            p, e, o = D(x, y)
            assert rel(e, golden_fwd["d_embed_gamma07"]) < tol
            assert rel(o, golden_fwd["d_out_gamma07"]) < tol
    finally:
        os.environ.pop("IEA_ACT_DTYPE", None)


def test_diffaugment_vs_golden_and_grad(golden_fwd):
    from iea_gan_b200.augment import DiffAugment
    from oracle import iea_oracle as O
    for key_in, key_out, seed, shape in (("x_real", "diffaug_out", 104, (40, 64, 64)), ("x_ns", "diffaug_ns_out", 106, (8, 32, 96))):
        x = golden_fwd[key_in]
        torch.manual_seed(seed)
        d = O.diffaug_draws(*shape)  # CPU stream, as in the golden run
        seq = [d["brightness"], d["saturation"], d["contrast"], d["tx"], d["ty"], d["ox"], d["oy"]]
        with noise.replay(seq):
            xg = x.cuda().requires_grad_(True)
            out = DiffAugment(xg, policy="color,translation,cutout")
        assert rel(out, golden_fwd[key_out]) < 1e-6
        g = torch.randn_like(out)
        out.backward(g)
        xr = x.clone().requires_grad_(True)
        O.diffaugment(xr, d).backward(g.cpu())
        assert rel(xg.grad, xr.grad) < 1e-5


def test_losses_vs_golden_and_grad(golden_fwd):
    from iea_gan_b200 import losses as LS
    from oracle import iea_oracle as O
    e, p, o, ef = (golden_fwd[k] for k in ("d_embed", "d_proxy", "d_out", "embed_fake_rand"))
    crit = LS.Conditional_Contrastive_loss("cuda", 40, False)
    eg, pg = e.cuda().requires_grad_(True), p.cuda().requires_grad_(True)
    lc = crit(eg, pg, None, None, 1.0, 0)
    assert abs(float(lc) - float(golden_fwd["loss_contra"])) < 1e-5
    lc.backward()
    er, pr = e.clone().requires_grad_(True), p.clone().requires_grad_(True)
    O.contrastive(er, pr).backward()
    assert rel(eg.grad, er.grad) < 1e-4 and rel(pg.grad, pr.grad) < 1e-4
    eg = e.cuda().requires_grad_(True)
    lu = LS.unif_loss(eg)
    assert abs(float(lu) - float(golden_fwd["loss_unif"])) < 1e-5
    lu.backward()
    er = e.clone().requires_grad_(True)
    O.uniformity(er).backward()
    assert rel(eg.grad, er.grad) < 1e-4
    fg = ef.cuda().requires_grad_(True)
    li = LS.IEA_loss(fg, e.cuda())
    assert abs(float(li) - float(golden_fwd["loss_iea"])) < 1e-6
    li.backward()
    fr = ef.clone().requires_grad_(True)
    O.iea(fr, e).backward()
    assert rel(fg.grad, fr.grad) < 1e-4
    og = o.cuda().requires_grad_(True)
    a, b = LS.loss_hinge_dis(og * 3 - 0.5, og * 2 + 0.3)
    c = LS.loss_hinge_gen(og)
    assert torch.allclose(torch.stack([a, b, c]).cpu(), golden_fwd["loss_hinge"], atol=1e-6)
    (a + 2 * b + 3 * c).backward()
    orr = o.clone().requires_grad_(True)
    a2, b2 = O.hinge_dis(orr * 3 - 0.5, orr * 2 + 0.3)
    (a2 + 2 * b2 + 3 * O.hinge_gen(orr)).backward()
    assert rel(og.grad, orr.grad) < 1e-5
    # two events: mean of per-event losses
    e2 = torch.cat([e, torch.nn.functional.normalize(torch.randn(40, 1024, generator=torch.Generator().manual_seed(5)), dim=1)])
    assert abs(float(LS.unif_loss(e2.cuda())) - float(O.uniformity(e2))) < 1e-5


@pytest.mark.parametrize("adt", ["fp32", "bf16"])
def test_train_step_vs_unmodified_train_fns(small_cfg, golden_step, adt):
    """One full D step + G step (incl. Adam on D between them): the five returned floats, every
    parameter's gradient norm and selected gradients against the golden run of the reference's
    unmodified train_fns.train.  Tolerances: fp32 activations 1e-3 on losses / 1e-2 relative on
    gradient norms; bf16 activations 3e-2 / 0.15 (gradients through 50+ bf16 layers)."""
    import iea_gan_b200 as P
    from iea_gan_b200.train_step import make_train_step
    from oracle import iea_oracle as O
    os.environ["IEA_ACT_DTYPE"] = adt
    try:
        cfg = dict(small_cfg, device="cuda")
        G, D, _, _ = nets(cfg)
        G.train(); D.train()
        GD = P.G_D(G, D)
        # the golden run's CPU random stream, replayed in order
        torch.manual_seed(202)
        draws = []
        for ph in range(2):
            z = torch.empty(40, cfg["dim_z"]).normal_(0, 1)
            rd = torch.randn(40, cfg["rdof_dim"])
            d = O.diffaug_draws(40, 64, 64)
            draws.append((z, rd, [d["brightness"], d["saturation"], d["contrast"], d["tx"], d["ty"], d["ox"], d["oy"]]))
        zs = iter([d[0] for d in draws])

        class Z:
            def sample_(self):
                return next(zs).cuda()
        with noise.replay([t for _, rd, aug in draws for t in [rd] + aug]):  # per phase: rdof, then the 7 DiffAugment draws
            train = make_train_step(G, D, GD, Z(), cfg)
            losses = train(golden_step["x"].cuda(), torch.arange(40, device="cuda"))
        ltol, gtol, ftol = (1e-3, 1e-2, 2e-2) if adt == "fp32" else (3e-2, 0.15, 0.2)
        for k, v in golden_step["losses"].items():
            assert abs(losses[k] - v) < ltol * max(1.0, abs(v)), (k, losses[k], v)
        bad = []
        # conv biases in front of a batch-norm have an exactly-zero gradient; with bf16 storage the
        # kernel's value is rounding noise (~1e-4), hence the absolute floor
        atol = 1e-6 if adt == "fp32" else 2e-3
        for tag, net, norms in (("G", G, golden_step["g_grad_norm"]), ("D", D, golden_step["d_grad_norm"])):
            for k, p in net.named_parameters():
                got = float(p.grad.norm())
                if abs(got - norms[k]) > gtol * max(norms[k], 1e-6) + atol:
                    bad.append((tag, k, got, norms[k]))
        assert not bad, bad[:10]
        for net, grads in ((G, golden_step["g_grads"]), (D, golden_step["d_grads"])):
            ps = dict(net.named_parameters())
            for k, v in grads.items():
                assert rel(ps[k].grad.reshape(-1)[:65536], v) < ftol, k
        for k, v in golden_step["g_buffers"].items():
            assert rel(G.state_dict()[k], v) < (1e-4 if adt == "fp32" else 3e-2), k
        for k, v in golden_step["d_buffers"].items():
            assert rel(D.state_dict()[k], v) < (1e-4 if adt == "fp32" else 3e-2), k
    finally:
        os.environ.pop("IEA_ACT_DTYPE", None)


@pytest.mark.parametrize("adt,t_out,t_w,t_dx", [("bf16", 1e-2, 1.2e-1, 2.5e-1), ("fp32", 1e-5, 1e-3, 1e-2)])
def test_full_size_discriminator_forward_backward_vs_oracle(adt, t_out, t_w, t_dx):
    """BASELINE.json's shape (40 images of 256x256, shipped widths): the CUDA Discriminator -- stem on the
    CUDA-core kernel, TMA-fed macro-tile tcgen05 convs, resident / streaming tcgen05 convs, tcgen05 attention
    (1024 x 256, gamma != 0), cluster split-K head linears, macro-tile weight gradients -- against the fp32 CPU
    oracle: the three outputs, the input gradient and weight gradients from the top, the middle (attention) and
    the bottom of the net.
    Tolerances.  fp32 activations (CUDA-core kernels, same arithmetic as the oracle in a different summation
    order): outputs 1e-5, weight gradients 1e-3, input gradient 1e-2 -- the input gradient is per pixel and
    passes 75 ReLU masks, a handful of which flip even at fp32 rounding level (measured 2.6e-3).  bf16
    activations and gradients (the tensor-core path): outputs 1e-2 (measured 3.5e-3), weight gradients 1.2e-1
    (measured 2-8 %, largest for the attention theta conv), input gradient 2.5e-1 (measured 0.18: no averaging
    over pixels, and ~0.2 % of the ReLU masks of every layer differ from the fp32 forward)."""
    import iea_gan_b200 as P
    from iea_gan_b200.default_config import shipped_config
    from oracle import iea_oracle as O
    cfg = shipped_config(H_base=1, device="cuda")
    torch.manual_seed(0)
    D = P.Discriminator(**cfg)
    with torch.no_grad():
        D.blocks[2][2].gamma.fill_(0.5)  # the attention block must contribute
    sd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    torch.manual_seed(21)
    x = torch.rand(40, 1, 256, 256) * 2 - 1
    y = torch.arange(40)
    go, ge = torch.randn(40), torch.randn(40, cfg["hypersphere_dim"])
    names = ["input_conv.weight", "linear0.weight", "linear1.weight", "blocks.0.0.conv2.weight",
             "blocks.2.2.theta.weight", "blocks.3.1.conv3.weight", "blocks.5.0.conv1.weight"]
    os.environ["IEA_ACT_DTYPE"] = adt
    try:
        D = D.cuda().train()
        xg = x.cuda().requires_grad_(True)
        p, e, o = D(xg, y.cuda())
        ((o * go.cuda()).sum() + (e * ge.cuda()).sum()).backward()
        torch.cuda.synchronize()
    finally:
        os.environ.pop("IEA_ACT_DTYPE", None)
    got = {n: dict(D.named_parameters())[n].grad.detach().cpu() for n in names}
    # oracle (fp32, CPU, autograd through the restated forward)
    for n in names:
        sd[n].requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    pr, er, orr = O.discriminator_forward(sd, dict(cfg, device="cpu"), xr, y, training=True)
    ((orr * go).sum() + (er * ge).sum()).backward()
    assert rel(p.detach(), pr) < 1e-5
    assert rel(e.detach(), er) < t_out
    assert rel(o.detach(), orr) < t_out
    assert rel(xg.grad, xr.grad) < t_dx
    for n in names:
        assert rel(got[n], sd[n].grad) < t_w, n


def test_full_size_discriminator_pooled_once_shortcut_fp32():
    """The opt-in DBlock shortcut that pools x once (IEA_DBLOCK_POOL_ONCE=1: iea_avgpool2_fwd, conv_sc on the low
    resolution window, conv4 with a plain residual, one un-pooling adjoint) is the same function: with fp32 activations
    it meets the tolerances of the default path against the CPU oracle."""
    os.environ["IEA_DBLOCK_POOL_ONCE"] = "1"
    try:
        test_full_size_discriminator_forward_backward_vs_oracle("fp32", 1e-5, 1e-3, 1e-2)
    finally:
        os.environ.pop("IEA_DBLOCK_POOL_ONCE", None)
