"""Correctness AT THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs 1-3: shipped widths, 256x256):

  * a full G+D train step on two events against the fp32 CPU oracle -- the five losses, the gradient of every
    parameter of G (the full-size Generator backward) and of D, the buffers (fp32 activations tight, bf16 stated);
  * the E = 8 train step bench.py times and the E = 16 sampling batch: one batched call must equal the same
    events run one at a time on the GPU (per-event batch-norm groups, the fused finalize's ticket path, per-event
    loss means, grouped SN backward) -- the size-independent property that extends the oracle pin from E <= 2 to
    the benchmarked batch sizes without minutes of CPU time;
  * a 20-step loss trajectory, bf16 against fp32 activations, on the small configuration: the justification
    for accepting the bf16 gradient tolerances.
"""
import os

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


def draws_for(cfg, seed, rows, hh, ww):
    from oracle import iea_oracle as O
    g = torch.Generator().manual_seed(seed)
    ph = []
    for _ in range(2):
        z = torch.randn(rows, cfg["dim_z"], generator=g)
        rd = torch.randn(rows, cfg["rdof_dim"], generator=g)
        ph.append((z, rd, O.diffaug_draws(rows, hh, ww, generator=g)))
    return ph


def replay_list(phases, rows=None):
    out = []
    for _, rd, d in phases:
        s = slice(None) if rows is None else rows
        out.append(("randn", rd[s]))
        out += [("rand", d[k][s]) for k in ("brightness", "saturation", "contrast")]
        out += [("randint", d[k][s]) for k in ("tx", "ty", "ox", "oy")]
    return out


class FixedZ:
    def __init__(self, phases, rows=None):
        self.it = iter([p[0] if rows is None else p[0][rows] for p in phases])

    def sample_(self):
        return next(self.it).cuda()


def fresh_nets(cfg, seed=0):
    import iea_gan_b200 as P
    torch.manual_seed(seed)
    G, D = P.Generator(**cfg), P.Discriminator(**cfg)
    with torch.no_grad():
        for m in D.modules():
            if hasattr(m, "gamma") and isinstance(m.gamma, torch.nn.Parameter):
                m.gamma.fill_(0.5)  # the attention block must contribute
    sg = {k: v.detach().clone() for k, v in G.state_dict().items()}
    sd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    return G.cuda().train(), D.cuda().train(), sg, sd


def gpu_step(cfg, phases, x, y, adt, rows=None, nets=None):
    import iea_gan_b200 as P
    from iea_gan_b200 import noise
    from iea_gan_b200.train_step import make_train_step
    os.environ["IEA_ACT_DTYPE"] = adt
    try:
        G, D, _, _ = nets if nets is not None else fresh_nets(cfg)
        n = x.shape[0]
        train = make_train_step(G, D, P.G_D(G, D), FixedZ(phases, rows), dict(cfg, batch_size=n))
        with noise.replay(replay_list(phases, rows)):
            losses = train(x.cuda(), y.cuda())
        torch.cuda.synchronize()
    finally:
        os.environ.pop("IEA_ACT_DTYPE", None)
    return G, D, losses


@pytest.fixture(scope="module")
def oracle_two_events():
    """fp32 CPU oracle of one full-size G+D step on two events (about a minute of host time), shared by both dtypes."""
    from iea_gan_b200.default_config import shipped_config
    from oracle import iea_oracle as O
    import iea_gan_b200 as P
    cfg = shipped_config(H_base=1, device="cuda", clip_norm=1e9)
    rows = 80
    phases = draws_for(cfg, 401, rows, 256, 256)
    torch.manual_seed(402)
    x = torch.rand(rows, 1, 256, 256) * 2 - 1
    x[x < 0.6] = -1.0  # sparse, PXD-like occupancy
    y = torch.arange(40).repeat(2)
    torch.manual_seed(0)
    ccfg = dict(cfg, device="cpu")
    Gc, Dc = P.Generator(**ccfg), P.Discriminator(**ccfg)
    with torch.no_grad():
        for m in Dc.modules():
            if hasattr(m, "gamma") and isinstance(m.gamma, torch.nn.Parameter):
                m.gamma.fill_(0.5)
    sg = {k: v.detach().clone() for k, v in Gc.state_dict().items()}
    sd = {k: v.detach().clone() for k, v in Dc.state_dict().items()}
    nz = dict(z_d=phases[0][0], rdof_d=phases[0][1], aug_d=phases[0][2], z_g=phases[1][0], rdof_g=phases[1][1],
              aug_g=phases[1][2])
    torch.set_num_threads(os.cpu_count() or 1)
    want = O.train_step(sg, sd, ccfg, x, y, nz)  # no optimizers: D is not stepped between the phases (see below)
    return cfg, phases, x, y, sg, sd, want


def grad_report(net, ref, attr=False, floor=1e-6):
    """[(name, rel-L2 error, cosine)] of every parameter gradient against `ref[name]`; gradients that are exactly
    zero in the reference (conv biases in front of a batch-norm: the mean subtraction cancels them) are skipped --
    what the kernels produce there is rounding noise around zero."""
    out = []
    for k, p in net.named_parameters():
        if k.startswith("blocks.") and ".conv" in k and k.endswith(".bias") and hasattr(net, "shared"):
            continue  # G's block conv biases all sit in front of a batch-norm: true gradient exactly zero
        rg = ref[k].grad if attr else ref[k]
        a, b = p.grad.detach().double().cpu().reshape(-1), rg.detach().double().cpu().reshape(-1)
        if float(b.norm()) < floor * max(1.0, float(b.numel()) ** 0.5):
            continue
        out.append((k, float((a - b).norm() / b.norm()), float(a @ b / (a.norm() * b.norm() + 1e-300))))
    return out


@pytest.mark.parametrize("adt", ["fp32", "bf16"])
def test_full_size_train_step_two_events_vs_oracle(oracle_two_events, adt):
    """Tolerances.  fp32 activations: losses 1e-3, EVERY parameter gradient of G and D within 2e-2 relative L2
    (a handful of ReLU masks flip at fp32 rounding level in the 150-layer G -> DiffAugment -> D chain), buffers 1e-4.
    bf16 activations / gradients: losses 3e-2; D's gradients (D phase) within 12 % (30 % for the attention
    theta / phi convs, whose logits feed an un-scaled softmax); G's gradients: cosine >= 0.9 with the fp32
    gradient (relative L2 <= 0.45).  Why G is looser: at RANDOM INIT (the only weights that exist for this
    benchmark) the 48-batch-norm generator amplifies any perturbation by ~1.5x per layer (tools/dbg/
    batch_dep_forward.py: a 3e-7 change of the conditioning vector becomes 2e-2 on the image), so the bf16
    rounding of every layer saturates at ~2 % on the image and ~15 % on gradients that crossed D and G backward;
    test_loss_trajectory_bf16_tracks_fp32 is the check that training is unaffected.
    D's optimizer is given lr = 0 so that both sides run the G phase on the same D weights."""
    cfg, phases, x, y, sg, sd, want = oracle_two_events
    G, D, _, _ = fresh_nets(cfg)
    for grp in D.optim.param_groups:
        grp["lr"] = 0.0
    G, D, got = gpu_step(cfg, phases, x, y, adt, nets=(G, D, None, None))
    ltol = 1e-3 if adt == "fp32" else 3e-2
    for k, v in want.items():
        assert abs(got[k] - v) < ltol * max(1.0, abs(v)), (k, got[k], v)
    rg, rd_ = grad_report(G, sg, True), grad_report(D, sd, True)
    assert len(rg) > 150 and len(rd_) > 100
    if adt == "fp32":
        bad = [t for t in rg + rd_ if t[1] > 2e-2]
    else:
        bad = [t for t in rd_ if t[1] > (0.30 if (".theta." in t[0] or ".phi." in t[0]) else 0.12)]
        bad += [t for t in rg if t[2] < 0.9 or t[1] > 0.45]
    assert not bad, (len(bad), bad[:8])
    btol = 1e-4 if adt == "fp32" else 3e-2
    for net, ref in ((G, sg), (D, sd)):
        st = net.state_dict()
        for k in ref:
            if k.endswith(("u0", "sv0", "stored_mean", "stored_var")):
                assert rel(st[k].float(), ref[k].float()) < btol, k
    print("full-size 2-event step (%s): worst rel-L2 G %.3g D %.3g, lowest cosine G %.4f"
          % (adt, max(t[1] for t in rg), max(t[1] for t in rd_), min(t[2] for t in rg)))


@pytest.mark.parametrize("adt", ["fp32", "bf16"])
def test_benchmarked_batches_equal_per_event_runs(adt):
    """E = 8 train step (BASELINE config 2/3 as bench.py runs it): gradients and losses of ONE batched step ==
    mean over 8 single-event steps started from the same weights and spectral-norm vectors; E = 16 sampling: one
    batched forward == single-event forwards.
    fp32 activations: the event logic itself (per-event batch-norm groups, ticketed finalize, per-event loss
    means, grouped SN backward) -- only fp32 summation order differs: losses 1e-3, gradients 2e-2, images 1e-3.
    bf16 (the benchmarked dtype): D's gradients 1e-2, images 5e-2, G's gradients cosine >= 0.9 -- a different
    batch size changes fp32 summation order in the split-K linears by ~3e-7, which the random-init generator
    amplifies exactly as it amplifies bf16 rounding (see the test above)."""
    from iea_gan_b200.default_config import shipped_config
    from iea_gan_b200 import noise
    cfg = shipped_config(H_base=1, device="cuda", clip_norm=1e9)
    E_ = 8
    rows = 40 * E_
    phases = draws_for(cfg, 501, rows, 256, 256)
    torch.manual_seed(502)
    x = torch.rand(rows, 1, 256, 256) * 2 - 1
    y = torch.arange(40).repeat(E_)
    G, D, _, _ = fresh_nets(cfg)
    for o in (G.optim, D.optim):
        for grp in o.param_groups:
            grp["lr"] = 0.0
    state0 = ({k: v.clone() for k, v in G.state_dict().items()}, {k: v.clone() for k, v in D.state_dict().items()})
    G, D, got = gpu_step(cfg, phases, x, y, adt, nets=(G, D, None, None))
    gb = ({k: p.grad.clone() for k, p in G.named_parameters()}, {k: p.grad.clone() for k, p in D.named_parameters()})
    acc = ({k: torch.zeros_like(v) for k, v in gb[0].items()}, {k: torch.zeros_like(v) for k, v in gb[1].items()})
    lsum = {}
    for e in range(E_):
        G.load_state_dict(state0[0]); D.load_state_dict(state0[1])  # same weights, same u0 for every event
        sl = slice(40 * e, 40 * e + 40)
        _, _, l1 = gpu_step(cfg, phases, x[sl], y[sl], adt, rows=sl, nets=(G, D, None, None))
        for i, net in enumerate((G, D)):
            for k, p in net.named_parameters():
                acc[i][k] += p.grad / E_
        for k, v in l1.items():
            lsum[k] = lsum.get(k, 0.0) + v / E_
    for k, v in lsum.items():
        assert abs(got[k] - v) < (1e-3 if adt == "fp32" else 1e-2) * max(1.0, abs(v)), (k, got[k], v)
    for i, net in enumerate((G, D)):
        for k, p in net.named_parameters():
            p.grad = gb[i][k]
    rg, rd_ = grad_report(G, acc[0]), grad_report(D, acc[1])
    if adt == "fp32":
        bad = [t for t in rg + rd_ if t[1] > 2e-2]
    else:
        bad = [t for t in rd_ if t[1] > 1e-2] + [t for t in rg if t[2] < 0.9]
    assert not bad, (len(bad), bad[:8])
    # ---- sampling, 16 events
    rows = 640
    g = torch.Generator().manual_seed(503)
    z, rd = torch.randn(rows, cfg["dim_z"], generator=g).cuda(), torch.randn(rows, cfg["rdof_dim"], generator=g)
    y = torch.arange(40).repeat(16).cuda()
    os.environ["IEA_ACT_DTYPE"] = adt
    try:
        G.load_state_dict(state0[0])
        with torch.no_grad(), noise.replay([rd]):
            img = G(z, y)
        for e in (0, 7, 15):
            G.load_state_dict(state0[0])
            sl = slice(40 * e, 40 * e + 40)
            with torch.no_grad(), noise.replay([rd[sl]]):
                one = G(z[sl], y[sl])
            assert rel(img[sl], one) < (1e-3 if adt == "fp32" else 5e-2), (e, rel(img[sl], one))
    finally:
        os.environ.pop("IEA_ACT_DTYPE", None)
    print("E=8 batched vs per-event (%s): worst rel-L2 G %.3g D %.3g" % (adt, max(t[1] for t in rg), max(t[1] for t in rd_)))


def test_loss_trajectory_bf16_tracks_fp32(small_cfg):
    """20 optimizer steps from the same seed with bf16 and with fp32 activations (same noise: the CUDA generator is
    re-seeded).  GAN training is itself chaotic, so the two runs can only be compared while the perturbation is
    small and statistically afterwards: the five reported losses agree within 3 % of the loss scale (|value| floor
    1.0) over the first 12 steps (measured: ~1 % up to step 16, then the trajectories separate), the means of the
    G / D losses over steps 10-19 agree within 10 %, nothing is NaN, and the two final parameter vectors are closer
    to each other than half the distance either travelled from the initial point.  This is what a user of the bf16 path accepts in place of per-gradient agreement."""
    import iea_gan_b200 as P
    from iea_gan_b200.train_step import make_train_step, NormalNoise
    cfg = dict(small_cfg, device="cuda")
    runs = {}
    for adt in ("fp32", "bf16"):
        os.environ["IEA_ACT_DTYPE"] = adt
        try:
            torch.manual_seed(0)
            G, D = P.Generator(**cfg).cuda().train(), P.Discriminator(**cfg).cuda().train()
            p0 = torch.cat([p.detach().reshape(-1) for p in list(G.parameters()) + list(D.parameters())]).clone()
            torch.manual_seed(77)
            torch.cuda.manual_seed(77)
            train = make_train_step(G, D, P.G_D(G, D), NormalNoise(40, cfg["dim_z"], "cuda"), cfg)
            y = torch.arange(40, device="cuda")
            traj = []
            for i in range(20):
                x = torch.rand(40, 1, 64, 64, device="cuda") * 2 - 1
                traj.append(train(x, y))
            p1 = torch.cat([p.detach().reshape(-1) for p in list(G.parameters()) + list(D.parameters())])
            runs[adt] = (traj, p0, p1)
        finally:
            os.environ.pop("IEA_ACT_DTYPE", None)
    (ta, p0, pa), (tb, _, pb) = runs["fp32"], runs["bf16"]
    dev = lambda lo, hi: max(abs(a[k] - b[k]) / max(1.0, abs(a[k])) for a, b in zip(ta[lo:hi], tb[lo:hi]) for k in a)
    assert all(v == v for t in tb for v in t.values())
    assert dev(0, 12) < 3e-2, dev(0, 12)
    mean = lambda t, k: sum(o[k] for o in t[10:]) / len(t[10:])
    for k in ("G_loss", "D_loss_real", "D_loss_fake"):
        assert abs(mean(ta, k) - mean(tb, k)) < 0.1 * max(1.0, abs(mean(ta, k))), (k, mean(ta, k), mean(tb, k))
    travelled = float((pa - p0).norm())
    assert travelled > 0 and float((pa - pb).norm()) < 0.5 * travelled
    print("trajectory: max loss deviation steps 0-11 %.3g, steps 12-19 %.3g; |p_bf16 - p_fp32| / |p_fp32 - p_0| = %.3g"
          % (dev(0, 12), dev(12, 20), float((pa - pb).norm()) / travelled))


def test_micro_batched_step_equals_whole_batch(small_cfg):
    """config["micro_events"]: 4 events processed 2 at a time with gradients accumulating in the flat buffers ==
    the 4-event step in one go (fp32 activations; only fp32 summation order differs, which the G -> D chain amplifies
    to a few 1e-3 on G's gradients): losses 1e-5, every gradient 2e-2, the update of the whole parameter vector 2e-2."""
    rows = 160
    cfg = dict(small_cfg, device="cuda", batch_size=rows)
    phases = draws_for(cfg, 701, rows, 64, 64)
    torch.manual_seed(702)
    x = torch.rand(rows, 1, 64, 64) * 2 - 1
    y = torch.arange(40).repeat(4)
    res = []
    for micro in (0, 2):
        c = dict(cfg, micro_events=micro)
        # DiffAugment draws are consumed per G_D call: hand each micro-batch its rows of the same draws
        import iea_gan_b200 as P
        from iea_gan_b200 import noise
        from iea_gan_b200.train_step import make_train_step
        os.environ["IEA_ACT_DTYPE"] = "fp32"
        try:
            G, D, _, _ = fresh_nets(c)
            train = make_train_step(G, D, P.G_D(G, D), FixedZ(phases), c)
            if micro:
                chunks = [slice(0, 80), slice(80, 160)]
                seq = [t for ph in phases for sl in chunks for t in replay_list([ph], sl)]
            else:
                seq = replay_list(phases)
            with noise.replay(seq):
                losses = train(x.cuda(), y.cuda())
        finally:
            os.environ.pop("IEA_ACT_DTYPE", None)
        res.append((losses, {k: p.grad.clone() for k, p in list(G.named_parameters()) + [("D." + k, p) for k, p in D.named_parameters()]},
                    torch.cat([p.detach().reshape(-1) for p in list(G.parameters()) + list(D.parameters())])))
    (la, ga, pa), (lb, gb, pb) = res
    for k in la:
        assert abs(la[k] - lb[k]) < 1e-5 * max(1.0, abs(la[k])), (k, la[k], lb[k])
    scale = max(float(v.norm()) for v in ga.values())
    bad = [(k, rel(gb[k], ga[k])) for k in ga if float(ga[k].norm()) > 1e-6 * scale and rel(gb[k], ga[k]) > 2e-2]
    assert not bad, bad[:8]
    assert rel(pb, pa) < 1e-3
