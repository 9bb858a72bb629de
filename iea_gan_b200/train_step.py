"""Host-side driver of one G+D training step on the B200 modules.

The reference's own step function (train_fns.py:20-206) runs unchanged on the
drop-in modules; this file is the repo's self-contained
equivalent for the shipped configuration (Contra head, split_D, DiffAugment,
IEA + uniformity losses, toggle_grads, G ortho-reg, EMA) so that tests and
bench.py need nothing from the reference tree.  Same order of operations, same
RNG draws per phase (z_.sample_(), randn rdof inside G, 7 DiffAugment draws).
"""
import torch

from . import losses


def toggle_grad(model, on):
    """utils.toggle_grad (utils/__init__.py); also takes a cached parameter list (the module-tree walk of
    .parameters() costs more host time than the whole optimizer step at this launch rate)."""
    for p in (model if isinstance(model, (list, tuple)) else model.parameters()):
        p.requires_grad = on


class NormalNoise:
    """z ~ N(0, var) refreshed in place by sample_() (utils.Distribution 'normal', utils/__init__.py:78-86)."""

    def __init__(self, rows, dim, device, var=1.0):
        self.t = torch.empty(rows, dim, device=device)
        self.var = var
        self.sample_()

    def sample_(self):
        self.t.normal_(0, self.var)  # the reference passes the variance as std, too (utils/__init__.py:86)
        return self.t


from .optim import OrthoReg, ortho as ortho_  # noqa: E402,F401  (utils.ortho as grouped kernels)
from .optim import FusedEMA as EMA  # noqa: E402  (utils.apply_ema as one multi-tensor launch)


def check_config(config):
    """This step function is the shipped configuration's path through train_fns.train (Contra head, split_D,
    toggled grads, one D step, no accumulation, no consistency regularisation).  Any other setting changes the
    objective the reference would optimise, so it is refused here instead of silently ignored; the reference's
    own train_fns.py runs on the drop-in modules for those."""
    want = {"num_D_steps": 1, "num_D_accumulations": 1, "num_G_accumulations": 1, "toggle_grads": True,
            "conditional_strategy": "Contra", "split_D": True, "Con_reg": False}
    bad = {k: config[k] for k, v in want.items() if k in config and config[k] != v}
    if bad:
        raise NotImplementedError("make_train_step is built for %s; got %s -- use the reference's train_fns.py on "
                                  "the drop-in modules for other settings" % (want, bad))


GRAPH_SUPPORTED = True  # make_train_step(cuda_graph=True): bench.py probes this


def make_train_step(G, D, GD, z_, config, ema=None, state=None, grad_hook=None, cuda_graph=False, graph_warmup=3):
    """Returns train(x, y) -> dict of the five floats train_fns.train returns.
    grad_hook(net) is called after each backward (explicit data-parallel all-reduce point; nets prepared with
    dp.attach() need none).

    cuda_graph=True: after `graph_warmup` eager steps (which size the caches and the allocator) the WHOLE step --
    both forwards and backwards, the losses, clip + Adam, EMA, and under dp.attach() the NCCL all-reduces -- is
    captured once into a CUDA graph and replayed: ~2000 kernel launches become one graph launch, which removes the
    host launch floor (81 ms at one event per step).  Shapes must then stay fixed; the random draws come from the
    CUDA generator as before (its Philox offset advances per replay); learning-rate / EMA-decay changes are picked
    up from device memory without re-capture."""
    check_config(config)
    contra = losses.Conditional_Contrastive_loss(None, config["batch_size"], config["pos_collected_numerator"])
    use_unif = bool(config.get("Uniformity_loss", True))
    use_iea = bool(config.get("IEA_loss", True))
    state = state if state is not None else {"itr": 0}
    g_params, d_params = list(G.parameters()), list(D.parameters())  # walked once, not ten times per step
    g_black = list(G.shared.parameters())
    d_ortho = OrthoReg(d_params, config["D_ortho"]) if config.get("D_ortho", 0.0) > 0.0 else None
    g_ortho = OrthoReg(g_params, config["G_ortho"], g_black) if config.get("G_ortho", 0.0) > 0.0 else None
    names = ("G_loss", "D_loss_real", "D_loss_fake", "unif_loss_d", "iea_loss")

    micro = int(config.get("micro_events", 0)) * 40  # rows per micro-batch (0: the whole batch at once)
    syncs = [n_.__dict__.get("_iea_grad_sync") for n_ in (D, G)]

    def set_sync(on):
        for sy in syncs:
            if sy is not None:
                sy.enabled = on

    def body(x, y):
        """One step; returns the (5,) device tensor of reported losses (no host synchronisation inside).
        With config["micro_events"] = m the batch is processed m events at a time and the gradients accumulate in
        the flat buffers (one optimizer step, one data-parallel all-reduce per net): same result, bounded memory --
        how 32 events per GPU (BASELINE configs[3] on 2 GPUs) fit."""
        n = x.shape[0]
        mb = micro if 0 < micro < n else n
        if n % mb:
            raise ValueError("%d rows do not split into micro-batches of %d" % (n, mb))
        chunks = [slice(r, r + mb) for r in range(0, n, mb)]
        w = 1.0 / len(chunks)
        G.optim.zero_grad()
        D.optim.zero_grad()
        toggle_grad(d_params, True)
        toggle_grad(g_params, False)
        t = 1.0
        # ---- D step (train_fns.py:49-139)
        z = z_.sample_()
        ers, rep = [], None
        for ci, sl in enumerate(chunks):
            set_sync(ci == len(chunks) - 1)
            pf, ef, d_fake, pr, er, d_real = GD(z[sl], y[sl], x[sl], y[sl], contra=True, train_G=False, split_D=True,
                                                diff_aug=config["diff_aug"])
            l_real, l_fake = losses.loss_hinge_dis(d_fake, d_real)
            d_loss = l_real + l_fake + config["contra_lambda"] * contra(er, pr, None, y[sl], t, 0)
            unif_d = torch.zeros((), device=d_loss.device)
            if use_unif:  # train_fns.py:122-124
                unif_d = losses.unif_loss(er)
                d_loss = d_loss + config["unif_lambda"] * unif_d
            (d_loss * w if len(chunks) > 1 else d_loss).backward()
            ers.append(er)
            r = torch.stack([l_real.detach(), l_fake.detach(), unif_d.detach()]) * w
            rep = r if rep is None else rep + r
        if grad_hook is not None:
            grad_hook(D)
        if d_ortho is not None:  # utils.ortho(D, D_ortho), train_fns.py:133-134
            d_ortho.apply()
        D.optim.step(clip_norm=config.get("clip_norm"))  # clip_grad_norm_ + Adam, fused (train_fns.py:136-139)
        # ---- G step (train_fns.py:142-192)
        toggle_grad(d_params, False)
        toggle_grad(g_params, True)
        G.optim.zero_grad()
        z = z_.sample_()
        repg = None
        for ci, sl in enumerate(chunks):
            set_sync(ci == len(chunks) - 1)
            pf, ef, d_fake = GD(z[sl], y[sl], contra=True, train_G=True, split_D=True, diff_aug=config["diff_aug"])
            g_loss = losses.loss_hinge_gen(d_fake) + config["contra_lambda"] * contra(ef, pf, None, y[sl], t, 0)
            iea_l = torch.zeros((), device=g_loss.device)
            if use_iea:  # train_fns.py:169-176: the G uniformity term sits INSIDE the IEA branch
                iea_l = losses.IEA_loss(ef, ers[ci])
                g_loss = g_loss + config["IEA_lambda"] * iea_l
                if use_unif:
                    g_loss = g_loss + config["unif_lambda"] * losses.unif_loss(ef)
            (g_loss * w if len(chunks) > 1 else g_loss).backward()
            r = torch.stack([g_loss.detach(), iea_l.detach()]) * w
            repg = r if repg is None else repg + r
        if grad_hook is not None:
            grad_hook(G)
        if g_ortho is not None:  # utils.ortho(G, G_ortho, blacklist=G.shared), train_fns.py:185-188
            g_ortho.apply()
        if config.get("clip_norm") is not None:  # the reference only steps G inside this branch (train_fns.py:190-192)
            G.optim.step(clip_norm=config["clip_norm"])
        if ema is not None:
            ema.update(state["itr"])
        return torch.stack([repg[0], rep[0], rep[1], rep[2], repg[1]])

    if not cuda_graph:
        def train(x, y):
            return dict(zip(names, body(x, y).tolist()))
        return train

    gs = {"calls": 0, "graph": None}

    def train(x, y):
        if gs["graph"] is None:
            gs["calls"] += 1
            if gs["calls"] <= graph_warmup:
                # warm-up on the SAME side stream the capture will use: autograd binds a parameter's gradient
                # accumulator node to the stream that was current when the node was built, and at the end of a
                # backward pass the caller's stream waits for every such leaf stream -- which a capturing stream may
                # only do for itself ("dependency created on uncaptured work in another stream" otherwise).
                side = gs.setdefault("side", torch.cuda.Stream())
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    vals = body(x, y)
                torch.cuda.current_stream().wait_stream(side)
                return dict(zip(names, vals.tolist()))
            # capture (the capture itself executes nothing: the replay below performs this call's step)
            from . import engine
            gs["x"], gs["y"] = x.detach().clone(), y.detach().clone()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            engine.GRAPH_KEEP = keep = []  # pinned host tables referenced by captured copies stay alive with the graph
            l0 = engine.LAUNCHES[0]
            try:
                # relaxed: the backward passes stage their job tables through pinned host buffers, and a pinned
                # allocation (cudaHostAlloc) is an 'unsafe' call under the default global capture mode
                with torch.cuda.graph(g, stream=gs.setdefault("side", torch.cuda.Stream()), capture_error_mode="relaxed"):
                    gs["vals"] = body(gs["x"], gs["y"])
            finally:
                engine.GRAPH_KEEP = None
            gs["graph"], gs["keep"] = g, keep
            train.launches_per_step = engine.LAUNCHES[0] - l0  # kernels of this repo inside one replay
        else:
            gs["x"].copy_(x, non_blocking=True)
            gs["y"].copy_(y, non_blocking=True)
        # everything host-dependent that the captured kernels read from device memory
        for o in (G.optim, D.optim):
            o.refresh_hyper()
        if ema is not None:
            ema.refresh_hyper(state["itr"])
        gs["graph"].replay()
        return dict(zip(names, gs["vals"].tolist()))
    train.launches_per_step = None
    return train
