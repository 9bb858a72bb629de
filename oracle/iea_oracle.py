"""CPU oracle for the IEA-GAN Generator/Discriminator hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg may import it, and only as the checker or the timed CPU
baseline -- never as the thing shipped.

It is a functional restatement (plain fp32 torch ops on a flat name->tensor
state dict that uses the reference's state-dict keys) of the algorithm in the
reference's model.py / layers.py / RRM.py / diff_aug.py / loss.py /
train_fns.py.  The arithmetic itself lives in third-party PyTorch (the
reference pins torch==1.11.0, requirements.txt:160; this image has 2.11.0).

Parity pin: tests/golden/make_golden.py imports the real reference from
/root/reference (build container only) and stores its outputs on seeded
inputs; tests/test_oracle_golden.py checks this file against those vectors.
The reference's own tests pin shapes only (tests/test_image_gen.py:36-38), so
these generated vectors are the only numerical pin ("parity pinned by
reference-run fixtures").

Multi-event extension: a batch of E events is E
independent reference forwards that share weights and share the spectral-norm
vectors of that call; batch-norm statistics, the RRM, and the contrastive /
IEA / uniformity losses are per event (groups of 40 consecutive rows).
"""
import math

import torch
import torch.nn.functional as F

IMGS = 40  # sensors per event: model.py:466 hard-wires 40 rows


# --------------------------------------------------------------------------
# architecture tables (model.py:74-136 G_arch, model.py:561-621 D_arch)
# --------------------------------------------------------------------------
_G_MULT = {
    512: ([16, 16, 8, 8, 4, 2, 1], [16, 8, 8, 4, 2, 1, 1]),
    256: ([16, 16, 8, 8, 4, 2], [16, 8, 8, 4, 2, 1]),
    128: ([16, 16, 8, 4, 2], [16, 8, 4, 2, 1]),
    64: ([16, 16, 8, 4], [16, 8, 4, 2]),
    32: ([4, 4, 4], [4, 4, 4]),
}
_D_MULT = {
    256: ([1, 2, 4, 8, 8, 16], [2, 4, 8, 8, 16, 16]),
    128: ([1, 2, 4, 8, 16], [2, 4, 8, 16, 16]),
    64: ([1, 2, 4, 8], [2, 4, 8, 16]),
}


def g_channels(cfg):
    i, o = _G_MULT[cfg["resolution"]]
    ch = cfg["G_ch"]
    return [ch * m for m in i], [ch * m for m in o]


def d_channels(cfg):
    i, o = _D_MULT[cfg["resolution"]]
    ch = cfg["D_ch"]
    return [ch * m for m in i], [ch * m for m in o]


def d_attention_stages(cfg):
    """Stage indices after which D carries a self-attention block
    (model.py:756-766: resolution list is res/2, res/4, ...)."""
    wanted = [int(t) for t in str(cfg.get("D_attn", "0")).split("_")]
    n = len(_D_MULT[cfg["resolution"]][0])
    return [s for s in range(n) if (cfg["resolution"] >> (s + 1)) in wanted]


# --------------------------------------------------------------------------
# spectral norm (layers.py:89-111 power_iteration, layers.py:151-165 SN.W_)
# --------------------------------------------------------------------------
def sn_weight(sd, name, training, eps):
    """Return W / sigma for the layer whose keys are name.weight/u0/sv0.
    One power iteration from the stored u; u0 and sv0 written only when
    training (layers.py:106-107, 161-164)."""
    w = sd[name + ".weight"]
    u = sd[name + ".u0"]
    w2 = w.reshape(w.shape[0], -1)
    with torch.no_grad():
        v = F.normalize(u @ w2, eps=eps)
        u_new = F.normalize(v @ w2.t(), eps=eps)
    sigma = ((v @ w2.t()) @ u_new.t()).squeeze()
    if training:
        with torch.no_grad():
            u.copy_(u_new)
            sd[name + ".sv0"].fill_(float(sigma))
    return w / sigma


class _Weights:
    """All spectrally-normalised weights of one forward call, computed once
    per call (multi-event rule) in first-use order."""

    def __init__(self, sd, training, eps):
        self.sd, self.training, self.eps, self.cache = sd, training, eps, {}

    def __call__(self, name):
        if name not in self.cache:
            if name + ".u0" in self.sd:
                self.cache[name] = sn_weight(self.sd, name, self.training, self.eps)
            else:
                self.cache[name] = self.sd[name + ".weight"]
        return self.cache[name]

    def bias(self, name):
        return self.sd.get(name + ".bias")


def _conv(w, x, name, k):
    return F.conv2d(x, w(name), w.bias(name), padding=k // 2)  # layers.py:198-206


def _linear(w, x, name):
    return F.linear(x, w(name), w.bias(name))  # layers.py:224


# --------------------------------------------------------------------------
# batch norm (layers.py:656-689 ccbn, layers.py:728-742 bn), per event
# --------------------------------------------------------------------------
def _batch_norm(sd, name, x, training, eps, weight=None, bias=None):
    outs = []
    for e in range(x.shape[0] // IMGS):  # running stats see the events in order
        outs.append(F.batch_norm(x[e * IMGS:(e + 1) * IMGS], sd[name + ".stored_mean"],
                                 sd[name + ".stored_var"], weight, bias, training, 0.1, eps))
    return torch.cat(outs, 0) if len(outs) > 1 else outs[0]


def ccbn_forward(sd, w, name, x, y, training, bn_eps):
    gain = (1 + _linear(w, y, name + ".gain")).view(y.shape[0], -1, 1, 1)
    bias = _linear(w, y, name + ".bias").view(y.shape[0], -1, 1, 1)
    return _batch_norm(sd, name, x, training, bn_eps) * gain + bias


# --------------------------------------------------------------------------
# RRM (RRM.py:10-16, 44-63, 98-109, 120-125)
# --------------------------------------------------------------------------
def rrm_forward(sd, w, name, x, heads):
    """x: (B, 40, E).  One pre-LN encoder block + final LayerNorm."""
    p = name + ".layers.0."
    b, s, e = x.shape
    d = e // heads

    def ln(t, key):
        return F.layer_norm(t, (e,), sd[key + ".weight"], sd[key + ".bias"])

    h = ln(x, p + "norm1")
    qkv = _linear(w, h, p + "self_attn.qkv_proj").reshape(b, s, heads, 3 * d).permute(0, 2, 1, 3)
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]  # per-head interleave, RRM.py:49-53
    att = F.softmax(q @ k.transpose(-2, -1) / math.sqrt(d), dim=-1)
    val = (att @ v).permute(0, 2, 1, 3).reshape(b, s, e)
    x = x + _linear(w, val, p + "self_attn.o_proj")
    h = ln(x, p + "norm2")
    h = _linear(w, F.relu(_linear(w, h, p + "linear_net.0")), p + "linear_net.3")
    x = x + h
    return ln(x, name + ".norm")


# --------------------------------------------------------------------------
# Generator (model.py:54-71 GBlock.forward, model.py:454-487 Generator.forward)
# --------------------------------------------------------------------------
def generator_forward(sd, cfg, z, y, rdof, training=True):
    """z (40E,dim_z), y (40E,) int64, rdof (40E,rdof_dim) -> (40E,1,res,res*H_base)."""
    w = _Weights(sd, training, cfg["SN_eps"])
    bn_eps = cfg.get("BN_eps", 1e-5)
    n = z.shape[0]
    emb = F.embedding(y, sd["shared.weight"])  # plain nn.Embedding, model.py:263,295-299
    c = _linear(w, torch.cat([emb, rdof], 1), "linear_f")
    c = rrm_forward(sd, w, "RR_G", c.view(n // IMGS, IMGS, -1), cfg["n_head_G"]).reshape(n, -1)
    c = torch.cat([c, z], 1)  # hier: the same 256-vector conditions every ccbn, model.py:471-473
    hb = cfg["H_base"]
    bw = cfg.get("bottom_width", 4)
    h = _linear(w, c, "linear").view(n, -1, bw, bw * hb)
    cin, cout = g_channels(cfg)
    for s in range(len(cin)):
        for g in range(cfg["G_depth"]):
            p = "blocks.%d.0." % (s * cfg["G_depth"] + g)
            last = g == cfg["G_depth"] - 1
            co = cout[s] if last else cin[s]
            x = h
            h = _conv(w, F.relu(ccbn_forward(sd, w, p + "bn1", x, c, training, bn_eps)), p + "conv1", 1)
            h = F.relu(ccbn_forward(sd, w, p + "bn2", h, c, training, bn_eps))
            if co != cin[s]:
                x = x[:, :co]  # channel drop, model.py:60-61
            if last:  # nearest x2 on both branches, model.py:63-65
                h = F.interpolate(h, scale_factor=2)
                x = F.interpolate(x, scale_factor=2)
            h = _conv(w, h, p + "conv2", 3)
            h = _conv(w, F.relu(ccbn_forward(sd, w, p + "bn3", h, c, training, bn_eps)), p + "conv3", 3)
            h = _conv(w, F.relu(ccbn_forward(sd, w, p + "bn4", h, c, training, bn_eps)), p + "conv4", 1)
            h = h + x
    h = _batch_norm(sd, "output_layer.0", h, training, 1e-5,  # layers.bn default eps, model.py:380-384
                    sd["output_layer.0.gain"], sd["output_layer.0.bias"])
    return torch.tanh(_conv(w, F.relu(h), "output_layer.2", 3))


# --------------------------------------------------------------------------
# Discriminator (model.py:541-557 DBlock, layers.py:283-300 Attention,
# model.py:902-937 Discriminator.forward, Contra head)
# --------------------------------------------------------------------------
def attention_forward(sd, w, p, x):
    n, ch, hh, ww = x.shape
    theta = _conv(w, x, p + "theta", 1).view(n, ch // 8, hh * ww)
    phi = F.max_pool2d(_conv(w, x, p + "phi", 1), 2).view(n, ch // 8, hh * ww // 4)
    g = F.max_pool2d(_conv(w, x, p + "g", 1), 2).view(n, ch // 2, hh * ww // 4)
    beta = F.softmax(theta.transpose(1, 2) @ phi, -1)  # no 1/sqrt(d), layers.py:293
    o = _conv(w, (g @ beta.transpose(1, 2)).view(n, ch // 2, hh, ww), p + "o", 1)
    return sd[p + "gamma"] * o + x


def discriminator_forward(sd, cfg, x, y, training=True):
    """x (40E,1,H,W), y (40E,) -> (proxy (40E,D), embed (40E,D), out (40E,))."""
    w = _Weights(sd, training, cfg["SN_eps"])
    n = x.shape[0]
    cin, cout = d_channels(cfg)
    attn_after = d_attention_stages(cfg)
    h = _conv(w, x, "input_conv", 3)
    for s in range(len(cin)):
        for d in range(cfg["D_depth"]):
            p = "blocks.%d.%d." % (s, d)
            down = d == 0  # every listed stage downsamples in its first block
            x0 = h
            t = F.relu(h) if (s > 0 or d > 0) else h  # model.py:745 preactivation
            t = _conv(w, t, p + "conv1", 1)
            t = _conv(w, F.relu(t), p + "conv2", 3)
            t = F.relu(_conv(w, F.relu(t), p + "conv3", 3))
            if down:
                t = F.avg_pool2d(t, 2)
                x0 = F.avg_pool2d(x0, 2)
            t = _conv(w, t, p + "conv4", 1)
            if p + "conv_sc.weight" in sd:  # learnable shortcut = concat, model.py:534-539
                x0 = torch.cat([x0, _conv(w, x0, p + "conv_sc", 1)], 1)
            h = t + x0
        if s in attn_after:
            h = attention_forward(sd, w, "blocks.%d.%d." % (s, cfg["D_depth"]), h)
    h = torch.sum(F.relu(h), [2, 3])
    out = _linear(w, h, "linear0").squeeze(-1)
    proxy = F.embedding(y, w("embed"))
    h = rrm_forward(sd, w, "RR_D", h.view(n // IMGS, IMGS, -1), cfg.get("n_head_D", 4)).reshape(n, -1)
    emb = _linear(w, h, "linear1")
    emb = F.layer_norm(emb, (emb.shape[1],), sd["norm.weight"], sd["norm.bias"])
    return F.normalize(proxy, dim=1), F.normalize(emb, dim=1), out


# --------------------------------------------------------------------------
# DiffAugment, policy "color,translation,cutout", one channel
# (diff_aug.py:23-102; draws made by the caller in the reference's order)
# --------------------------------------------------------------------------
def diffaug_draws(n, hh, ww, device="cpu", generator=None):
    """The seven RNG draws of one DiffAugment call, in the reference's order
    (diff_aug.py:25,33,41,50-55,74-85)."""
    kw = dict(device=device, generator=generator)
    r = [torch.rand(n, 1, 1, 1, **kw) for _ in range(3)]
    sx, sy = int(hh * 0.125 + 0.5), int(ww * 0.125 + 0.5)
    tx = torch.randint(-sx, sx + 1, [n, 1, 1], **kw)
    ty = torch.randint(-sy, sy + 1, [n, 1, 1], **kw)
    ch, cw = int(hh * 0.5 + 0.5), int(ww * 0.5 + 0.5)
    ox = torch.randint(0, hh + (1 - ch % 2), [n, 1, 1], **kw)
    oy = torch.randint(0, ww + (1 - cw % 2), [n, 1, 1], **kw)
    return dict(brightness=r[0], saturation=r[1], contrast=r[2], tx=tx, ty=ty, ox=ox, oy=oy)


def diffaugment(x, d):
    n, c, hh, ww = x.shape
    assert c == 1
    x = x + (d["brightness"] - 0.5)
    # saturation: (x - mean_c x) * 2r + mean_c x is the identity for one channel
    m = x.mean(dim=[1, 2, 3], keepdim=True)
    x = (x - m) * (d["contrast"] + 0.5) + m
    ii = torch.arange(hh).view(1, hh, 1) + d["tx"]  # source row index
    jj = torch.arange(ww).view(1, 1, ww) + d["ty"]
    inside = ((ii >= 0) & (ii < hh) & (jj >= 0) & (jj < ww)).unsqueeze(1)
    src = x[torch.arange(n).view(n, 1, 1), 0, ii.clamp(0, hh - 1), jj.clamp(0, ww - 1)].unsqueeze(1)
    x = torch.where(inside, src, torch.zeros_like(src))
    ch, cw = int(hh * 0.5 + 0.5), int(ww * 0.5 + 0.5)
    r0 = d["ox"] - ch // 2
    c0 = d["oy"] - cw // 2
    rows = torch.arange(hh).view(1, hh, 1)
    cols = torch.arange(ww).view(1, 1, ww)
    # rows covered by clamp(r0 + [0,ch), 0, hh-1): a box hanging over the border still hits the border row
    rhit = (rows >= r0.clamp(0, hh - 1)) & (rows <= (r0 + ch - 1).clamp(0, hh - 1))
    chit = (cols >= c0.clamp(0, ww - 1)) & (cols <= (c0 + cw - 1).clamp(0, ww - 1))
    return x * (~(rhit & chit)).unsqueeze(1).to(x.dtype)


# --------------------------------------------------------------------------
# losses (loss.py:8-9, 14-27, 30-38, 79-132), per event then averaged
# --------------------------------------------------------------------------
def hinge_dis(d_fake, d_real):
    return torch.mean(F.relu(1.0 - d_real)), torch.mean(F.relu(1.0 + d_fake))


def hinge_gen(d_fake):
    return -torch.mean(d_fake)


def _per_event(fn, *ts):
    e = ts[0].shape[0] // IMGS
    return sum(fn(*[t[i * IMGS:(i + 1) * IMGS] for t in ts]) for i in range(e)) / e


def contrastive(embed, proxy, temperature=1.0, margin=0.0):
    def one(em, pr):
        n = em.shape[0]
        sim = F.cosine_similarity(em.unsqueeze(1), em.unsqueeze(0), dim=-1)
        off = ~torch.eye(n, dtype=torch.bool)
        zone = torch.exp((sim[off].view(n, n - 1) - margin) / temperature)
        pos = torch.exp((F.cosine_similarity(em, pr, dim=-1) - margin) / temperature)
        return -torch.log(temperature * (pos / (pos + zone.sum(1)))).mean()
    return _per_event(one, embed, proxy)


def iea(k_f, k_r):
    def one(f, r):
        with torch.no_grad():
            pr = F.softmax(r @ r.t(), dim=-1)
        lq = F.log_softmax(f @ f.t(), dim=-1)
        return F.kl_div(lq, pr, reduction="batchmean")
    return _per_event(one, k_f, k_r)


def uniformity(x, t=2):
    return _per_event(lambda a: torch.pdist(a, p=2).pow(2).mul(-t).exp().mean().log(), x)


# --------------------------------------------------------------------------
# one G+D step (train_fns.py:23-205, Contra branch of the shipped config)
# --------------------------------------------------------------------------
def ortho_grad(params, strength, skip=()):
    """utils/__init__.py:843-859 modified orthogonal regularisation gradient."""
    with torch.no_grad():
        for name, p in params.items():
            if p.dim() < 2 or name in skip or p.grad is None:
                continue
            m = p.view(p.shape[0], -1)
            gram = m @ m.t()
            gram.fill_diagonal_(0.0)
            p.grad += strength * (2 * gram @ m).view_as(p)


def set_requires_grad(sd, names, flag):
    for k in names:
        sd[k].requires_grad_(flag)


def param_names(sd):
    return [k for k in sd if not (k.endswith(".u0") or k.endswith(".sv0")
                                  or k.endswith("stored_mean") or k.endswith("stored_var"))]


def train_step(sd_g, sd_d, cfg, x, y, noise, opt_g=None, opt_d=None):
    """One step of train_fns.train.  noise = dict(z_d, rdof_d, aug_d, z_g, rdof_g,
    aug_g) holds every random draw of the step.  Returns the losses; leaves
    .grad on the parameters (so tests can read them) and, when optimisers
    are given, applies them like the reference (D always; G only when
    clip_norm is not None, train_fns.py:190-192)."""
    pg, pd = param_names(sd_g), param_names(sd_d)
    for k in pg:
        sd_g[k].grad = None
    for k in pd:
        sd_d[k].grad = None
    set_requires_grad(sd_d, pd, True)
    set_requires_grad(sd_g, pg, False)
    # ---- D step: fakes generated under no_grad but in training mode (model.py:973-978)
    with torch.no_grad():
        g_z = generator_forward(sd_g, cfg, noise["z_d"], y, noise["rdof_d"], True)
        g_z = diffaugment(g_z, noise["aug_d"])
    _, _, d_fake = discriminator_forward(sd_d, cfg, g_z, y, True)
    proxy_r, embed_r, d_real = discriminator_forward(sd_d, cfg, x, y, True)  # real x not augmented
    l_real, l_fake = hinge_dis(d_fake, d_real)
    unif_d = uniformity(embed_r)
    d_loss = l_real + l_fake + cfg["contra_lambda"] * contrastive(embed_r, proxy_r) \
        + cfg["unif_lambda"] * unif_d
    d_loss.backward()
    if cfg.get("D_ortho", 0.0) > 0:
        ortho_grad({k: sd_d[k] for k in pd}, cfg["D_ortho"])
    if cfg.get("clip_norm") is not None:
        torch.nn.utils.clip_grad_norm_([sd_d[k] for k in pd], cfg["clip_norm"])
    if opt_d is not None:
        opt_d.step()
    # ---- G step
    set_requires_grad(sd_d, pd, False)
    set_requires_grad(sd_g, pg, True)
    g_z = generator_forward(sd_g, cfg, noise["z_g"], y, noise["rdof_g"], True)
    g_z = diffaugment(g_z, noise["aug_g"])
    proxy_f, embed_f, d_fake = discriminator_forward(sd_d, cfg, g_z, y, True)
    iea_l = iea(embed_f, embed_r.detach())
    g_loss = hinge_gen(d_fake) + cfg["contra_lambda"] * contrastive(embed_f, proxy_f) \
        + cfg["IEA_lambda"] * iea_l + cfg["unif_lambda"] * uniformity(embed_f)
    g_loss.backward()
    if cfg.get("G_ortho", 0.0) > 0:
        ortho_grad({k: sd_g[k] for k in pg}, cfg["G_ortho"], skip=("shared.weight",))
    if cfg.get("clip_norm") is not None:
        torch.nn.utils.clip_grad_norm_([sd_g[k] for k in pg], cfg["clip_norm"])
        if opt_g is not None:
            opt_g.step()
    return dict(G_loss=float(g_loss.detach()), D_loss_real=float(l_real.detach()), D_loss_fake=float(l_fake.detach()),
                unif_loss_d=float(unif_d.detach()), iea_loss=float(iea_l.detach()))


def generate_postprocess(imgs):
    """model.py:1139-1147: 7-ADU cut, [-1,1]->[0,1], 256^x - 1, clamp, crop 3 rows."""
    imgs = F.threshold(imgs, -0.26, -1)
    imgs = torch.pow(256, imgs * 0.5 + 0.5).add(-1).clamp(0, 255)
    return imgs[:, 0, 3:-3, :]


def preprocess_events(images_u8, draws=None, scale=4e-3, pad=3):
    """utils/dataloader.py:69-77 on a decoded uint8 batch (N,H,W): zero-pad `pad` rows top and bottom, ToTensor,
    fn_lognorm255 (utils/norm.py:8-18), + scale*U[0,1) (utils/noise.py:32-35), Normalize(0.5, 0.5)."""
    n, h, w = images_u8.shape
    x = torch.zeros(n, 1, h + 2 * pad, w)
    x[:, 0, pad:pad + h] = torch.log(255 * (images_u8.float() / 255) + 1) / math.log(256)
    if draws is not None:
        x = x + scale * draws
    return (x - 0.5) / 0.5
