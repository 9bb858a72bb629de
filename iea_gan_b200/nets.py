"""IEA-GAN Generator / Discriminator (BigGAN-deep bottleneck blocks, RRM, Contra head).

Host-side mirror of the reference's model.py operator interface: same class
names, constructor keywords (incl. the **kwargs sink), attribute names,
state-dict keys, parameter initialisation order (so equal seeds give equal
weights) and forward signatures.  The forward/backward arithmetic is executed
by the fused sm_100a kernels through `engine`; there is no CPU path.

Reference interface: model.py:16-71 (GBlock), :74-136 (G_arch), :139-487
(Generator), :490-557 (DBlock), :561-621 (D_arch), :624-944 (Discriminator),
:949-1121 (G_D), :1124-1148 (Model, generate).
"""
import functools

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.optim as optim
from torch.nn import init

from . import sn_layers as layers
from . import relational as RRM
from . import engine as E
from .augment import DiffAugment
from .optim import FusedAdam

# channel multipliers per output resolution: (in, out)
_G_TABLE = {512: ([16, 16, 8, 8, 4, 2, 1], [16, 8, 8, 4, 2, 1, 1]),
            256: ([16, 16, 8, 8, 4, 2], [16, 8, 8, 4, 2, 1]),
            128: ([16, 16, 8, 4, 2], [16, 8, 4, 2, 1]),
            64: ([16, 16, 8, 4], [16, 8, 4, 2]),
            32: ([4, 4, 4], [4, 4, 4])}
_D_TABLE = {512: ([1, 1, 2, 4, 8, 8, 16], [1, 2, 4, 8, 8, 16, 16]),
            256: ([1, 2, 4, 8, 8, 16], [2, 4, 8, 8, 16, 16]),
            128: ([1, 2, 4, 8, 16], [2, 4, 8, 16, 16]),
            64: ([1, 2, 4, 8], [2, 4, 8, 16])}


def _attn_set(attention):
    return {int(t) for t in str(attention).split("_")}


def G_arch(ch=64, attention="64", ksize="333333", dilation="111111"):
    want, arch = _attn_set(attention), {}
    for res, (mi, mo) in _G_TABLE.items():
        stages = [res >> (len(mi) - 1 - i) for i in range(len(mi))]
        arch[res] = {"in_channels": [ch * m for m in mi], "out_channels": [ch * m for m in mo],
                     "upsample": [True] * len(mi), "resolution": stages,
                     "attention": {r: r in want for r in stages}}
    return arch


def D_arch(ch=64, attention="64", ksize="333333", dilation="111111"):
    want, arch = _attn_set(attention), {}
    for res, (mi, mo) in _D_TABLE.items():
        stages = [max(res >> (i + 1), 4) for i in range(len(mi))]
        arch[res] = {"in_channels": [ch * m for m in mi], "out_channels": [ch * m for m in mo],
                     "downsample": [True] * len(mi), "resolution": stages,
                     "attention": {r: r in want for r in set(stages)}}
    return arch


def _make_activation(name):
    if name == "inplace_relu":
        return nn.ReLU(inplace=True)
    if name == "relu":
        return nn.ReLU(inplace=False)
    raise NotImplementedError("activation function %s not built (fused prologues implement ReLU)" % name)


def _ortho_init(net, style):
    """Same module walk and RNG consumption as model.py:430-452 / :878-900."""
    net.param_count = 0
    for m in net.modules():
        if isinstance(m, (nn.Conv2d, nn.Linear, nn.Embedding)):
            if style == "ortho":
                init.orthogonal_(m.weight)
            elif style == "N02":
                init.normal_(m.weight, 0, 0.02)
            elif style in ("glorot", "xavier"):
                init.xavier_uniform_(m.weight)
            else:
                print("Init style not recognized...")
            net.param_count += sum(p.data.nelement() for p in m.parameters())


def _make_optim(net, lr, b1, b2, eps, sched_version, kwargs):
    net.lr, net.B1, net.B2, net.adam_eps = lr, b1, b2, eps
    # torch.optim.Adam's update (model.py:410-416, 858-864) as one multi-tensor kernel; same state_dict layout
    net.optim = FusedAdam(params=net.parameters(), lr=lr, betas=(b1, b2), weight_decay=0, eps=eps)
    if sched_version == "CosAnnealLR":
        net.lr_sched = optim.lr_scheduler.CosineAnnealingLR(net.optim, T_max=kwargs["num_epochs"],
                                                            eta_min=lr / 4, last_epoch=-1)
    elif sched_version == "CosAnnealWarmRes":
        net.lr_sched = optim.lr_scheduler.CosineAnnealingWarmRestarts(net.optim, T_0=10, T_mult=2, eta_min=lr / 4)
    else:
        net.lr_sched = None


class GBlock(nn.Module):
    """ccbn-ReLU-1x1, ccbn-ReLU-[up]-3x3, ccbn-ReLU-3x3, ccbn-ReLU-1x1, + channel-dropped
    (and upsampled) skip (model.py:54-71)."""

    def __init__(self, in_channels, out_channels, which_conv=layers.SNConv2d, which_bn=layers.bn,
                 activation=None, upsample=None, channel_ratio=4):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.hidden_channels = in_channels // channel_ratio
        self.which_conv, self.which_bn, self.activation, self.upsample = which_conv, which_bn, activation, upsample
        hid = self.hidden_channels
        self.conv1 = which_conv(in_channels, hid, kernel_size=1, padding=0)
        self.conv2 = which_conv(hid, hid)
        self.conv3 = which_conv(hid, hid)
        self.conv4 = which_conv(hid, out_channels, kernel_size=1, padding=0)
        self.bn1 = which_bn(in_channels)
        self.bn2 = which_bn(hid)
        self.bn3 = which_bn(hid)
        self.bn4 = which_bn(hid)

    def forward(self, x, y):
        return E.module_gblock(self, x, y)


class Generator(nn.Module):
    def __init__(self, G_ch=64, G_depth=2, dim_z=128, bottom_width=4, resolution=256, G_kernel_size=3,
                 G_attn="64", n_classes=40, H_base=1, num_G_SVs=1, num_G_SV_itrs=1, attn_type="sa",
                 G_shared=True, shared_dim=128, rdof_dim=4, hier=True, cross_replica=False, mybn=False,
                 G_activation="relu", G_lr=5e-5, G_B1=0.0, G_B2=0.999, adam_eps=1e-8, BN_eps=1e-5,
                 SN_eps=1e-12, G_init="ortho", G_mixed_precision=False, G_fp16=False, skip_init=False,
                 no_optim=False, sched_version="default", RRM_prx_G=True, prior_embed=False, n_head_G=2,
                 G_param="SN", norm_style="bn", device="cuda", **kwargs):
        super().__init__()
        self.ch, self.G_depth, self.dim_z, self.bottom_width = G_ch, G_depth, dim_z, bottom_width
        self.H_base, self.resolution, self.kernel_size, self.attention = H_base, resolution, G_kernel_size, G_attn
        self.n_classes, self.G_shared = n_classes, G_shared
        self.shared_dim = shared_dim if shared_dim > 0 else dim_z
        self.hier, self.cross_replica, self.mybn = hier, cross_replica, mybn
        self.activation = _make_activation(G_activation)
        self.init, self.G_param, self.norm_style = G_init, G_param, norm_style
        self.BN_eps, self.SN_eps, self.fp16 = BN_eps, SN_eps, G_fp16
        self.arch = G_arch(self.ch, self.attention)[resolution]
        self.RRM_prx_G, self.n_head_G, self.prior_embed = RRM_prx_G, n_head_G, prior_embed
        self.rdof_dim, self.device = rdof_dim, device
        if G_param != "SN" or not G_shared or not hier or prior_embed or not RRM_prx_G:
            raise NotImplementedError("built: G_param='SN', G_shared, hier, RRM_prx_G, no prior_embed "
                                      "(config.json:11,17,21,103,110)")
        if attn_type != "sa":
            raise NotImplementedError("built: attn_type='sa'")
        sn = dict(num_svs=num_G_SVs, num_itrs=num_G_SV_itrs, eps=SN_eps)
        self.which_conv = functools.partial(layers.SNConv2d, kernel_size=3, padding=1, **sn)
        self.which_linear = functools.partial(layers.SNLinear, **sn)
        self.which_embedding = nn.Embedding  # G's class embedding is never spectrally normalised
        self.which_bn = functools.partial(
            layers.ccbn, which_linear=functools.partial(self.which_linear, bias=False),
            cross_replica=cross_replica, mybn=mybn, input_size=self.shared_dim + dim_z,
            norm_style=norm_style, eps=BN_eps)

        self.shared = self.which_embedding(n_classes, self.shared_dim)
        self.linear_f = self.which_linear(self.shared_dim + rdof_dim, 128)
        self.RR_G = RRM.RelationalReasoning(num_layers=1, input_dim=128, dim_feedforward=128,
                                            which_linear=nn.Linear, num_heads=n_head_G, dropout=0.0,
                                            hidden_dim=128)
        cin, cout = self.arch["in_channels"], self.arch["out_channels"]
        self.linear = self.which_linear(dim_z + self.shared_dim, cin[0] * (bottom_width ** 2 * H_base))
        blocks = []
        for s in range(len(cout)):
            for g in range(G_depth):
                last = g == G_depth - 1
                up = functools.partial(F.interpolate, scale_factor=2) if (self.arch["upsample"][s] and last) else None
                blocks.append([GBlock(cin[s], cout[s] if g else cin[s], which_conv=self.which_conv,
                                      which_bn=self.which_bn, activation=self.activation, upsample=up)])
            if self.arch["attention"][self.arch["resolution"][s]]:
                print("Adding attention layer in G at resolution %d" % self.arch["resolution"][s])
                blocks[-1].append(layers.Attention(cout[s], self.which_conv))
        self.blocks = nn.ModuleList([nn.ModuleList(b) for b in blocks])
        self.output_layer = nn.Sequential(layers.bn(cout[-1], cross_replica=cross_replica, mybn=mybn),
                                          self.activation, self.which_conv(cout[-1], 1))
        if not skip_init:
            self.init_weights()
        if no_optim:
            return
        _make_optim(self, G_lr, G_B1, G_B2, adam_eps, sched_version, kwargs)

    def init_weights(self):
        _ortho_init(self, self.init)
        print("Param count for Gs initialized parameters: %d" % self.param_count)

    def forward(self, z, y):
        """z (40E, dim_z) float, y (40E,) int64 -> (40E, 1, res, res*H_base) in [-1, 1].
        Draws rdof = randn(40E, rdof_dim) from the device generator first (model.py:466)."""
        return E.generator_forward(self, z, y)


class DBlock(nn.Module):
    """[ReLU]-1x1, ReLU-3x3, ReLU-3x3, ReLU-[avgpool]-1x1 + concat shortcut (model.py:541-557)."""

    def __init__(self, in_channels, out_channels, which_conv=layers.SNConv2d, wide=True, preactivation=True,
                 activation=None, downsample=None, channel_ratio=4):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.hidden_channels = out_channels // channel_ratio
        self.which_conv, self.preactivation = which_conv, preactivation
        self.activation, self.downsample = activation, downsample
        hid = self.hidden_channels
        self.conv1 = which_conv(in_channels, hid, kernel_size=1, padding=0)
        self.conv2 = which_conv(hid, hid)
        self.conv3 = which_conv(hid, hid)
        self.conv4 = which_conv(hid, out_channels, kernel_size=1, padding=0)
        self.learnable_sc = in_channels != out_channels
        if self.learnable_sc:
            self.conv_sc = which_conv(in_channels, out_channels - in_channels, kernel_size=1, padding=0)

    def forward(self, x):
        return E.module_dblock(self, x)


class Discriminator(nn.Module):
    def __init__(self, D_ch=64, D_wide=True, D_depth=2, resolution=256, D_kernel_size=3, D_attn="64",
                 n_classes=40, attn_type="sa", num_D_SVs=1, num_D_SV_itrs=1, D_activation="relu",
                 conditional_strategy="Proj", D_lr=2e-4, D_B1=0.0, D_B2=0.999, adam_eps=1e-8, SN_eps=1e-12,
                 output_dim=1, D_init="ortho", D_mixed_precision=False, D_fp16=False, sched_version="default",
                 skip_init=False, D_param="SN", hypersphere_dim=512, nonlinear_embed=False,
                 normalize_embed=True, prior_embed=False, RRM_prx_D=False, RRM_embed=False, n_head_D=4,
                 **kwargs):
        super().__init__()
        self.ch, self.D_wide, self.D_depth, self.resolution = D_ch, D_wide, D_depth, resolution
        self.kernel_size, self.attention, self.n_classes = D_kernel_size, D_attn, n_classes
        self.activation = _make_activation(D_activation)
        self.init, self.D_param, self.SN_eps, self.fp16 = D_init, D_param, SN_eps, D_fp16
        self.RRM_prx_D, self.RRM_embed, self.prior_embed = RRM_prx_D, RRM_embed, prior_embed
        self.conditional_strategy, self.nonlinear_embed = conditional_strategy, nonlinear_embed
        self.normalize_embed, self.n_head_D = normalize_embed, n_head_D
        self.arch = D_arch(self.ch, self.attention)[resolution]
        if (D_param != "SN" or conditional_strategy != "Contra" or not RRM_embed or RRM_prx_D or prior_embed
                or nonlinear_embed or not normalize_embed or attn_type != "sa" or output_dim != 1):
            raise NotImplementedError("built: the shipped Contra head (SN, RRM_embed, normalize_embed; "
                                      "config.json:85-106)")
        sn = dict(num_svs=num_D_SVs, num_itrs=num_D_SV_itrs, eps=SN_eps)
        self.which_conv = functools.partial(layers.SNConv2d, kernel_size=3, padding=1, **sn)
        self.which_linear = functools.partial(layers.SNLinear, **sn)
        self.which_embedding = functools.partial(layers.SNEmbedding, **sn)

        cin, cout = self.arch["in_channels"], self.arch["out_channels"]
        self.input_conv = self.which_conv(1, cin[0])
        blocks = []
        for s in range(len(cout)):
            stage = [DBlock(cin[s] if d == 0 else cout[s], cout[s], which_conv=self.which_conv, wide=D_wide,
                            activation=self.activation, preactivation=(s > 0 or d > 0),
                            downsample=nn.AvgPool2d(2) if (self.arch["downsample"][s] and d == 0) else None)
                     for d in range(D_depth)]
            if self.arch["attention"][self.arch["resolution"][s]]:
                print("Adding attention layer in D at resolution %d" % self.arch["resolution"][s])
                stage.append(layers.Attention(cout[s], self.which_conv))
            blocks.append(stage)
        self.blocks = nn.ModuleList([nn.ModuleList(b) for b in blocks])
        self.linear0 = self.which_linear(cout[-1], output_dim)
        # the reference hard-codes a 512-wide RRM here (model.py:788-797), so D_ch must be 32
        self.RR_D = RRM.RelationalReasoning(num_layers=1, input_dim=cout[-1], dim_feedforward=512,
                                            num_heads=n_head_D, dropout=0.0, hidden_dim=512,
                                            which_linear=self.which_linear)
        self.norm = nn.LayerNorm(hypersphere_dim)
        self.linear1 = self.which_linear(cout[-1], hypersphere_dim)
        self.embed = self.which_embedding(n_classes, hypersphere_dim)
        if not skip_init:
            self.init_weights()
        _make_optim(self, D_lr, D_B1, D_B2, adam_eps, sched_version, kwargs)

    def init_weights(self):
        _ortho_init(self, self.init)
        print("Param count for Ds initialized parameters: %d" % self.param_count)

    def forward(self, x, y=None):
        """x (40E,1,H,W), y (40E,) -> (cls_proxy (40E,D), cls_embed (40E,D), out (40E,)), both L2-normalised."""
        return E.discriminator_forward(self, x, y)


class G_D(nn.Module):
    """G then DiffAugment (fakes only) then D on fake [and real] (model.py:956-1121, the
    split_D / contra return variants train_fns.py uses)."""

    def __init__(self, G, D):
        super().__init__()
        self.G, self.D = G, D

    def forward(self, z, gy, x=None, dy=None, x_aug=None, contra=True, train_G=False, return_G_z=False,
                split_D=False, diff_aug=True, pixel_reg=False):
        if not (contra and split_D) or x_aug is not None or pixel_reg or return_G_z:
            raise NotImplementedError("built: contra=True, split_D=True, no x_aug / pixel_reg "
                                      "(the call shapes of train_fns.py:105-110,161-164)")
        with torch.set_grad_enabled(train_G):
            G_z = self.G(z, gy)
            if diff_aug:
                G_z = DiffAugment(G_z, policy="color,translation,cutout")
        fake = self.D(G_z, gy)
        if train_G:
            return fake
        return fake + self.D(x, dy)


class Model(Generator):
    def __init__(self, config: dict):
        assert isinstance(config, dict), "Expected configuration dictionary"
        super().__init__(**config)


def generate(model):
    """Sample one event and convert to ADU counts (model.py:1130-1148): 7-ADU cut,
    256^x - 1, clamp, crop rows 3:-3.  The post-process runs fused on the GPU; only the
    final (40, 250, W) tensor crosses to the host."""
    device = next(model.parameters()).device
    with torch.no_grad():
        latents = torch.randn(40, model.dim_z, device=device)
        labels = torch.arange(40, dtype=torch.long, device=device)
        return E.adu_postprocess(model(latents, labels)).cpu()
