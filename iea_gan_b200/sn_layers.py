"""Spectrally-normalised layers, conditional batch-norm and BigGAN self-attention.

Host-side mirror of the reference's `layers.py` operator interface (same class
names, constructor arguments, attribute / state-dict names and RNG consumption)
whose arithmetic runs in the sm_100a kernels of libiea_sm100.so.  There is no
CPU path: calling forward without the CUDA extension raises.

Reference interface: layers.py:89-111 (power_iteration), :121-165 (SN),
:169-206 (SNConv2d), :210-224 (SNLinear), :230-259 (SNEmbedding),
:262-300 (Attention), :622-694 (ccbn), :698-742 (bn).
"""
import torch
import torch.nn as nn

from . import engine as E


class identity(nn.Module):
    def forward(self, tensor):
        return tensor


def power_iteration(W, u_, update=True, eps=1e-12):
    """One power-iteration step for the leading singular value (layers.py:89-111).
    Returns ([sigma], [u'], [v]); u_[0] is overwritten when `update`.  Runs the
    grouped spectral-norm kernel on a single layer; sigma carries no autograd
    history here (the layers below differentiate through it in their own
    backward kernels)."""
    assert len(u_) == 1, "only num_svs == 1 is built (shipped config: num_G_SVs = num_D_SVs = 1)"
    sig, un, v = E.power_iteration_single(W, u_[0], update, eps)
    return [sig], [un], [v]


class SN(object):
    """Mix-in holding the power-iteration state u0 (1,out) and sv0 (1,) (layers.py:121-148)."""

    def __init__(self, num_svs, num_itrs, num_outputs, transpose=False, eps=1e-12):
        if num_svs != 1 or num_itrs != 1 or transpose:
            raise NotImplementedError("iea_gan_b200 builds num_svs=1, num_itrs=1, transpose=False "
                                      "(the only setting the reference's config uses)")
        self.num_itrs, self.num_svs, self.transpose, self.eps = num_itrs, num_svs, transpose, eps
        for i in range(num_svs):
            self.register_buffer("u%d" % i, torch.randn(1, num_outputs))
            self.register_buffer("sv%d" % i, torch.ones(1))

    @property
    def u(self):
        return [getattr(self, "u%d" % i) for i in range(self.num_svs)]

    @property
    def sv(self):
        return [getattr(self, "sv%d" % i) for i in range(self.num_svs)]

    def W_(self):
        """weight / sigma after one power iteration from the stored u (layers.py:151-165).
        Differentiable w.r.t. weight (rank-1 corrected backward kernel)."""
        return E.sn_weight_standalone(self)


class SNConv2d(nn.Conv2d, SN):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, num_svs=1, num_itrs=1, eps=1e-12):
        nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        SN.__init__(self, num_svs, num_itrs, out_channels, eps=eps)
        k = self.kernel_size
        if not (k[0] == k[1] and k[0] in (1, 3) and self.stride == (1, 1) and self.dilation == (1, 1)
                and self.groups == 1 and self.padding == (k[0] // 2, k[0] // 2)):
            raise NotImplementedError("built: 1x1 / 3x3, stride 1, 'same' padding (all the reference nets use)")

    def forward(self, x):
        return E.module_conv(self, x)


class SNLinear(nn.Linear, SN):
    def __init__(self, in_features, out_features, bias=True, num_svs=1, num_itrs=1, eps=1e-12):
        nn.Linear.__init__(self, in_features, out_features, bias)
        SN.__init__(self, num_svs, num_itrs, out_features, eps=eps)

    def forward(self, x):
        return E.module_linear(self, x)


class SNEmbedding(nn.Embedding, SN):
    """u has num_embeddings entries (layers.py:256)."""

    def __init__(self, num_embeddings, embedding_dim, padding_idx=None, max_norm=None, norm_type=2,
                 scale_grad_by_freq=False, sparse=False, _weight=None, num_svs=1, num_itrs=1, eps=1e-12):
        nn.Embedding.__init__(self, num_embeddings, embedding_dim, padding_idx, max_norm, norm_type,
                              scale_grad_by_freq, sparse, _weight)
        SN.__init__(self, num_svs, num_itrs, num_embeddings, eps=eps)

    def forward(self, x):
        return E.module_embedding(self, x)


class Attention(nn.Module):
    """BigGAN non-local block: theta/phi/g 1x1 convs, 2x2 max-pool on phi and g,
    softmax(theta^T phi) without 1/sqrt(d), o conv, gamma*o + x (layers.py:262-300)."""

    def __init__(self, ch, which_conv=SNConv2d, name="attention"):
        super().__init__()
        self.ch, self.which_conv = ch, which_conv
        self.theta = which_conv(ch, ch // 8, kernel_size=1, padding=0, bias=False)
        self.phi = which_conv(ch, ch // 8, kernel_size=1, padding=0, bias=False)
        self.g = which_conv(ch, ch // 2, kernel_size=1, padding=0, bias=False)
        self.o = which_conv(ch // 2, ch, kernel_size=1, padding=0, bias=False)
        self.gamma = nn.Parameter(torch.tensor(0.0), requires_grad=True)

    def forward(self, x, y=None):
        return E.module_attention(self, x)


def fused_bn(x, mean, var, gain=None, bias=None, eps=1e-5):
    """x * rsqrt(var + eps) * gain - (mean * rsqrt(var + eps) * gain - bias) with per-channel mean / var
    (layers.py:505-517).  gain / bias: None, (1,C,1,1) or per-sample (N,C,1,1); gradients flow to x, gain and
    bias (mean / var are treated as given statistics, which is how myBN's eval branch uses this function)."""
    c = x.shape[1]
    return E.bn_functional(x, gain, bias, stored_mean=mean.detach().reshape(c).float().contiguous(),
                           stored_var=var.detach().reshape(c).float().contiguous(), training=False, mode=0, eps=eps,
                           momentum=0.0)


def manual_bn(x, gain=None, bias=None, return_mean_var=False, eps=1e-5):
    """Batch-norm from mean-of-squares statistics, differentiable through them (layers.py:522-542).
    With return_mean_var also (mean, biased var) -- (C,) for one event, (E, C) for a batch of E events."""
    out = E.bn_functional(x, gain, bias, stored_mean=None, stored_var=None, training=True, mode=1, eps=eps,
                          momentum=0.0, want_stats=return_mean_var)
    if return_mean_var:
        y, m, v = out
        return y, m.squeeze(0), v.squeeze(0)
    return out


class myBN(nn.Module):
    """The reference's hand-written BN with standing statistics (layers.py:547-599): running variance is the
    BIASED batch variance, accumulate_standing sums the statistics and counts the calls, eval divides."""

    def __init__(self, num_channels, eps=1e-5, momentum=0.1):
        super().__init__()
        self.momentum, self.eps = momentum, eps
        self.register_buffer("stored_mean", torch.zeros(num_channels))
        self.register_buffer("stored_var", torch.ones(num_channels))
        self.register_buffer("accumulation_counter", torch.zeros(1))
        self.accumulate_standing = False

    def reset_stats(self):
        self.stored_mean[:] = 0
        self.stored_var[:] = 0
        self.accumulation_counter[:] = 0

    def forward(self, x, gain, bias):
        return E.module_mybn(self, x, gain, bias)


class ccbn(nn.Module):
    """Class-conditional BN: batch statistics per event, gain = 1 + Lg(y), bias = Lb(y)
    applied after normalisation (layers.py:656-689)."""

    def __init__(self, output_size, input_size, which_linear, eps=1e-5, momentum=0.1,
                 cross_replica=False, mybn=False, norm_style="bn"):
        super().__init__()
        self.output_size, self.input_size = output_size, input_size
        self.gain = which_linear(input_size, output_size)
        self.bias = which_linear(input_size, output_size)
        self.eps, self.momentum = eps, momentum
        self.cross_replica, self.mybn, self.norm_style = cross_replica, mybn, norm_style
        if mybn:
            self.bn = myBN(output_size, eps, momentum)
        elif norm_style in ("bn", "in"):
            self.register_buffer("stored_mean", torch.zeros(output_size))
            self.register_buffer("stored_var", torch.ones(output_size))
        if cross_replica:
            raise NotImplementedError("cross_replica needs the sync_batchnorm package the reference does not ship "
                                      "(layers.py:647-648)")
        if norm_style != "bn":
            raise NotImplementedError("built: norm_style='bn' (config.json:18)")

    def forward(self, x, y):
        return E.module_ccbn(self, x, y)

    def extra_repr(self):
        return "out: {output_size}, in: {input_size}, cross_replica={cross_replica}".format(**self.__dict__)


class bn(nn.Module):
    """Plain BN with learnable per-channel gain and bias (layers.py:698-742)."""

    def __init__(self, output_size, eps=1e-5, momentum=0.1, cross_replica=False, mybn=False):
        super().__init__()
        self.output_size = output_size
        self.gain = nn.Parameter(torch.ones(output_size), requires_grad=True)
        self.bias = nn.Parameter(torch.zeros(output_size), requires_grad=True)
        self.eps, self.momentum, self.cross_replica, self.mybn = eps, momentum, cross_replica, mybn
        if cross_replica:
            raise NotImplementedError("cross_replica needs the sync_batchnorm package the reference does not ship "
                                      "(layers.py:720-721)")
        if mybn:
            self.bn = myBN(output_size, self.eps, self.momentum)
        else:
            self.register_buffer("stored_mean", torch.zeros(output_size))
            self.register_buffer("stored_var", torch.ones(output_size))

    def forward(self, x, y=None):
        return E.module_bn(self, x)


def prior(y, device="cuda", norm=True):
    raise NotImplementedError("layers.prior reads a features.csv the reference does not ship "
                              "(layers.py:19); prior_embed=False in config.json")
