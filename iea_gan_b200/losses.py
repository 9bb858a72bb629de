"""Hinge, 2C contrastive, IEA and uniformity losses as fused forward/backward kernels.

Host-side mirror of the reference's loss.py (loss.py:8-9 unif_loss, :14-27
IEA_loss, :30-38 hinge, :41-44 l2_loss, :79-132 Conditional_Contrastive_loss).
Rows are grouped by event (40 consecutive rows); a (40E, D) input yields the
mean of the E per-event losses, which for E = 1 is the reference's value.
No host round-trips: the reference's numpy diagonal mask (loss.py:91-97) is
index arithmetic inside the kernel.
"""
import torch

from . import engine as E


def unif_loss(x, t=2):
    return E.loss_uniformity(x, float(t))


def IEA_loss(k_f, k_r):
    return E.loss_iea(k_f, k_r)


def loss_hinge_dis(dis_fake, dis_real):
    return E.loss_hinge_dis(dis_fake, dis_real)


def loss_hinge_gen(dis_fake):
    return E.loss_hinge_gen(dis_fake)


def l2_loss(dis_real, dis_aug_real):
    return E.loss_l2(dis_real, dis_aug_real)


class Conditional_Contrastive_loss(torch.nn.Module):
    def __init__(self, device, batch_size, pos_collected_numerator):
        super().__init__()
        self.device, self.batch_size = device, batch_size
        self.pos_collected_numerator = pos_collected_numerator
        if pos_collected_numerator:
            raise NotImplementedError("built: pos_collected_numerator=False (config.json:87)")

    def forward(self, inst_embed, proxy, negative_mask, labels, temperature, margin):
        """-mean log( t * e^{(cos(e_i,p_i)-m)/t} / (e^{(cos(e_i,p_i)-m)/t} + sum_{j!=i} e^{(cos(e_i,e_j)-m)/t}) );
        negative_mask and labels are unused in this branch, as in the reference."""
        return E.loss_contrastive(inst_embed, proxy, float(temperature), float(margin))
