"""DiffAugment (color, translation, cutout) as one fused, differentiable gather kernel.

Host-side mirror of the reference's diff_aug.py (diff_aug.py:10-20 DiffAugment,
:23-102 rand_*, :105-109 AUGMENT_FNS).  The random draws are the reference's
own torch.rand / torch.randint calls in the reference's order (so equal seeds
give equal augmentations); the arithmetic -- brightness, contrast about the
per-image mean, integer shift with zero fill, box cut-out -- runs in
libiea_sm100.so (iea_diffaug_fwd / iea_diffaug_bwd).  Saturation is the exact
identity for the single-channel PXD images and only consumes its draw.
"""
import torch

from . import engine as E
from . import noise


def _draw_color(x, which):
    return noise.rand((x.size(0), 1, 1, 1), x.device)


def _draw_translation(x, ratio=0.125):
    sx, sy = int(x.size(2) * ratio + 0.5), int(x.size(3) * ratio + 0.5)
    tx = noise.randint(-sx, sx + 1, [x.size(0), 1, 1], x.device)
    ty = noise.randint(-sy, sy + 1, [x.size(0), 1, 1], x.device)
    return tx, ty


def _draw_cutout(x, ratio=0.5):
    ch, cw = int(x.size(2) * ratio + 0.5), int(x.size(3) * ratio + 0.5)
    ox = noise.randint(0, x.size(2) + (1 - ch % 2), [x.size(0), 1, 1], x.device)
    oy = noise.randint(0, x.size(3) + (1 - cw % 2), [x.size(0), 1, 1], x.device)
    return ox, oy, ch, cw


def _apply(x, **draws):
    return E.diffaug_apply(x, draws)


def rand_brightness(x):
    return _apply(x, brightness=_draw_color(x, 0))


def rand_saturation(x):
    r = _draw_color(x, 1)
    if x.size(1) != 1:
        raise NotImplementedError("built for single-channel PXD images (saturation is the identity)")
    return _apply(x, saturation=r)


def rand_contrast(x):
    return _apply(x, contrast=_draw_color(x, 2))


def rand_translation(x, ratio=0.125):
    tx, ty = _draw_translation(x, ratio)
    return _apply(x, tx=tx, ty=ty)


def rand_cutout(x, ratio=0.5):
    ox, oy, ch, cw = _draw_cutout(x, ratio)
    return _apply(x, ox=ox, oy=oy, cut_h=ch, cut_w=cw)


AUGMENT_FNS = {
    "color": [rand_brightness, rand_saturation, rand_contrast],
    "translation": [rand_translation],
    "cutout": [rand_cutout],
}


def DiffAugment(x, policy="", channels_first=True):
    """Same draws, same order as diff_aug.py:10-20; the stages named in `policy` are
    executed by ONE fused kernel when they appear in the canonical order
    color -> translation -> cutout (the only order model.py:971 uses), otherwise
    stage by stage."""
    if not policy:
        return x
    if not channels_first:
        x = x.permute(0, 3, 1, 2)
    stages = policy.split(",")
    order = [s for s in ("color", "translation", "cutout") if s in stages]
    if stages == order and x.size(1) == 1:
        draws = {}
        for s in stages:
            if s == "color":
                draws["brightness"] = _draw_color(x, 0)
                draws["saturation"] = _draw_color(x, 1)
                draws["contrast"] = _draw_color(x, 2)
            elif s == "translation":
                draws["tx"], draws["ty"] = _draw_translation(x)
            else:
                draws["ox"], draws["oy"], draws["cut_h"], draws["cut_w"] = _draw_cutout(x)
        x = E.diffaug_apply(x, draws)
    else:
        for s in stages:
            for f in AUGMENT_FNS[s]:
                x = f(x)
    if not channels_first:
        x = x.permute(0, 2, 3, 1)
    return x.contiguous()
