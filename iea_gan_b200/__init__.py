"""iea_gan_b200 -- B200-native (sm_100a) hot path of IEA-GAN.

Drop-in replacements for the reference's `model`, `layers`, `RRM`, `diff_aug`
and `loss` modules (put `iea_gan_b200/dropin` on sys.path to import them under
those names).  All arithmetic runs in libiea_sm100.so (hand-written CUDA for
sm_100a, C ABI declared in include/iea_b200.h); there is no CPU fallback.
"""
from . import sn_layers, relational, augment, losses, nets  # noqa: F401
from .nets import Generator, Discriminator, G_D, Model, generate  # noqa: F401

__version__ = "0.1.0"
