"""Drop-in module `RRM`: put this directory on sys.path ahead of the reference's
root and train.py / train_fns.py import the B200 implementation under the
reference's own module name."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from iea_gan_b200.relational import *  # noqa: F401,F403
from iea_gan_b200 import relational as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
