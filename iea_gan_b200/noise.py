"""Where the nets draw their in-forward noise.

The reference draws inside its forwards: `torch.randn(40, rdof_dim)` at the top of
Generator.forward (model.py:466) and three `torch.rand` + four `torch.randint` in DiffAugment
(diff_aug.py:25-85).  By default these modules issue exactly the same torch calls on the device
generator, so equal seeds give equal noise.  `replay(draws)` substitutes a prepared sequence (one
tensor per draw, consumed in call order) -- how the parity tests hand the CPU reference's draws to the
CUDA path, and how a caller replays recorded noise -- without touching torch's global functions.
"""
import threading

import torch

_state = threading.local()


def _next(kind, device):
    seq = getattr(_state, "seq", None)
    if seq is None:
        return None
    try:
        want, t = next(seq)
    except StopIteration:
        raise RuntimeError("noise.replay: the forward asked for more draws than were supplied (%s)" % kind)
    if want is not None and want != kind:
        raise RuntimeError("noise.replay: draw order mismatch, forward asks for %s, next supplied is %s" % (kind, want))
    return t.to(device)


def randn(rows, cols, device):
    t = _next("randn", device)
    return torch.randn(rows, cols, device=device) if t is None else t


def rand(size, device):
    t = _next("rand", device)
    return torch.rand(*size, dtype=torch.float32, device=device) if t is None else t.view(*size)


def randint(lo, hi, size, device):
    t = _next("randint", device)
    return torch.randint(lo, hi, size=size, device=device) if t is None else t.view(*size)


class replay:
    """with noise.replay([t0, t1, ...]) or [("randn", t0), ("rand", t1), ...]: the next draws, in order."""

    def __init__(self, draws):
        self.draws = [d if isinstance(d, tuple) else (None, d) for d in draws]

    def __enter__(self):
        self.prev = getattr(_state, "seq", None)
        _state.seq = iter(self.draws)
        return self

    def __exit__(self, *exc):
        _state.seq = self.prev
        return False
