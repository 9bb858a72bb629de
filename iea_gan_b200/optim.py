"""Per-step parameter update on the B200: flat gradient buffers, multi-tensor Adam (+ gradient-norm clipping)
and multi-tensor EMA -- SURVEY.md section 8(f) N2.

Reference behaviour restated here: `torch.optim.Adam(params, lr, betas=(B1, B2), weight_decay=0, eps=adam_eps)`
built in Generator / Discriminator.__init__ (model.py:410-416, 858-864), `torch.nn.utils.clip_grad_norm_`
(train_fns.py:133-136, 190-191) and `utils.apply_ema` (utils/__init__.py:809-837).  The arithmetic runs in
csrc/optim.cu (iea_mt_sqnorm / iea_mt_adam / iea_mt_lerp): three launches per net per step instead of ~10 per
tensor; there is no CPU path.

`FusedAdam` is a `torch.optim.Optimizer`: same constructor arguments, `param_groups`, `state_dict()` layout
(`step`, `exp_avg`, `exp_avg_sq` per parameter) as torch's Adam, so the reference's checkpoint code
(utils/__init__.py:689-726) and LR schedulers work on it unchanged.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from ._lib import ptr
from .engine import K as call  # (counts the launches for bench.py's gpu_launches and honours the per-call profiler)

CHUNK = 65536  # elements per multi-tensor work item (one 256-thread block each)


def _chunk_table(rows, device):
    """rows: (p, g, m, v, ema, numel) with tensors or None -> device table of iea_mt_chunk, count."""
    recs = []
    for ti, (p, g, m, v, e, n) in enumerate(rows):
        for o in range(0, n, CHUNK):
            recs.append((p.data_ptr() + 4 * o, g.data_ptr() + 4 * o if g is not None else 0,
                         m.data_ptr() + 4 * o if m is not None else 0, v.data_ptr() + 4 * o if v is not None else 0,
                         e.data_ptr() + 4 * o if e is not None else 0, min(CHUNK, n - o), ti))
    arr = np.zeros(len(recs), dtype=np.dtype([("p", "u8"), ("g", "u8"), ("m", "u8"), ("v", "u8"), ("ema", "u8"),
                                              ("n", "i4"), ("tensor", "i4")]))
    assert arr.dtype.itemsize == C.sizeof(L.MtChunk)
    for i, r in enumerate(recs):
        arr[i] = r
    return torch.from_numpy(arr.view(np.uint8)).to(device), len(recs)


class FlatGrads:
    """One fp32 buffer that holds the gradient of every parameter of a net; `p.grad` is a view into it.
    The weight-gradient kernels write straight into the views (engine.Tape.galloc), so the data-parallel
    all-reduce runs in place on ONE buffer with no pack / unpack and the optimizer's
    chunk table never changes."""

    def __init__(self, params):
        params = [p for p in params]
        self.device = params[0].device
        off, self.offsets = 0, {}
        for p in params:
            self.offsets[id(p)] = off
            off += (p.numel() + 63) // 64 * 64  # 256-byte aligned views: 16-byte vector accesses everywhere
        self.buf = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.views = {id(p): self.buf[self.offsets[id(p)]:self.offsets[id(p)] + p.numel()].view(p.shape) for p in params}
        self.ptrs = {k: v.data_ptr() for k, v in self.views.items()}
        self.key = tuple(p.data_ptr() for p in params)

    @staticmethod
    def of(net):
        """The net's flat gradient buffer, (re)built when the parameters moved (net.to(device) after construction)."""
        fl = net.__dict__.get("_iea_flat")
        ps = net.__dict__.get("_iea_params")
        if ps is None:
            ps = net.__dict__["_iea_params"] = list(net.parameters())
        if fl is None or fl.key != tuple(p.data_ptr() for p in ps):
            fl = net.__dict__["_iea_flat"] = FlatGrads(ps)
        return fl


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (no weight decay, no amsgrad) as two launches per parameter group; optional
    fused gradient-norm clipping: step(clip_norm=c) == clip_grad_norm_(params, c); step()."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("FusedAdam: weight_decay=0, amsgrad=False (model.py:410-416, 858-864)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))
        self._cache = {}

    def _group_state(self, gi, group, ps):
        dev = ps[0].device
        st = self._cache.get(gi)
        pkey = tuple(p.data_ptr() for p in ps)
        if st is None or st["pkey"] != pkey:
            # Adam moments: one buffer per group, per-parameter views exposed through self.state (checkpoint layout)
            total = sum((p.numel() + 63) // 64 * 64 for p in ps)
            scalars = torch.zeros(8, dtype=torch.float32, device=dev)
            mbuf = torch.zeros(total, dtype=torch.float32, device=dev)
            vbuf = torch.zeros(total, dtype=torch.float32, device=dev)
            off = 0
            for p in ps:
                n = p.numel()
                m, v = mbuf[off:off + n].view(p.shape), vbuf[off:off + n].view(p.shape)
                old = self.state.get(p)
                if old:  # resumed from a checkpoint or moved to another device: carry the moments over
                    m.copy_(old["exp_avg"])
                    v.copy_(old["exp_avg_sq"])
                    scalars[0] = float(old["step"])
                self.state[p] = {"step": scalars[0], "exp_avg": m, "exp_avg_sq": v}
                off += (n + 63) // 64 * 64
            st = self._cache[gi] = {"pkey": pkey, "gkey": None, "scalars": scalars, "mbuf": mbuf, "vbuf": vbuf,
                                    "hyper": torch.zeros(2, dtype=torch.float32, device=dev), "lr": None}
        gkey = tuple(p.grad.data_ptr() for p in ps)
        if st["gkey"] != gkey:  # (never again once the gradients live in a FlatGrads buffer)
            rows = [(p, p.grad, self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"], None, p.numel()) for p in ps]
            st["table"], st["n"] = _chunk_table(rows, dev)
            st["partial"] = torch.empty(st["n"], dtype=torch.float32, device=dev)
            st["gkey"] = gkey
        return st

    @torch.no_grad()
    def step(self, closure=None, clip_norm=None):
        loss = closure() if closure is not None else None
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = L.require_device(ps[0])
            with torch.cuda.device(dev):
                st = self._group_state(gi, group, ps)
                if st["lr"] != group["lr"] and not torch.cuda.is_current_stream_capturing():
                    st["hyper"][0:1].fill_(group["lr"])
                    st["lr"] = group["lr"]
                s = L.stream()
                part = None
                if clip_norm is not None:
                    call("iea_mt_sqnorm", ptr(st["table"]), st["n"], ptr(st["partial"]), s)
                    part = st["partial"]
                b1, b2 = group["betas"]
                call("iea_mt_adam", ptr(st["table"]), st["n"], ptr(part), float(clip_norm or 0.0), b1, b2,
                     group["eps"], ptr(st["scalars"]), ptr(st["hyper"]), s, launches=2)
        return loss

    def refresh_hyper(self):
        """Push a changed learning rate to the device scalar the (possibly graph-captured) kernels read."""
        for gi, group in enumerate(self.param_groups):
            st = self._cache.get(gi)
            if st is not None and st["lr"] != group["lr"]:
                st["hyper"][0:1].fill_(group["lr"])
                st["lr"] = group["lr"]

    def grad_norm(self, gi=0):
        """Total gradient norm measured by the last step(clip_norm=...) (a device scalar; no sync)."""
        return self._cache[gi]["scalars"][4]

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._cache = {}  # moments were replaced by the loaded tensors: rebuild the tables around them


class FusedEMA:
    """utils.apply_ema (utils/__init__.py:809-837) over whole state dicts as ONE launch:
    target = decay*target + (1-decay)*source for every floating-point entry (parameters, u0, sv0, running
    statistics); before `start_itr` the target is pegged to the source (decay 0), as in the reference."""

    def __init__(self, source, target, decay=0.9999, start_itr=0):
        self.source, self.target, self.decay, self.start_itr = source, target, decay, start_itr
        self._st = None
        sd, td = source.state_dict(), target.state_dict()
        with torch.no_grad():
            for k in sd:
                td[k].copy_(sd[k])

    def _state(self):
        sd, td = self.source.state_dict(), self.target.state_dict()
        keys = [k for k in sd if sd[k].is_floating_point()]
        key = tuple(sd[k].data_ptr() for k in keys) + tuple(td[k].data_ptr() for k in keys)
        if self._st is None or self._st["key"] != key:
            dev = sd[keys[0]].device
            L.require_device(sd[keys[0]])
            rows = [(td[k], sd[k], None, None, None, sd[k].numel()) for k in keys]
            table, n = _chunk_table(rows, dev)
            self._st = {"key": key, "table": table, "n": n, "hyper": torch.zeros(2, dtype=torch.float32, device=dev),
                        "decay": None, "device": dev}
        return self._st

    @torch.no_grad()
    def refresh_hyper(self, itr=None):
        """Decay switch at start_itr, pushed to the device scalar the kernel reads."""
        st = self._state()
        decay = 0.0 if (itr and itr < self.start_itr) else self.decay
        if st["decay"] != decay:
            with torch.cuda.device(st["device"]):
                st["hyper"][1:2].fill_(decay)
            st["decay"] = decay
        return st

    @torch.no_grad()
    def update(self, itr=None):
        capturing = torch.cuda.is_current_stream_capturing()
        st = self._state() if capturing else self.refresh_hyper(itr)  # (a captured fill would pin the decay)
        with torch.cuda.device(st["device"]):
            call("iea_mt_lerp", ptr(st["table"]), st["n"], ptr(st["hyper"]), L.stream())


class OrthoReg:
    """utils.ortho (utils/__init__.py:843-859) for a fixed parameter list as three grouped launches
    (csrc/ortho.cu).  Parameters with fewer than 2 axes and blacklisted ones (G.shared, train_fns.py:187) are
    skipped, as in the reference."""

    TALL = 2048  # rows above which W (W^T W) - diag(|w|^2) W replaces the rows x rows Gram matrix

    def __init__(self, params, strength=1e-4, blacklist=()):
        self.params = [p for p in params if p.dim() >= 2 and not any(p is b for b in blacklist)]
        self.strength = strength
        self._st = None

    def _state(self):
        ps = [p for p in self.params if p.grad is not None]
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
        if self._st is not None and self._st["key"] == key:
            return self._st
        dev = ps[0].device
        L.require_device(ps[0])
        items = (L.OrthoItem * len(ps))()
        geo, gsz, rsz, psz = [], 0, 0, 0
        for p in ps:
            rows, cols = p.shape[0], p.numel() // p.shape[0]
            tall = rows > cols and rows > self.TALL
            d = cols if tall else rows
            ks = max(1, min(32, rows // 512)) if tall else 1  # the long K of W^T W is split over ks CTAs per tile
            geo.append((rows, cols, tall, d, gsz, rsz, ks, psz))
            gsz += d * d
            rsz += rows if tall else 0
            psz += ks * d * d if tall else 0
        gram = torch.empty(max(gsz, 1), dtype=torch.float32, device=dev)
        rn = torch.empty(max(rsz, 1), dtype=torch.float32, device=dev)
        gpart = torch.empty(max(psz, 1), dtype=torch.float32, device=dev)
        rn_rows, gt, at, rb = [], [], [], []
        for i, (p, (rows, cols, tall, d, go, ro, ks, po)) in enumerate(zip(ps, geo)):
            it = items[i]
            it.w, it.grad, it.gram = p.data_ptr(), p.grad.data_ptr(), gram.data_ptr() + 4 * go
            it.rownorm = rn.data_ptr() + 4 * ro if tall else None
            it.gram_part = gpart.data_ptr() + 4 * po if tall else None
            it.rows, it.cols, it.tall, it.strength, it.ksplits = rows, cols, int(tall), self.strength, ks
            if tall:
                rn_rows += [(i, r) for r in range(rows)]
                rb += [(i, e) for e in range(0, d * d, 256)]
            td = (d + 63) // 64
            gt += [(i, a, b, k) for a in range(td) for b in range(td) for k in range(ks)]
            at += [(i, a, b, 0) for a in range((rows + 63) // 64) for b in range((cols + 63) // 64)]
        dev_i32 = lambda rows_: torch.tensor(np.array(rows_, dtype=np.int32).reshape(-1), dtype=torch.int32, device=dev)
        self._st = {"key": key, "items": torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8).to(dev),
                    "rn_rows": dev_i32(rn_rows) if rn_rows else None, "n_rn": len(rn_rows), "gt": dev_i32(gt),
                    "n_gt": len(gt), "rb": dev_i32(rb) if rb else None, "n_rb": len(rb), "at": dev_i32(at), "n_at": len(at),
                    "gram": gram, "rn": rn, "gpart": gpart, "device": dev}
        return self._st

    @torch.no_grad()
    def apply(self):
        if not any(p.grad is not None for p in self.params):
            return
        st = self._state()
        with torch.cuda.device(st["device"]):
            call("iea_ortho_grouped", ptr(st["items"]), ptr(st["rn_rows"]), st["n_rn"], ptr(st["gt"]), st["n_gt"],
                 ptr(st["rb"]), st["n_rb"], ptr(st["at"]), st["n_at"], L.stream(), launches=4 if st["n_rn"] else 2)


def ortho(model, strength=1e-4, blacklist=None):
    """Drop-in for utils.ortho(model, strength, blacklist) (utils/__init__.py:843-859; `utils.ortho =
    iea_gan_b200.optim.ortho` is the one-line hook-up): same gradient update, grouped kernels."""
    plans = model.__dict__.setdefault("_iea_ortho", {})
    bl = tuple(blacklist or ())
    key = (float(strength), tuple(id(b) for b in bl))
    plan = plans.get(key)
    if plan is None:
        plan = plans[key] = OrthoReg(model.parameters(), strength, bl)
    plan.apply()
