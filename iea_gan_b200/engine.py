"""Execution engine: schedules the fused sm_100a kernels of libiea_sm100.so for the
Generator / Discriminator forward and backward passes.

Design (B200-first, not a port of the reference's layer-by-layer autograd graph):
  * activations live in HBM as NHWC bf16 (fp32 when IEA_ACT_DTYPE=fp32 for tight
    parity runs); parameters stay fp32 in the reference's layout;
  * one grouped launch runs the spectral-norm power iteration of every layer of a
    net and repacks the weights for the conv kernels (`SNGroup`); W/sigma is never
    materialised, 1/sigma is an epilogue scale;
  * batch-norm statistics come out of the producer conv's epilogue and the
    normalise + gain/bias + ReLU (+ nearest-upsample / avg-pool) is the consumer
    conv's prologue, so every activation is written once and read once;
  * all 96 ccbn gain/bias linears of G are ONE grouped GEMM on the shared
    conditioning vector;
  * backward is a small explicit tape (`Tape`) per net call, bridged to torch
    autograd by one `autograd.Function` per net, so parameter gradients honour
    `requires_grad` toggling and `no_grad`.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L
from ._lib import call, ptr, dt

IMGS = 40  # images per event (model.py:466)


def act_dtype():
    return torch.float32 if os.environ.get("IEA_ACT_DTYPE", "bf16") == "fp32" else torch.bfloat16


def conv_impl():
    return {"auto": L.IMPL_AUTO, "generic": L.IMPL_GENERIC, "tcgen05": L.IMPL_TCGEN05}[
        os.environ.get("IEA_CONV_IMPL", "auto")]


LAUNCHES = [0]  # kernels launched through the C ABI (bench.py reports it)


def K(name, *args, launches=1):
    LAUNCHES[0] += launches
    return call(name, *args)


def cdiv(a, b):
    return (a + b - 1) // b


# ------------------------------------------------------------------ tape
class Var:
    """A tensor flowing through a net with an optional gradient slot.  The kernels see it as a
    matrix of rows x ld with a channel window [c0, c0+c)."""
    __slots__ = ("t", "c0", "c", "ld", "g", "ds", "bn", "need", "sc_var")

    def __init__(self, t, need=True, c0=0, c=None):
        self.t = t
        self.ld = t.shape[-1]
        self.c0 = c0
        self.c = self.ld if c is None else c
        self.g = None
        self.ds = None      # (ds1, ds2, rows_per_event): batch-norm statistics gradient
        self.bn = None      # (partials, tiles_per_event, count_per_event)
        self.need = need
        self.sc_var = None  # Var of a shortcut conv that wrote a channel window of this tensor

    def off(self, t=None):
        t = self.t if t is None else t
        return t.data_ptr() + self.c0 * t.element_size()


class Tape:
    def __init__(self, record):
        self.record = record
        self.nodes = []
        self.pgrads = {}

    def add(self, fn):
        if self.record:
            self.nodes.append(fn)

    def backward(self):
        for fn in reversed(self.nodes):
            fn()
        self.nodes = []

    def pgrad(self, param, g):
        k = id(param)
        if k in self.pgrads:
            acc = self.pgrads[k]
            K("iea_axpby", ptr(g), L.F32, 1.0, ptr(acc), L.F32, 1.0, ptr(acc), L.F32, g.numel(), L.stream())
        else:
            self.pgrads[k] = g


def _accum_target(v):
    """(gradient buffer of v, beta): beta = 1 when a previous consumer already wrote it."""
    if v.g is None:
        v.g = torch.empty_like(v.t)
        return v.g, 0.0
    return v.g, 1.0


def add_grad(v, g):
    if v.g is None:
        v.g = g
    else:
        K("iea_axpby", ptr(g), dt(g), 1.0, ptr(v.g), dt(v.g), 1.0, ptr(v.g), dt(v.g), g.numel(), L.stream())


def _f32(n, device):
    return torch.empty(max(int(n), 1), dtype=torch.float32, device=device)


# ------------------------------------------------------------------ spectral norm group
class SNLayer:
    """One (optionally spectrally-normalised) weight inside an SNGroup."""
    __slots__ = ("mod", "weight", "rows", "cin", "taps", "spectral", "pack_dtype", "wp", "wd", "index",
                 "colscale", "group", "wd_ld", "shared")

    def inv_sigma(self):
        return self.group.cur[0][self.index:self.index + 1]

    def u(self):
        o = self.group.u_off[self.index]
        return self.group.cur[1][o:o + self.rows]

    def v(self):
        o = self.group.v_off[self.index]
        return self.group.cur[2][o:o + self.cin * self.taps]

    def saved(self):
        """(inv_sigma, u', v) of the CURRENT forward call, to be captured by backward closures."""
        return self.inv_sigma(), self.u(), self.v()


class SNGroup:
    """All weights of a net that go through the grouped power-iteration / repack kernel
    (layers.py:89-165).  Built once per net; `run()` is called at the top of each forward."""

    def __init__(self):
        self.layers = []
        self.key = None
        self.shared_packs = []
        self.tables = {}
        self.cur = None

    def add(self, mod, pack_dtype):
        w = mod.weight
        l = SNLayer()
        l.mod, l.weight = mod, w
        l.rows = w.shape[0]
        l.cin, l.taps = (w.shape[1], w.shape[2] * w.shape[3]) if w.dim() == 4 else (w.shape[1], 1)
        l.spectral = 1 if hasattr(mod, "u0") else 0
        l.pack_dtype = pack_dtype
        l.index = len(self.layers)
        l.group = self
        l.colscale = None
        l.wp = l.wd = None
        l.wd_ld = l.rows
        l.shared = False
        self.layers.append(l)
        return l

    def add_shared(self, mods, pack_dtype):
        """Several linears on the same input packed into one [sum rows][cin] matrix with a
        per-output-column 1/sigma vector (the grouped ccbn GEMM)."""
        ls = [self.add(m, pack_dtype) for m in mods]
        for l in ls:
            l.shared = True
        self.shared_packs.append(ls)
        return ls

    def _signature(self):
        ls = self.layers
        return (ls[0].weight.data_ptr(), ls[-1].weight.data_ptr(), ls[0].weight.device, len(ls))

    def _build(self, device):
        ls = self.layers
        total = {torch.float32: 0, torch.bfloat16: 0}
        al = lambda n: (n + 63) // 64 * 64
        offs = []
        for l in ls:
            if l.shared:
                offs.append(None)
                continue
            offs.append(total[l.pack_dtype])
            total[l.pack_dtype] += 2 * al(l.rows * l.cin * l.taps)
        shared_info = []
        for grp in self.shared_packs:
            rows, cin = sum(l.rows for l in grp), grp[0].cin
            shared_info.append((total[grp[0].pack_dtype], rows, cin))
            total[grp[0].pack_dtype] += 2 * al(rows * cin)
        self.packs = {d: torch.zeros(max(n, 1), dtype=d, device=device) for d, n in total.items()}
        self.colscales = []
        for (o, rows, cin), grp in zip(shared_info, self.shared_packs):
            buf = self.packs[grp[0].pack_dtype]
            n = rows * cin
            wp = buf[o:o + n].view(rows, cin)
            wd = buf[o + al(n):o + al(n) + n].view(cin, rows)
            cs = torch.ones(rows, dtype=torch.float32, device=device)
            self.colscales.append((wp, wd, cs))
            r0 = 0
            for l in grp:
                l.wp, l.wd, l.wd_ld, l.colscale = wp[r0:r0 + l.rows], wd[:, r0:r0 + l.rows], rows, cs[r0:r0 + l.rows]
                r0 += l.rows
        for l, o in zip(ls, offs):
            if o is None:
                continue
            buf = self.packs[l.pack_dtype]
            n = l.rows * l.cin * l.taps
            l.wp = buf[o:o + n].view(l.rows, l.taps, l.cin)
            l.wd = buf[o + al(n):o + al(n) + n].view(l.cin, l.taps, l.rows)
        chunks, metas, scratch, uo, vo = [], [], 0, 0, 0
        self.u_off, self.v_off = [], []
        for i, l in enumerate(ls):
            cols = l.cin * l.taps
            rows_per = max(1, 32768 // cols)
            c0 = len(chunks)
            for r in range(0, l.rows, rows_per):
                chunks.append((i, r, min(l.rows, r + rows_per)))
            metas.append((c0, len(chunks) - c0, scratch))
            scratch += (len(chunks) - c0) * cols + l.rows
            self.u_off.append(uo)
            self.v_off.append(vo)
            uo += l.rows
            vo += cols
        self.metas = metas
        self.chunks = torch.tensor(np.array(chunks, dtype=np.int32).reshape(-1), dtype=torch.int32, device=device)
        self.n_chunks = len(chunks)
        self.scratch = torch.empty(scratch, dtype=torch.float32, device=device)
        self.max_cols = max(l.cin * l.taps for l in ls)
        self.inv_sigma = torch.ones(len(ls), dtype=torch.float32, device=device)
        self.u_new = torch.zeros(uo, dtype=torch.float32, device=device)
        self.v_new = torch.zeros(vo, dtype=torch.float32, device=device)
        self.sig_scratch = torch.zeros(len(ls), dtype=torch.float32, device=device)
        self.tables = {}
        self.u_dst = [l.mod.u0.view(-1) for l in ls if l.spectral]
        self.u_src = [self.u_new[self.u_off[l.index]:self.u_off[l.index] + l.rows] for l in ls if l.spectral]

    def _table(self, training, need_bwd):
        key = (bool(training), bool(need_bwd))
        if key not in self.tables:
            ls = self.layers
            arr = (L.SnLayer * len(ls))()
            for i, l in enumerate(ls):
                a = arr[i]
                c0, nch, so = self.metas[i]
                a.w = l.weight.data_ptr()
                if l.spectral:
                    a.u_in = l.mod.u0.data_ptr()
                    a.sigma_out = l.mod.sv0.data_ptr() if training else self.sig_scratch.data_ptr() + 4 * i
                a.u_out = self.u_new.data_ptr() + 4 * self.u_off[i]
                a.v_out = self.v_new.data_ptr() + 4 * self.v_off[i]
                a.inv_sigma_out = self.inv_sigma.data_ptr() + 4 * i
                if l.colscale is not None:
                    a.colscale_out, a.colscale_n = l.colscale.data_ptr(), l.rows
                a.pack_fprop = l.wp.data_ptr()
                a.pack_dgrad = l.wd.data_ptr() if need_bwd else None
                a.rows, a.cin, a.taps, a.pack_dgrad_ld = l.rows, l.cin, l.taps, l.wd_ld
                a.pack_dtype = L.F32 if l.pack_dtype == torch.float32 else L.BF16
                a.spectral = l.spectral
                a.eps = float(l.mod.eps) if l.spectral else 0.0
                a.chunk0, a.nchunks, a.scratch_off = c0, nch, so
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self.tables[key] = host.to(self.inv_sigma.device)
        return self.tables[key]

    def run(self, training, need_bwd):
        """One grouped power iteration + repack (4 launches for the whole net).  After it the layer
        handles expose wp / wd and, through saved(), THIS call's (1/sigma, u', v): when a backward
        will follow they are snapshots, so two forwards before one backward keep their own values."""
        ls = self.layers
        L.require_device(ls[0].weight)
        sig = self._signature()
        if sig != self.key:
            self._build(ls[0].weight.device)
            self.key = sig
        K("iea_sn_power_iter", ptr(self._table(training, need_bwd)), len(ls), ptr(self.chunks), self.n_chunks,
          ptr(self.scratch), self.max_cols, L.stream(), launches=4)
        if training and self.u_dst:  # u0 <- u' (layers.py:106-107); sv0 was written by the kernel
            torch._foreach_copy_(self.u_dst, self.u_src)
        if need_bwd:
            self.cur = (self.inv_sigma.clone(), self.u_new.clone(), self.v_new.clone())
        else:
            self.cur = (self.inv_sigma, self.u_new, self.v_new)


# ------------------------------------------------------------------ fused conv op
def _desc(n, h, w, cin, cout, k, x_t, x_ptr, x_ld, in_mode, in_relu, scale, shift, wp, out_scale,
          out_scale_stride, bias, res, res_mode, res_c, acc_c0, y_t, y_ptr, y_ld, act, stats, in_bcast=0):
    d = L.ConvDesc()
    d.n, d.h, d.w, d.cin, d.cout, d.ksize = n, h, w, cin, cout, k
    d.x, d.x_dtype, d.x_ld, d.in_mode, d.in_relu = x_ptr, dt(x_t), x_ld, in_mode, int(in_relu)
    d.in_scale, d.in_shift, d.in_bcast = ptr(scale), ptr(shift), in_bcast
    d.wpack, d.w_dtype = wp.data_ptr(), dt(wp)
    d.out_scale, d.out_scale_stride = ptr(out_scale), out_scale_stride
    d.bias = ptr(bias)
    if res is not None:
        d.res, d.res_dtype, d.res_ld, d.res_mode, d.res_c = res.off(), dt(res.t), res.ld, res_mode, res_c
    d.acc_c0 = acc_c0
    d.y, d.y_dtype, d.y_ld, d.act = y_ptr, dt(y_t), y_ld, act
    d.stats = ptr(stats)
    d.impl = conv_impl()
    return d


class ScaleShift:
    """Per-(image, channel) affine of a batch-norm, consumed by a conv prologue."""
    __slots__ = ("scale", "shift", "dscale", "dshift")

    def __init__(self, scale, shift):
        self.scale, self.shift, self.dscale, self.dshift = scale, shift, None, None


def conv(tape, xv, layer, n, h, w, k, *, bias=None, in_mode=L.IN_DIRECT, in_relu=False, ss=None, res=None,
         res_mode=L.IN_DIRECT, res_c=0, acc_c0=-1, out=None, act=L.ACT_NONE, stats=False, out_dtype=None,
         out_shape=None, grouped=None):
    """y = act(conv_k(T(x)) * (1/sigma) + bias + residual); see iea_conv_desc in include/iea_b200.h.
    `layer` is an SNLayer.  grouped = (wp, wd, colscale, cout, layers) runs several layers that share
    the input as one GEMM with a per-output-column 1/sigma (the ccbn gain/bias linears)."""
    cin = xv.c
    dev = xv.t.device
    if grouped is not None:
        wp_t, wd_t, colscale, cout, wls = grouped
        osc, oss = colscale, 1
    else:
        wp_t, wd_t, colscale, cout, wls = layer.wp, layer.wd, None, layer.rows, [layer]
        osc, oss = layer.inv_sigma(), 0
    if out is None:
        od = out_dtype or act_dtype()
        yv = Var(torch.empty(out_shape or (n, h, w, cout), dtype=od, device=dev))
    else:
        yv = out
    y = yv.t
    M = n * h * w
    st = torch.empty((cdiv(M, 128), cout, 2), dtype=torch.float32, device=dev) if stats else None
    scale = ss.scale if ss is not None else None
    shift = ss.shift if ss is not None else None
    d = _desc(n, h, w, cin, cout, k, xv.t, xv.off(), xv.ld, in_mode, in_relu, scale, shift, wp_t, osc, oss,
              bias, res, res_mode, res_c, acc_c0, y, yv.off(), yv.ld, act, st)
    K("iea_conv_fprop", C.byref(d), L.stream())
    if stats:
        rpe = IMGS * h * w
        yv.bn = (st, rpe // 128, rpe) if rpe % 128 == 0 else None
    if not tape.record:
        return yv
    saved = [l.saved() for l in wls]

    def fwd_desc():  # the forward geometry / prologue, for wgrad and the prologue backward
        return _desc(n, h, w, cin, cout, k, xv.t, xv.off(), xv.ld, in_mode, in_relu, scale, shift, wp_t,
                     None, 0, None, None, 0, 0, -1, y, yv.off(), yv.ld, 0, None)

    def dgrad_into(g_t, g_ptr, g_ld, dst, dst_ptr, accumulate):
        if colscale is not None:  # per-input-column 1/sigma of the grouped GEMM
            dd = _desc(n, h, w, cout, cin, k, g_t, g_ptr, g_ld, 0, 0, colscale, torch.zeros_like(colscale), wd_t,
                       None, 0, None, None, 0, 0, 0 if accumulate else -1, dst, dst_ptr, cin, 0, None, in_bcast=1)
        else:
            dd = _desc(n, h, w, cout, cin, k, g_t, g_ptr, g_ld, 0, 0, None, None, wd_t, saved[0][0], 0,
                       None, None, 0, 0, 0 if accumulate else -1, dst, dst_ptr, cin, 0, None)
        K("iea_conv_fprop", C.byref(dd), L.stream())

    def bw():
        if yv.g is None:
            return
        g_t = yv.g
        g_ptr, g_ld = yv.off(g_t), yv.ld
        if yv.ds is not None or act == L.ACT_TANH:
            ds1, ds2, rpe_ = yv.ds if yv.ds is not None else (None, None, 1)
            geff = torch.empty((M, cout), dtype=g_t.dtype, device=dev)
            K("iea_conv_out_bwd", g_ptr, dt(g_t), g_ld, yv.off(), dt(y), yv.ld, act, ptr(ds1), ptr(ds2), M, rpe_,
              cout, ptr(geff), dt(geff), L.stream())
            g_t, g_ptr, g_ld = geff, geff.data_ptr(), cout
        if bias is not None and bias.requires_grad:
            db = torch.empty(cout, dtype=torch.float32, device=dev)
            K("iea_colsum", g_ptr, dt(g_t), g_ld, M, cout, ptr(db), 0.0, ptr(_f32(300 * cout, dev)), L.stream(),
              launches=2)
            tape.pgrad(bias, db)
        if res is not None and res.need:
            rg, beta = _accum_target(res)
            K("iea_residual_bwd", g_ptr, dt(g_t), g_ld, n, h, w, res_c, res_mode, res.off(rg), dt(rg), res.ld,
              res.c, beta, L.stream())
        if yv.sc_var is not None:
            yv.sc_var.g = g_t  # the shortcut conv wrote channels >= acc_c0 of y: same gradient tensor
        if any(l.weight.requires_grad for l in wls):
            kdim = cin * k * k
            nsplit = max(1, min(64, M // 4096))
            gpart = torch.empty((nsplit, cout, kdim), dtype=torch.float32, device=dev)
            dfw = fwd_desc()
            K("iea_conv_wgrad", C.byref(dfw), g_ptr, dt(g_t), g_ld, ptr(gpart), nsplit, L.stream())
            r0 = 0
            for l, (isg, u_, v_) in zip(wls, saved):
                if l.weight.requires_grad:
                    dw = torch.empty_like(l.weight)
                    gp = gpart if len(wls) == 1 else gpart[:, r0:r0 + l.rows].contiguous()
                    K("iea_sn_weight_bwd", ptr(gp), nsplit, ptr(l.weight), ptr(u_), ptr(v_), ptr(isg), l.spectral,
                      ptr(dw), 0.0, l.rows, l.cin, l.taps, ptr(_f32(520, dev)), L.stream(), launches=2)
                    tape.pgrad(l.weight, dw)
                r0 += l.rows
        plain = in_mode == L.IN_DIRECT and not in_relu and ss is None and xv.c == xv.ld
        if xv.need and plain:
            xg, beta = _accum_target(xv)
            dgrad_into(g_t, g_ptr, g_ld, xg, xg.data_ptr(), beta != 0.0)
        elif xv.need or ss is not None:
            da = torch.empty((M, cin), dtype=g_t.dtype, device=dev)
            dgrad_into(g_t, g_ptr, g_ld, da, da.data_ptr(), False)
            xg, beta, xg_ptr = None, 0.0, None
            if xv.need:
                xg, beta = _accum_target(xv)
                xg_ptr = xv.off(xg)
            dsc = dsh = None
            if ss is not None:
                dsc, dsh = torch.empty_like(ss.scale), torch.empty_like(ss.shift)
                ss.dscale, ss.dshift = dsc, dsh
            dfw = fwd_desc()
            K("iea_conv_input_bwd", C.byref(dfw), ptr(da), dt(da), xg_ptr, dt(xg) if xg is not None else 0, xv.ld,
              beta, ptr(dsc), ptr(dsh), L.stream())
    tape.add(bw)
    return yv


# ------------------------------------------------------------------ batch norm glue
def ensure_stats(xv, n, h, w):
    """Batch-norm partial sums of xv: from the producer's epilogue when present, else one
    stand-alone pass."""
    if xv.bn is None:
        rpe = IMGS * h * w
        tiles = max(1, min(64, rpe // 512))
        part = torch.empty((n // IMGS, tiles, xv.c, 2), dtype=torch.float32, device=xv.t.device)
        K("iea_bn_stats", xv.off(), dt(xv.t), xv.ld, n * h * w, rpe, xv.c, tiles, ptr(part), L.stream())
        xv.bn = (part, tiles, rpe)
    return xv.bn


def bn_affine(tape, xv, n, h, w, *, gain, gain_ld, gain_add, bias, bias_ld, stored_mean, stored_var, training,
              eps, momentum=0.1, dgain=None, dbias=None, dgb_ld=0, gain_param=None, bias_param=None):
    """Finalize the batch statistics of xv into a per-(n,c) ScaleShift (layers.py:656-689, 728-742).
    ccbn: gain/bias are raw device addresses of column windows of the grouped-GEMM output (row
    stride gain_ld) and dgain/dbias the matching windows of its gradient buffer; plain bn:
    gain_param/bias_param are the (C,) parameters."""
    c, dev, events = xv.c, xv.t.device, n // IMGS
    part, tiles, count = ensure_stats(xv, n, h, w) if training else (None, 0, IMGS * h * w)
    scale = torch.empty((n, c), dtype=torch.float32, device=dev)
    shift = torch.empty((n, c), dtype=torch.float32, device=dev)
    mean = torch.empty((events, c), dtype=torch.float32, device=dev)
    rstd = torch.empty((events, c), dtype=torch.float32, device=dev)
    K("iea_bn_finalize", ptr(part), events, tiles, count, IMGS, c, gain, gain_ld, gain_add, bias, bias_ld,
      ptr(stored_mean), ptr(stored_var), int(training), momentum, eps, ptr(mean), ptr(rstd), ptr(scale), ptr(shift),
      L.stream())
    ss = ScaleShift(scale, shift)
    if tape.record:
        def bw():
            if ss.dscale is None:
                return
            ds1 = torch.empty((events, c), dtype=torch.float32, device=dev)
            ds2 = torch.empty((events, c), dtype=torch.float32, device=dev)
            if gain_param is not None:
                dg = torch.empty(c, dtype=torch.float32, device=dev)
                db = torch.empty(c, dtype=torch.float32, device=dev)
                K("iea_bn_finalize_bwd", ptr(ss.dscale), ptr(ss.dshift), ptr(scale), ptr(mean), ptr(rstd), events,
                  IMGS, count, c, gain, gain_ld, gain_add, ptr(dg), 0, ptr(db), 0, 1, int(training), ptr(ds1),
                  ptr(ds2), L.stream())
                if gain_param.requires_grad:
                    tape.pgrad(gain_param, dg)
                if bias_param.requires_grad:
                    tape.pgrad(bias_param, db)
            else:
                K("iea_bn_finalize_bwd", ptr(ss.dscale), ptr(ss.dshift), ptr(scale), ptr(mean), ptr(rstd), events,
                  IMGS, count, c, gain, gain_ld, gain_add, dgain, dgb_ld, dbias, dgb_ld, 0, int(training),
                  ptr(ds1), ptr(ds2), L.stream())
            if training and xv.need:
                xv.ds = (ds1, ds2, count)
        tape.add(bw)
    return ss


# ------------------------------------------------------------------ small ops with tape
def layernorm(tape, xv, ln):
    x = xv.t
    rows, dim = x.shape[0], x.shape[-1]
    y = torch.empty_like(x)
    mean, rstd = _f32(rows, x.device), _f32(rows, x.device)
    K("iea_layernorm_fwd", ptr(x), ptr(ln.weight), ptr(ln.bias), rows, dim, ln.eps, ptr(y), ptr(mean), ptr(rstd),
      L.stream())
    yv = Var(y)
    if tape.record:
        def bw():
            if yv.g is None:
                return
            need_p = ln.weight.requires_grad
            dx = torch.empty_like(x)
            dg = _f32(dim, x.device) if need_p else None
            db = _f32(dim, x.device) if need_p else None
            K("iea_layernorm_bwd", ptr(yv.g), ptr(x), ptr(ln.weight), ptr(mean), ptr(rstd), rows, dim, ptr(dx),
              ptr(dg), ptr(db), 1, L.stream(), launches=2)
            if need_p:
                tape.pgrad(ln.weight, dg)
                tape.pgrad(ln.bias, db)
            if xv.need:
                add_grad(xv, dx)
        tape.add(bw)
    return yv


def mha_core(tape, qkvv, events, seq, heads, d):
    qkv = qkvv.t
    val = torch.empty((events * seq, heads * d), dtype=torch.float32, device=qkv.device)
    att = torch.empty((events, heads, seq, seq), dtype=torch.float32, device=qkv.device)
    K("iea_mha_fwd", ptr(qkv), events, seq, heads, d, ptr(val), ptr(att), L.stream())
    vv = Var(val)
    if tape.record:
        def bw():
            if vv.g is None:
                return
            dq = torch.empty_like(qkv)
            K("iea_mha_bwd", ptr(vv.g), ptr(qkv), ptr(att), events, seq, heads, d, ptr(dq), L.stream())
            add_grad(qkvv, dq)
        tape.add(bw)
    return vv, att


def linear(tape, xv, layer, *, bias=None, in_relu=False, res=None, out_dtype=torch.float32):
    """F.linear(x, W/sigma, b) [+ residual] on a (rows, K) feature matrix: the 1x1 conv with h=w=1."""
    n = xv.t.shape[0]
    return conv(tape, xv, layer, n, 1, 1, 1, bias=bias, in_relu=in_relu, res=res,
                res_c=layer.rows if res is not None else 0, out_dtype=out_dtype, out_shape=(n, layer.rows))


def rrm(tape, xv, blocks, final_norm, sn):
    """RelationalReasoning.forward on rows grouped by event: x (40E, dim) fp32 (RRM.py:98-125)."""
    events = xv.t.shape[0] // IMGS
    for blk in blocks:
        at = blk.self_attn
        h1 = layernorm(tape, xv, blk.norm1)
        qkv = linear(tape, h1, sn[at.qkv_proj], bias=at.qkv_proj.bias)
        val, _ = mha_core(tape, qkv, events, IMGS, at.num_heads, at.head_dim)
        x1 = linear(tape, val, sn[at.o_proj], bias=at.o_proj.bias, res=xv)
        h2 = layernorm(tape, x1, blk.norm2)
        l0, l3 = blk.linear_net[0], blk.linear_net[3]
        f = linear(tape, h2, sn[l0], bias=l0.bias)
        xv = linear(tape, f, sn[l3], bias=l3.bias, in_relu=True, res=x1)  # ReLU fused as the prologue
    if final_norm is not None:
        xv = layernorm(tape, xv, final_norm)
    return xv


def cat_cols(tape, a, b):
    """torch.cat([a, b], 1) on small fp32 feature matrices (host-side plumbing)."""
    v = Var(torch.cat([a.t, b.t], 1))
    if tape.record:
        ca = a.t.shape[1]

        def bw():
            if v.g is None:
                return
            if a.need:
                add_grad(a, v.g[:, :ca].contiguous())
            if b.need:
                add_grad(b, v.g[:, ca:].contiguous())
        tape.add(bw)
    return v


def embedding(tape, idx, weight, sn_layer=None):
    """F.embedding(idx, W [/sigma]) (model.py:462; layers.py:259 for the SN variant)."""
    n, dim = idx.shape[0], weight.shape[1]
    scale = sn_layer.inv_sigma() if sn_layer is not None else None
    out = torch.empty((n, dim), dtype=torch.float32, device=weight.device)
    K("iea_embedding_fwd", ptr(idx), ptr(weight), ptr(scale), n, dim, ptr(out), L.stream())
    v = Var(out)
    if tape.record and weight.requires_grad:
        saved = sn_layer.saved() if sn_layer is not None else None

        def bw():
            if v.g is None:
                return
            G = torch.empty_like(weight)  # gradient w.r.t. the normalised table
            K("iea_embedding_bwd", ptr(idx), ptr(v.g), None, n, dim, weight.shape[0], ptr(G), L.stream())
            if saved is not None:
                dw = torch.empty_like(weight)
                K("iea_sn_weight_bwd", ptr(G), 1, ptr(weight), ptr(saved[1]), ptr(saved[2]), ptr(saved[0]), 1,
                  ptr(dw), 0.0, weight.shape[0], dim, 1, ptr(_f32(520, weight.device)), L.stream(), launches=2)
                G = dw
            tape.pgrad(weight, G)
        tape.add(bw)
    return v


def nchw_to_nhwc(tape, xv, n, c, hh, ww):
    """(n, c*hh*ww) features viewed NCHW by the reference (model.py:477-479) -> NHWC activation."""
    src = xv.t
    out = torch.empty((n, hh, ww, c), dtype=act_dtype(), device=src.device)
    K("iea_nchw_to_nhwc", ptr(src), dt(src), ptr(out), dt(out), n, c, hh * ww, L.stream())
    ov = Var(out)
    if tape.record:
        def bw():
            if ov.g is None:
                return
            g = ov.g
            if ov.ds is not None:
                ds1, ds2, rpe = ov.ds
                ge = torch.empty_like(g)
                K("iea_conv_out_bwd", ptr(g), dt(g), c, ptr(out), dt(out), c, 0, ptr(ds1), ptr(ds2), n * hh * ww,
                  rpe, c, ptr(ge), dt(ge), L.stream())
                g = ge
            dx = torch.empty((n, c * hh * ww), dtype=torch.float32, device=src.device)
            K("iea_nhwc_to_nchw", ptr(g), dt(g), ptr(dx), L.F32, n, c, hh * ww, L.stream())
            add_grad(xv, dx)
        tape.add(bw)
    return ov


# ------------------------------------------------------------------ Generator
class GPlan:
    """Static per-Generator schedule: the SN group and the grouped ccbn GEMM layout."""

    def __init__(self, G):
        from . import sn_layers as SL
        self.sn = SNGroup()
        self.h = {}
        big = act_dtype()
        self.h[G.linear_f] = self.sn.add(G.linear_f, torch.float32)
        for blk in G.RR_G.layers:
            for m in (blk.self_attn.qkv_proj, blk.self_attn.o_proj, blk.linear_net[0], blk.linear_net[3]):
                self.h[m] = self.sn.add(m, torch.float32)
        self.h[G.linear] = self.sn.add(G.linear, big)
        self.ccbn = []
        for bl in G.blocks:
            for m in bl:
                if isinstance(m, SL.Attention):
                    raise NotImplementedError("G_attn: the shipped config has no attention in G (config.json:27)")
                for cv in (m.conv1, m.conv2, m.conv3, m.conv4):
                    self.h[cv] = self.sn.add(cv, big)
                self.ccbn += [m.bn1, m.bn2, m.bn3, m.bn4]
        oc = G.output_layer[2]
        self.h[oc] = self.sn.add(oc, big)
        mods, off, self.gb_off = [], 0, {}
        for b in self.ccbn:
            self.gb_off[b] = (off, off + b.output_size)
            off += 2 * b.output_size
            mods += [b.gain, b.bias]
        self.gb_cols = off
        self.gb_layers = self.sn.add_shared(mods, torch.float32)
        for m, l in zip(mods, self.gb_layers):
            self.h[m] = l


def _plan(net, cls):
    p = net.__dict__.get("_iea_plan")
    if p is None or p[0] != act_dtype():
        p = (act_dtype(), cls(net))
        net.__dict__["_iea_plan"] = p
    return p[1]


def _gblock(tape, blk, x, n, hh, ww, plan, gb, dgb, training):
    """One bottleneck GBlock as 4 fused convs (model.py:54-71): every ccbn+ReLU(+upsample) is the
    prologue of the conv that consumes it, every conv's epilogue emits the next BN's statistics, and
    the channel-dropped (upsampled) skip is the epilogue residual of conv4."""
    h_ = plan.h
    up = blk.upsample is not None
    gld = plan.gb_cols

    def aff(bnm, xv, h, w):
        g0, b0 = plan.gb_off[bnm]
        return bn_affine(tape, xv, n, h, w, gain=gb.data_ptr() + 4 * g0, gain_ld=gld, gain_add=1.0,
                         bias=gb.data_ptr() + 4 * b0, bias_ld=gld, stored_mean=bnm.stored_mean,
                         stored_var=bnm.stored_var, training=training, eps=bnm.eps,
                         dgain=(dgb.data_ptr() + 4 * g0) if dgb is not None else None,
                         dbias=(dgb.data_ptr() + 4 * b0) if dgb is not None else None, dgb_ld=gld)
    h1 = conv(tape, x, h_[blk.conv1], n, hh, ww, 1, bias=blk.conv1.bias, in_relu=True,
              ss=aff(blk.bn1, x, hh, ww), stats=True)
    ho, wo = (hh * 2, ww * 2) if up else (hh, ww)
    h2 = conv(tape, h1, h_[blk.conv2], n, ho, wo, 3, bias=blk.conv2.bias, in_relu=True,
              ss=aff(blk.bn2, h1, hh, ww), in_mode=L.IN_UP2 if up else L.IN_DIRECT, stats=True)
    h3 = conv(tape, h2, h_[blk.conv3], n, ho, wo, 3, bias=blk.conv3.bias, in_relu=True,
              ss=aff(blk.bn3, h2, ho, wo), stats=True)
    out = conv(tape, h3, h_[blk.conv4], n, ho, wo, 1, bias=blk.conv4.bias, in_relu=True,
               ss=aff(blk.bn4, h3, ho, wo), res=x, res_mode=L.IN_UP2 if up else L.IN_DIRECT,
               res_c=blk.out_channels, stats=True)
    return out, ho, wo


def _g_body(G, tape, z, y, rdof):
    plan = _plan(G, GPlan)
    sn, h_ = plan.sn, plan.h
    training = G.training
    n = z.t.shape[0]
    if n % IMGS:
        raise ValueError("a batch of %d rows is not a whole number of 40-image events" % n)
    sn.run(training, tape.record)
    emb = embedding(tape, y, G.shared.weight)
    c = cat_cols(tape, emb, Var(rdof, need=False))
    c = linear(tape, c, h_[G.linear_f], bias=G.linear_f.bias)
    c = rrm(tape, c, G.RR_G.layers, G.RR_G.norm, h_)
    cond = cat_cols(tape, c, z)  # hier: the same 256-vector conditions every ccbn (model.py:471-473)
    # all ccbn gain/bias linears as ONE GEMM: (40E x 256) . (256 x sum 2C), per-column 1/sigma
    wp, wd, cs = sn.colscales[0]
    gbv = conv(tape, cond, None, n, 1, 1, 1, out_dtype=torch.float32, out_shape=(n, plan.gb_cols),
               grouped=(wp, wd, cs, plan.gb_cols, plan.gb_layers))
    gb, dgb = gbv.t, None
    if tape.record:
        dgb = torch.zeros_like(gb)  # the bn backward kernels write their (dgain, dbias) windows into it
        gbv.g = dgb
    bwid, hb = G.bottom_width, G.H_base
    hh, ww = bwid, bwid * hb
    lin = linear(tape, cond, h_[G.linear], bias=G.linear.bias)
    h = nchw_to_nhwc(tape, lin, n, G.arch["in_channels"][0], hh, ww)
    for bl in G.blocks:
        for blk in bl:
            h, hh, ww = _gblock(tape, blk, h, n, hh, ww, plan, gb, dgb, training)
    obn, oconv = G.output_layer[0], G.output_layer[2]
    ss = bn_affine(tape, h, n, hh, ww, gain=ptr(obn.gain), gain_ld=0, gain_add=0.0, bias=ptr(obn.bias), bias_ld=0,
                   stored_mean=obn.stored_mean, stored_var=obn.stored_var, training=training, eps=obn.eps,
                   momentum=obn.momentum, gain_param=obn.gain, bias_param=obn.bias)
    img = conv(tape, h, h_[oconv], n, hh, ww, 3, bias=oconv.bias, in_relu=True, ss=ss, act=L.ACT_TANH,
               out_dtype=torch.float32)
    return img, hh, ww


# ------------------------------------------------------------------ autograd bridge
class _NetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, body, n_in, *args):
        tape = Tape(True)
        in_vars = [Var(a, need=bool(a.requires_grad)) for a in args[:n_in]]
        outs = body(tape, *in_vars)
        ctx.tape, ctx.in_vars, ctx.outs, ctx.params, ctx.n_in = tape, in_vars, outs, args[n_in:], n_in
        res = tuple(o.t for o in outs)
        return res if len(res) > 1 else res[0]

    @staticmethod
    def backward(ctx, *gouts):
        tape = ctx.tape
        if tape is None:
            raise RuntimeError("iea_gan_b200: a net call can be back-propagated only once")
        for o, g in zip(ctx.outs, gouts):
            if g is not None:
                o.g = g.contiguous().to(o.t.dtype).view(o.t.shape)
        tape.backward()
        gin = [v.g.view(v.t.shape) if (v.need and v.g is not None and ctx.needs_input_grad[2 + i]) else None
               for i, v in enumerate(ctx.in_vars)]
        gp = [tape.pgrads.get(id(p)) if ctx.needs_input_grad[2 + ctx.n_in + j] else None
              for j, p in enumerate(ctx.params)]
        ctx.tape = None
        return (None, None) + tuple(gin) + tuple(gp)


def run_net(body, inputs, params):
    """Execute body(tape, *in_vars) -> [Var] with a tape (under autograd) or without one."""
    record = torch.is_grad_enabled() and (any(a.requires_grad for a in inputs) or any(p.requires_grad for p in params))
    if record:
        return _NetFn.apply(body, len(inputs), *inputs, *params)
    with torch.no_grad():
        outs = body(Tape(False), *[Var(a, need=False) for a in inputs])
    res = tuple(o.t for o in outs)
    return res if len(res) > 1 else res[0]


def _plain(t, dtype=torch.float32):
    if type(t) is not torch.Tensor:
        t = t.as_subclass(torch.Tensor)  # utils.Distribution instances flow in from train_fns.py
    return t.contiguous() if t.dtype == dtype else t.to(dtype).contiguous()


def generator_forward(G, z, y):
    z = _plain(z)
    L.require_device(z)
    n = z.shape[0]
    # same draw as model.py:466: torch.randn on the device generator before anything else (40E rows)
    rdof = torch.randn(n, G.rdof_dim, device=z.device)
    y = _plain(y, torch.int64)
    geo = {}

    def body(tape, zv):
        img, hh, ww = _g_body(G, tape, zv, y, rdof)
        geo["hw"] = (hh, ww)
        return [img]
    out = run_net(body, [z], list(G.parameters()))
    return out.view(n, 1, *geo["hw"])  # one channel: NHWC and NCHW coincide


def adu_postprocess(img):
    n, _, h, w = img.shape
    img = img.contiguous()
    out = torch.empty((n, h - 6, w), dtype=torch.float32, device=img.device)
    K("iea_adu_postprocess", ptr(img), n, h, w, ptr(out), L.stream())
    return out


def __getattr__(name):
    def _missing(*a, **k):
        raise NotImplementedError("engine.%s is not built yet" % name)
    return _missing
