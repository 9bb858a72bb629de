"""Ablation timing of one conv layer on the macro-tile tcgen05 kernel (needs a -DIEA_THIN_DBG build).
usage: python tools/ablate_layer.py cin cout k hw events [res] [dbg ...]"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if sys.argv[1] == "child":
    import torch
    from iea_gan_b200 import engine as E_, _lib as L
    import iea_gan_b200.sn_layers as SL
    cin, cout, k, hw, ev, res = (int(a) for a in sys.argv[2:8])
    n, h, w = 40 * ev, hw, hw
    m = SL.SNConv2d(cin, cout, k, padding=k // 2, eps=1e-6).cuda()
    grp = E_.SNGroup(); l = grp.add(m, E_.act_dtype()); grp.run(True, False)
    x = torch.randn(n, h, w, cin, device="cuda").bfloat16()
    ss = E_.ScaleShift(torch.rand(n, cin, device="cuda") + 0.5, torch.randn(n, cin, device="cuda"))
    rv = E_.Var(torch.randn(n, h // 2, w // 2, 2 * cout, device="cuda").bfloat16()) if res else None
    tape = E_.Tape(False)
    fn = lambda: E_.conv(tape, E_.Var(x, need=False), l, n, h, w, k, bias=m.bias, in_relu=True, ss=ss, stats=True,
                         res=rv, res_mode=L.IN_UP2 if res else 0, res_c=cout if res else 0)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): fn()
    b.record(); torch.cuda.synchronize()
    print("RESULT %.4f ms" % (a.elapsed_time(b) / 5))
else:
    shape = sys.argv[1:7]
    for dbg in sys.argv[7:] or ["0", "1", "2", "4", "8", "16", "31"]:
        out = subprocess.run([sys.executable, __file__, "child"] + shape, env=dict(os.environ, IEA_TC2_DBG=dbg), capture_output=True, text=True)
        print("dbg=%s" % dbg, [l for l in out.stdout.splitlines() if l.startswith("RESULT")] or out.stderr[-300:])
