"""Ablation timing of the resident tcgen05 conv kernel (IEA_TC2_DBG bits: 1 no prologue transform,
2 no MMA issue, 4 no output stores, 8 no statistics, 16 no cp.async loads; IEA_TC2_OCC=2 forces two
CTAs per SM on the 16-channel variants).  Profiling aid only: the switches exist in builds made with
`make -C iea_gan_b200/csrc NVCC="nvcc -DIEA_THIN_DBG"`; the product kernel has them compiled out.
usage: python tools/ablate.py [dbg[:occ] ...]"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from iea_gan_b200 import engine as E_
    import bench
    r = bench.top_kernel_roofline(E_, 4, 6556.2, "measured")
    print("RESULT", r["ms_per_launch"], r["achieved"])
else:
    cfgs = sys.argv[1:] or ["0", "1", "2", "4", "8", "16", "9", "14", "15", "31"]
    for c in cfgs:
        dbg, _, occ = c.partition(":")
        env = dict(os.environ, IEA_TC2_DBG=dbg)
        if occ:
            env["IEA_TC2_OCC"] = occ
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print("dbg=%s occ=%s" % (dbg, occ or "auto"), [l for l in out.stdout.splitlines() if l.startswith("RESULT")] or out.stderr[-400:])
