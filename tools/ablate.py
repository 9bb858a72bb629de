"""Ablation timing of the resident tcgen05 conv kernel (IEA_TC2_DBG bits: 1 no prologue transform,
2 no MMA issue, 4 no output stores, 8 no statistics, 16 no cp.async loads).  Profiling aid only."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from iea_gan_b200 import engine as E_
    import bench
    r = bench.top_kernel_roofline(E_, 4, 6556.2, "measured")
    print("RESULT", r["ms_per_launch"], r["achieved"])
else:
    for dbg in (0, 1, 2, 4, 8, 16, 1 | 8, 2 | 4 | 8, 1 | 2 | 4 | 8, 31):
        env = dict(os.environ, IEA_TC2_DBG=str(dbg))
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout
        print("dbg=%2d" % dbg, [l for l in out.splitlines() if l.startswith("RESULT")])
