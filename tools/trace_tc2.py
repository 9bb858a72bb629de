"""Pipeline trace of CTA 0 of the resident tcgen05 conv kernel (needs a -DIEA_TC2_TRACE build).
slots: 0 producer item landed, 1 transformed, 2 arrived on full, 3 MMA saw full, 4 MMA committed,
5 epilogue starts waiting, 6 epilogue saw tfull, 7 epilogue released the accumulator."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iea_gan_b200 import engine as E_, _lib as L
import bench
r = bench.top_kernel_roofline(E_, 4, 6556.2, "measured")
print("RESULT", r["ms_per_launch"], r["achieved"])
buf = (ctypes.c_longlong * (8 * 512))()
lib = ctypes.CDLL(L.LIB_PATH)
fn = getattr(lib, "iea_debug_thin_trace", None) or lib.iea_debug_tc2_trace
print("rc", fn(buf))
t = [[buf[s * 512 + i] for i in range(512)] for s in range(8)]
t0 = min(x for x in t[0][:8] if x)
names = ["landed", "xformed", "arrived", "mma_full", "mma_commit", "epi_wait", "epi_full", "epi_done"]
print("tile " + " ".join("%10s" % n for n in names))
for i in list(range(0, 12)) + list(range(30, 60)):
    print("%4d " % i + " ".join("%10d" % (t[s][i] - t0) for s in range(8)))
for s in range(8):
    d = [t[s][i + 1] - t[s][i] for i in range(20, 60)]
    print(names[s], "mean delta per item", sum(d) / len(d))
