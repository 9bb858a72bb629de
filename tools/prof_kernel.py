"""One of bench.py's ROOFLINE_LAYERS launched a few times on its own (for ncu: the LAST launch is the warm one).
usage: python tools/prof_kernel.py <tag> [launches]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from iea_gan_b200 import engine as E_

# extra layers for kernel work (not part of bench.py's roofline candidates)
bench.ROOFLINE_LAYERS.update({
    "l128_k3_16": (16, 16, 128, 128, 3, 0, True, "iea_conv_fprop 128->128 3x3 @16x16 (tc::conv_tc_kernel)"),
    "l128_k3_8": (8, 8, 128, 128, 3, 0, True, "iea_conv_fprop 128->128 3x3 @8x8 (tc::conv_tc_kernel)"),
})
tag = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
call, bytes_, extra, label = bench.roofline_layer(E_, tag)
for _ in range(reps):
    call()
torch.cuda.synchronize()
print(tag, label, bytes_, extra)
