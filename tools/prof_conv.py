"""ncu target: a few launches of the convolution kernels at the VAE's dominant shapes.
    ncu --set full --clock-control none --import-source on -k regex:conv_ -s 4 -c 4 \
        -o gpurun_out/prof_conv python tools/prof_conv.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import _lib, ops           # noqa: E402

BF16 = torch.bfloat16
SHAPES = [(96, 96, 4, 720, 1280, 3), (192, 192, 4, 360, 640, 3), (384, 384, 4, 180, 320, 3), (128, 128, 4, 720, 1280, 1)]
if os.environ.get("PROF_CONV_SHAPES") == "thin":
    SHAPES = [(16, 128, 4, 720, 1280, 1), (16, 96, 4, 720, 1280, 3), (128, 3, 4, 720, 1280, 1), (96, 3, 4, 720, 1280, 3)]

if __name__ == "__main__":
    torch.set_grad_enabled(False)
    flags = int(sys.argv[1], 0) if len(sys.argv) > 1 else 0
    _lib.lib().m4d_set_debug_flags(flags)
    for rep in range(2):                    # first pass = warm-up launches (skipped with -s 4)
        for (cin, cout, T, H, W, kt) in SHAPES:
            x = torch.randn(T, H, W, cin, device="cuda", dtype=BF16)
            w = torch.randn(cout, cin, kt, 3, 3, device="cuda", dtype=BF16) * 0.02
            out = torch.empty(T, H, W, cout, device="cuda", dtype=BF16)
            ops.conv_cl(x, ops.pack_conv_weight(w, 16 if cin % 32 else 32), None, cout, (kt, 3, 3), pad=(kt - 1, 1, 1), out=out)
            torch.cuda.synchronize()
