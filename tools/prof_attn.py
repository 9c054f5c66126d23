"""One attention launch sequence for an ncu capture of the shipped kernel (or, in a development
build, a variant: PROF_FLAGS=0x112 ...).  Shape: B=2, L=8192, 40 heads (2560 CTAs, ~17 waves).

    ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -o gpurun_out/attn_r02 python tools/prof_attn.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import _lib, ops          # noqa: E402

L = int(os.environ.get("ATTN_L", "8192"))
flags = int(os.environ.get("PROF_FLAGS", "0"), 0)
if flags:
    _lib.dev_set_flags(flags)
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(2, L, 40, 128, device="cuda", dtype=torch.bfloat16, generator=g) for _ in range(3))
for _ in range(3):
    o = ops.attention(q, k, v)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
