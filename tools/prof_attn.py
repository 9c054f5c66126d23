"""ncu target: two launches of the self-attention kernel (first = warm-up).
    ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 \
        -o gpurun_out/prof_attn python tools/prof_attn.py [L] [heads] [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import _lib, ops           # noqa: E402

if __name__ == "__main__":
    torch.set_grad_enabled(False)
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    flags = int(os.environ.get("M4D_DEBUG_FLAGS", "0"), 0)
    _lib.lib().m4d_set_debug_flags(flags)
    q, k, v = (torch.randn(B, L, N, 128, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    out = torch.empty_like(q)
    for _ in range(2):
        ops.attention(q, k, v, out=out)
        torch.cuda.synchronize()
