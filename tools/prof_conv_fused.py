"""ncu target: the fused conv + residual + RMS_norm epilogue at the decoder's 96-channel 720p shape.
    ncu --set full --import-source on -k regex:conv_halo -s 2 -c 2 -o gpurun_out/prof_conv_fused python tools/prof_conv_fused.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import ops           # noqa: E402

BF16 = torch.bfloat16
if __name__ == "__main__":
    torch.set_grad_enabled(False)
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    H, W = (720, 1280) if C == 96 else (360, 640)
    x = torch.randn(T, H, W, C, device="cuda", dtype=BF16)
    res = torch.randn(T, H, W, C, device="cuda", dtype=BF16)
    w = torch.randn(C, C, 3, 3, 3, device="cuda", dtype=BF16) * 0.02
    b = torch.randn(C, device="cuda", dtype=BF16) * 0.1
    g = torch.ones(C, device="cuda", dtype=BF16)
    wp = ops.pack_conv_weight(w)
    for rep in range(2):
        ops.conv3x3_rmsnorm_cl(x, wp, b, C, 3, g, want_raw=False)                     # conv1 form
        ops.conv3x3_rmsnorm_cl(x, wp, b, C, 3, g, want_raw=True, residual=res)        # conv2 form
        torch.cuda.synchronize()
    # timing without the profiler
    for name, kw in (("plain", None), ("norm only", dict(want_raw=False)), ("raw+norm+residual", dict(want_raw=True, residual=res))):
        ts = []
        for _ in range(3):
            s, e = torch.cuda.Event(True), torch.cuda.Event(True)
            s.record()
            if kw is None:
                ops.conv_cl(x, wp, b, C, (3, 3, 3), pad=(2, 1, 1))
            else:
                ops.conv3x3_rmsnorm_cl(x, wp, b, C, 3, g, **kw)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        print(f"{name}: {min(ts):.3f} ms", flush=True)
