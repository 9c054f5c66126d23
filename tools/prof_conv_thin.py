"""ncu / timing target: the output-bound thin-input convs of the adaptors (16 -> 128 channels, 3x3 per frame,
GroupNorm statistics in the epilogue) and the thin-output ones (128 -> 3 padded to 16).
    ncu --set full --import-source on -k regex:conv_halo -s 2 -c 1 -o gpurun_out/prof_thin python tools/prof_conv_thin.py 8"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import ops           # noqa: E402

BF16 = torch.bfloat16
if __name__ == "__main__":
    torch.set_grad_enabled(False)
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    H, W = 720, 1280
    x16 = torch.randn(T, H, W, 16, device="cuda", dtype=BF16)
    x128 = torch.randn(T, H, W, 128, device="cuda", dtype=BF16)
    w_in = ops.pack_conv_weight(torch.randn(128, 16, 3, 3, device="cuda", dtype=BF16) * 0.05, 16)
    w_mid = ops.pack_conv_weight(torch.randn(128, 128, 3, 3, device="cuda", dtype=BF16) * 0.02, 32)
    b = torch.randn(128, device="cuda", dtype=BF16) * 0.1
    cases = (("16->128 + stats", lambda: ops.conv3x3_gnstats_cl(x16, w_in, b, 128)),
             ("16->128 plain", lambda: ops.conv_cl(x16, w_in, b, 128, (1, 3, 3), pad=(0, 1, 1))),
             ("128->128 + stats", lambda: ops.conv3x3_gnstats_cl(x128, w_mid, b, 128)),
             ("128->128 plain", lambda: ops.conv_cl(x128, w_mid, b, 128, (1, 3, 3), pad=(0, 1, 1))),
             ("128->128 + stats + residual", lambda: ops.conv3x3_gnstats_cl(x128, w_mid, b, 128, residual=x128)))
    for name, fn in cases:
        fn(); fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            s, e = torch.cuda.Event(True), torch.cuda.Event(True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        out_gb = T * H * W * 128 * 2 / 1e9
        print(f"{name}: {min(ts):.3f} ms  (output {out_gb:.2f} GB -> {out_gb / min(ts) * 1e3:.0f} GB/s written)", flush=True)
