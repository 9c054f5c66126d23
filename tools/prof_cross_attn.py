"""ncu / timing target: the cross-attention launch of the headline workload (B 2, Lq 50400, 40 heads, text 512 +
image 257 keys in one launch, m4d_attention_fwd_seg2).
    ncu --set full --import-source on -k regex:attn_fwd -s 2 -c 1 -o gpurun_out/prof_xattn python tools/prof_cross_attn.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import ops           # noqa: E402

if __name__ == "__main__":
    torch.set_grad_enabled(False)
    B, L, N, Lk, seg = 2, 50400, 40, 769, 512
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(B, L, N, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Lk, N, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Lk, N, 128, device="cuda", generator=g).bfloat16()
    out = torch.empty_like(q)
    for _ in range(3):
        ops.attention_seg2(q, k, v, seg, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); ops.attention_seg2(q, k, v, seg, out=out); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    fl = 4.0 * B * N * L * Lk * 128
    print(f"seg2 cross-attention: {min(ts):.3f} ms  {fl / min(ts) / 1e9:.0f} TFLOP/s", flush=True)
    # the same keys as ONE segment (no mid epilogue, one exact tile): what the segment boundary costs
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); ops.attention(q, k, v, out=out); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    print(f"one segment, 769 keys: {min(ts):.3f} ms  {fl / min(ts) / 1e9:.0f} TFLOP/s", flush=True)
    k8 = k[:, :768].contiguous(); v8 = v[:, :768].contiguous()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); ops.attention(q, k8, v8, out=out); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    print(f"one segment, 768 keys (6 tiles): {min(ts):.3f} ms  {4.0 * B * N * L * 768 * 128 / min(ts) / 1e9:.0f} TFLOP/s", flush=True)
