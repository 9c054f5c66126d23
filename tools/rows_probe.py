"""HBM-bound row / layout kernels at the headline shapes, one launch sequence for an ncu capture and
CUDA-event GB/s for the same launches ("ncu evidence for every HBM kernel").

    ncu --set full --clock-control none -o gpurun_out/rows_r02 python tools/rows_probe.py --once
    python tools/rows_probe.py > gpurun_out/rows_r02_events.json

Algorithmic bytes = what the kernel must read + write once (stated per entry); achieved GB/s =
bytes / CUDA-event time (best of 5 with a 256 MB L2 flush in between), against the measured copy
peak in MEASURED_PEAKS.json.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import ops          # noqa: E402

BF16 = torch.bfloat16


def main():
    once = "--once" in sys.argv
    dev = "cuda"
    B, L, C, heads = 2, 50400, 5120, 40
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s, dt=BF16: torch.randn(*s, device=dev, dtype=torch.float32, generator=g).to(dt)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    cases = []

    # ---- DiT row kernels, 720p / 14B: x fp32 [2, 50400, 5120]
    x = rn(B, L, C, dt=torch.float32)
    em = rn(B, 6, C, dt=torch.float32)
    cases.append(("layernorm_modulate (AdaLN, fp32 in -> bf16 out) [2,50400,5120]", B * L * C * (4 + 2),
                  lambda: ops.layernorm_modulate(x, None, None, em[:, 0], em[:, 1], 6 * C, L, 1e-6)))
    w = rn(C); b_ = rn(C)
    cases.append(("layernorm_modulate (affine norm3) [2,50400,5120]", B * L * C * (4 + 2),
                  lambda: ops.layernorm_modulate(x, w, b_, eps=1e-6)))
    q = rn(B, L, C)
    from more4d_b200.dit import build_freqs
    fr = build_freqs(128)
    cos, sin = fr.real.float().contiguous().to(dev), fr.imag.float().contiguous().to(dev)
    grid = torch.tensor([[14, 45, 80]] * B, device=dev, dtype=torch.int32)
    cases.append(("rmsnorm_rope (WanRMSNorm + 3-axis RoPE, in place) [2,50400,5120]", B * L * C * (2 + 2),
                  lambda: ops.rmsnorm_rope_(q, w, heads, 1e-6, cos, sin, grid)))
    cases.append(("rmsnorm_rope (RoPE only, no norm weight) [2,50400,5120]", B * L * C * (2 + 2),
                  lambda: ops.rmsnorm_rope_(q, None, heads, 1e-6, cos, sin, grid)))
    lat = rn(1, 16, 13, 90, 160); nz = rn(2, 16, 13, 90, 160)
    cases.append(("cfg_euler_step [1,16,13,90,160]", lat.numel() * 2 * 4,
                  lambda: ops.cfg_euler_step_(lat, nz[0:1], nz[1:2], 6.0, -0.01)))
    xin, yin = rn(B, 16, 13, 90, 160), rn(B, 48, 13, 90, 160)
    cases.append(("patchify x|y -> rows [2,50400-ish,256]", (xin.numel() + yin.numel()) * 2 * 2,
                  lambda: ops.patchify(xin, yin)))
    tok = rn(B, L, 64)
    cases.append(("unpatchify [2,50400,64] -> [2,16,13,90,160]", B * (L - 3600) * 64 * 2 * 2,
                  lambda: ops.unpatchify(tok, 3600, 16, 13, 90, 160)))

    # ---- VAE row kernels: 4 frames of the 720p decoder tail / adaptor
    T, H, W = 4, 720, 1280
    a96 = rn(T, H, W, 96); gm = rn(96)
    cases.append(("rmsnorm_silu_cl [4,720,1280,96]", a96.numel() * 4,
                  lambda: ops.rmsnorm_silu_cl(a96, gm)))
    a128 = rn(T, H, W, 128); gw, gb = rn(128), rn(128)
    cases.append(("groupnorm_swish_cl (stats + apply = 2 launches; 2 reads + 1 write) [4,720,1280,128]",
                  a128.numel() * 2 * 3, lambda: ops.groupnorm_swish_cl(a128, gw, gb)))
    a192 = rn(T, 360, 640, 192)
    cases.append(("upsample2x_cl [4,360,640,192] -> [4,720,1280,192]", a192.numel() * 2 * 5,
                  lambda: ops.upsample2x_cl(a192)))
    pl = rn(3, T, H, W)
    cases.append(("planar_to_cl [3,4,720,1280] -> [...,16]", pl.numel() * 2 + T * H * W * 16 * 2,
                  lambda: ops.planar_to_cl(pl, 16)))
    enc_out = rn(13, 90, 160, 32)                     # where the VAE uses it: the encoder's 32-channel output
    cases.append(("cl_to_planar [13,90,160,32] -> [32,...] (encoder output)", enc_out.numel() * 4,
                  lambda: ops.cl_to_planar(enc_out)))
    s = rn(14400, 14400, dt=torch.float32)
    cases.append(("softmax_rows fp32 [14400,14400] -> bf16 (VAE attention)", s.numel() * 6,
                  lambda: ops.softmax_rows(s, 384 ** -0.5)))
    tr = rn(14400, 384)
    cases.append(("transpose_bf16 [14400,384]", tr.numel() * 4, lambda: ops.transpose_bf16(tr)))

    peak = 6454.0
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs", peak)
    out = []
    for name, nbytes, fn in cases:
        if once:
            fn()
            torch.cuda.synchronize()
            continue
        fn(); fn()
        best = 1e9
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        gbs = nbytes / best / 1e6
        out.append({"kernel": name, "algorithmic_bytes": nbytes, "ms": best, "achieved_gbs": gbs,
                    "frac_of_measured_hbm_peak": gbs / peak})
        print(f"{name}: {best:.3f} ms {gbs:.0f} GB/s ({gbs / peak:.2f})", file=sys.stderr, flush=True)
    if not once:
        print(json.dumps({"hbm_peak_gbs": peak, "kernels": out}, indent=1))


if __name__ == "__main__":
    main()
