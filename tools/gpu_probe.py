"""First-contact diagnostics for the tcgen05 kernels on a real B200 (development tool).
Usage: python tools/gpu_probe.py {gemm|attn|rows}"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from more4d_b200 import _lib, ops           # noqa: E402
from oracle import dit_oracle as O          # noqa: E402

BF16 = torch.bfloat16


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(BF16)


def gemm():
    for (M, N, K) in [(128, 256, 64), (128, 256, 256), (300, 320, 192), (4096, 5120, 5120)]:
        a, w = rnd((M, K), 1), rnd((N, K), 2, 0.05)
        out = ops.linear(a.cuda(), w.cuda(), None)
        torch.cuda.synchronize()
        ref = (a.cuda().float() @ w.cuda().float().t())
        e = rel(out.float(), ref)
        print(f"gemm {M}x{N}x{K}: rel={e:.3e}", flush=True)
        if e > 1e-2 and M <= 300:
            d = (out.float() - ref).abs().cpu()
            print("  err by 32-col block:", [round(float(d[:, c:c + 32].mean()), 3) for c in range(0, N, 32)])
            print("  err by 32-row block:", [round(float(d[r:r + 32].mean()), 3) for r in range(0, M, 32)])
            # K-slice probe: which k-columns contribute correctly?
            for k0 in range(0, K, 16):
                a2 = torch.zeros_like(a); a2[:, k0:k0 + 16] = a[:, k0:k0 + 16]
                o2 = ops.linear(a2.cuda(), w.cuda(), None).float()
                r2 = a2.cuda().float() @ w.cuda().float().t()
                print(f"  k-slice {k0}: rel={rel(o2, r2):.3e}")
    # timing of the big one
    M, N, K = 8192, 5120, 5120
    a, w = rnd((M, K), 1).cuda(), rnd((N, K), 2, 0.05).cuda()
    out = torch.empty(M, N, device="cuda", dtype=BF16)
    for _ in range(3):
        ops.linear(a, w, None, out=out)
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(10):
        ops.linear(a, w, None, out=out)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"gemm {M}x{N}x{K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    s.record()
    for _ in range(10):
        torch.matmul(a, w.t(), out=out)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"cublas same shape: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)


def attn():
    ar = O.Arith(True)
    for flags in (0, 1, 2, 3):
        _lib.lib().m4d_set_debug_flags(flags)
        for (B, Lq, Lk, N) in [(1, 128, 128, 1), (1, 256, 384, 2), (2, 300, 257, 2)]:
            q, k, v = rnd((B, Lq, N, 128), 1), rnd((B, Lk, N, 128), 2), rnd((B, Lk, N, 128), 3)
            out = ops.attention(q.cuda(), k.cuda(), v.cuda())
            torch.cuda.synchronize()
            ref = O.attention(q, k, v, None, ar)
            print(f"attn flags={flags} B{B} Lq{Lq} Lk{Lk} N{N}: rel={rel(out.float().cpu(), ref):.3e} "
                  f"finite={bool(torch.isfinite(out).all())}", flush=True)
    _lib.lib().m4d_set_debug_flags(0)
    # structured probes: V = one-hot rows => O[q] = P[q, :] picks; uniform P (q=0) => O = mean(V)
    B, L, N = 1, 128, 1
    q = torch.zeros(B, L, N, 128, dtype=BF16)
    k = rnd((B, L, N, 128), 2)
    v = rnd((B, L, N, 128), 3)
    out = ops.attention(q.cuda(), k.cuda(), v.cuda()).float().cpu()
    print("uniform-P probe: rel vs mean(V) =", rel(out[0, :, 0], v[0, :, 0].float().mean(0, keepdim=True).expand(L, 128)))
    # timing
    for (B, L, N) in [(1, 8192, 12), (1, 21840, 12), (2, 50400, 40)]:
        q, k, v = (torch.randn(B, L, N, 128, device="cuda", dtype=BF16) for _ in range(3))
        out = torch.empty_like(q)
        ops.attention(q, k, v, out=out)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        ops.attention(q, k, v, out=out)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        fl = 4.0 * B * N * L * L * 128
        print(f"attn B{B} L{L} N{N}: {ms:.2f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
        if L <= 21840:
            with torch.nn.attention.sdpa_kernel(torch.nn.attention.SDPBackend.FLASH_ATTENTION):
                qq, kk, vv = (t.transpose(1, 2) for t in (q, k, v))
                o2 = torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
                torch.cuda.synchronize()
                s.record()
                o2 = torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
                e.record(); torch.cuda.synchronize()
                print(f"   torch SDPA(flash): {s.elapsed_time(e):.2f} ms  {fl/s.elapsed_time(e)/1e9:.1f} TFLOP/s;"
                      f" rel vs ours {rel(out.float(), o2.transpose(1, 2).float()):.3e}", flush=True)


def gemm_shapes():
    """The four GEMM shapes of a 14B block at 720p (M = 2*50400), CUDA-event timed."""
    M, C, Fd = 100800, 5120, 13824
    x = torch.randn(M, C, device="cuda", dtype=BF16)
    hbuf = torch.randn(M, Fd, device="cuda", dtype=BF16)
    res = torch.randn(M, C, device="cuda", dtype=torch.float32)
    gate = torch.randn(2, 6, C, device="cuda", dtype=torch.float32)
    wq = torch.randn(C, C, device="cuda", dtype=BF16) * 0.02
    w0 = torch.randn(Fd, C, device="cuda", dtype=BF16) * 0.02
    w2 = torch.randn(C, Fd, device="cuda", dtype=BF16) * 0.02
    b5, bF = torch.randn(C, device="cuda", dtype=BF16), torch.randn(Fd, device="cuda", dtype=BF16)
    outb = torch.empty(M, C, device="cuda", dtype=BF16)
    cases = [
        ("qkv   bias->bf16       K5120  N5120 ", lambda: ops.linear(x, wq, b5, out=outb), 2.0 * M * C * C),
        ("o     gate-residual    K5120  N5120 ", lambda: ops.linear(x, wq, b5, ops.EPI_GATE_RESIDUAL_F32, out=res, residual=res,
                                                                   gate=gate[:, 2], gate_batch_stride=6 * C, rows_per_batch=M // 2), 2.0 * M * C * C),
        ("ffn0  bias+gelu        K5120  N13824", lambda: ops.linear(x, w0, bF, ops.EPI_GELU_TANH, out=hbuf), 2.0 * M * C * Fd),
        ("ffn0' bias->bf16       K5120  N13824", lambda: ops.linear(x, w0, bF, out=hbuf), 2.0 * M * C * Fd),
        ("ffn2  gate-residual    K13824 N5120 ", lambda: ops.linear(hbuf, w2, b5, ops.EPI_GATE_RESIDUAL_F32, out=res, residual=res,
                                                                   gate=gate[:, 5], gate_batch_stride=6 * C, rows_per_batch=M // 2), 2.0 * M * C * Fd),
    ]
    for name, fn, fl in cases:
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            s_, e_ = torch.cuda.Event(True), torch.cuda.Event(True)
            s_.record(); fn(); e_.record(); torch.cuda.synchronize()
            ts.append(s_.elapsed_time(e_))
        ms = min(ts)
        print(f"{name}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    # correctness spot check of the rasterisation on a many-panel shape
    a = rnd((1000, 512), 1); w = rnd((13824, 512), 2, 0.05)
    o = ops.linear(a.cuda(), w.cuda(), None).float()
    ref = a.cuda().float() @ w.cuda().float().t()
    print("raster check rel:", rel(o, ref))


def gemm2():
    """2-CTA GEMM (debug flag 0x1000): correctness on small / ragged shapes, then timing."""
    _lib.lib().m4d_set_debug_flags(0x1000)
    for (M, N, K) in [(1024, 256, 64), (1024, 256, 512), (1300, 512, 192), (4096, 5120, 5120)]:
        a, w, b = rnd((M, K), 1), rnd((N, K), 2, 0.05), rnd((N,), 3, 0.1)
        out = ops.linear(a.cuda(), w.cuda(), b.cuda())
        torch.cuda.synchronize()
        ref = (a.cuda().float() @ w.cuda().float().t() + b.cuda().float())
        print(f"gemm2 {M}x{N}x{K}: rel={rel(out.float(), ref):.3e}", flush=True)
    for flags, name in ((0x2000, "1-CTA"), (0x1000, "2-CTA"), (0x2000, "1-CTA"), (0x1000, "2-CTA")):
        _lib.lib().m4d_set_debug_flags(flags)
        print("==", name, flush=True)
        gemm_shapes()
    _lib.lib().m4d_set_debug_flags(0)


def gemm2_k():
    """Per-k-block vs per-tile cost of the 1-CTA and 2-CTA GEMM kernels."""
    M, N = 16384, 5120
    for K in (512, 2048, 8192):
        a = torch.randn(M, K, device="cuda", dtype=BF16)
        w = torch.randn(N, K, device="cuda", dtype=BF16) * 0.02
        out = torch.empty(M, N, device="cuda", dtype=BF16)
        for flags, name in ((0x2000, "1-CTA"), (0x1000, "2-CTA")):
            _lib.lib().m4d_set_debug_flags(flags)
            for _ in range(2):
                ops.linear(a, w, None, out=out)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                s_, e_ = torch.cuda.Event(True), torch.cuda.Event(True)
                s_.record(); ops.linear(a, w, None, out=out); e_.record(); torch.cuda.synchronize()
                ts.append(s_.elapsed_time(e_))
            ms = min(ts)
            print(f"K={K} {name}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    _lib.lib().m4d_set_debug_flags(0)


def attn_sweep():
    """Polynomial-exp2 share sweep (debug flags 0x10 | PP) at the bench shape + parity."""
    ar = O.Arith(True)
    q, k, v = rnd((1, 300, 2, 128), 1, 2.0), rnd((1, 300, 2, 128), 2, 2.0), rnd((1, 300, 2, 128), 3)
    ref = O.attention(q, k, v, None, ar)
    B, L, N = 2, 50400, 40
    Q, K, V = (torch.randn(B, L, N, 128, device="cuda", dtype=BF16) for _ in range(3))
    out = torch.empty_like(Q)
    fl = 4.0 * B * N * L * L * 128
    variants = [(0, 0, 0, 0), (0, 0, 0, 1), (0, 0, 0, 0), (0, 0, 0, 1)]     # (PP, VAR, PACE, split)
    if len(sys.argv) > 2:
        variants = [tuple(int(x) for x in a.split(",")) for a in sys.argv[2:]]
    for pp, var, pace, split in variants:
        if split:
            _lib.lib().m4d_set_debug_flags(0x4000000 | 0x2000000)
        else:
            _lib.lib().m4d_set_debug_flags(0x4000000 | 0x200 | 0x100 | (pp << 4) | var | 0x1000000 | (pace << 20))
        o = ops.attention(q.cuda(), k.cuda(), v.cuda())
        e = rel(o.float().cpu(), ref)
        ops.attention(Q, K, V, out=out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            s_, e_ = torch.cuda.Event(True), torch.cuda.Event(True)
            s_.record()
            ops.attention(Q, K, V, out=out)
            e_.record(); torch.cuda.synchronize()
            ts.append(s_.elapsed_time(e_))
        ms = min(ts)
        print(f"PP={pp} VAR={var} PACE={pace} split={split}: parity rel={e:.3e}  {ms:.2f} ms  {fl/ms/1e9:.1f} TFLOP/s (all: {[round(t,1) for t in ts]})", flush=True)
    _lib.lib().m4d_set_debug_flags(0)


def conv():
    """Halo-staging conv kernel vs the per-tap kernel (debug flag 0x10000) and the oracle; the
    descriptor base-offset variant (0x20000); then timing at the VAE's dominant shapes."""
    from oracle import vae_oracle as V
    ar = V.Arith(True)
    L = _lib.lib()

    def cl(x):
        return x[0].permute(1, 2, 3, 0).contiguous().cuda()

    def ncthw(y):
        return y.permute(3, 0, 1, 2).unsqueeze(0).float().cpu()

    for (cin, cout, T, H, W, kt) in [(64, 64, 1, 16, 16, 1), (64, 64, 2, 20, 40, 3), (96, 96, 3, 12, 20, 3),
                                     (32, 192, 2, 8, 16, 3), (192, 384, 2, 9, 17, 3), (128, 128, 2, 33, 47, 1)]:
        x = rnd((1, cin, T, H, W), 1)
        w = rnd((cout, cin, kt, 3, 3), 2, (cin * 9 * kt) ** -0.5)
        b = rnd((cout,), 3, 0.1)
        if kt == 3:
            ref = V.causal_conv3d(x.float(), w, b, ar)
        else:
            ref = V.conv2d_frames(x.float(), w[:, :, 0], b, ar, stride=1, pad=(1, 1, 1, 1))
        wp = ops.pack_conv_weight(w.cuda())
        for flags, name in ((0x10000, "per-tap"), (0, "halo"), (0x20000, "halo+base_offset")):
            L.m4d_set_debug_flags(flags)
            y = ops.conv_cl(cl(x), wp, b.cuda(), cout, (kt, 3, 3), pad=(kt - 1, 1, 1))
            torch.cuda.synchronize()
            print(f"conv {cin}->{cout} T{T} {H}x{W} kt{kt} [{name}]: rel={rel(ncthw(y), ref):.3e}", flush=True)
    L.m4d_set_debug_flags(0)
    for (cin, cout, T, H, W) in [(96, 96, 8, 720, 1280), (192, 192, 8, 360, 640), (384, 384, 8, 180, 320),
                                 (384, 384, 13, 90, 160), (128, 128, 8, 720, 1280)]:
        kt = 1 if cin == 128 else 3
        x = torch.randn(T, H, W, cin, device="cuda", dtype=BF16)
        w = torch.randn(cout, cin, kt, 3, 3, device="cuda", dtype=BF16) * 0.02
        wp = ops.pack_conv_weight(w)
        out = torch.empty(T, H, W, cout, device="cuda", dtype=BF16)
        fl = 2.0 * T * H * W * cin * cout * 9 * kt
        for flags, name in ((0x10000, "per-tap"), (0, "halo"), (0x20000, "halo+bo")):
            L.m4d_set_debug_flags(flags)
            for _ in range(2):
                ops.conv_cl(x, wp, None, cout, (kt, 3, 3), pad=(kt - 1, 1, 1), out=out)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                s_, e_ = torch.cuda.Event(True), torch.cuda.Event(True)
                s_.record(); ops.conv_cl(x, wp, None, cout, (kt, 3, 3), pad=(kt - 1, 1, 1), out=out); e_.record()
                torch.cuda.synchronize()
                ts.append(s_.elapsed_time(e_))
            ms = min(ts)
            print(f"conv {cin}->{cout} T{T} {H}x{W} kt{kt} [{name}]: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    L.m4d_set_debug_flags(0)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    {"gemm": gemm, "attn": attn, "attn_sweep": attn_sweep, "gemm_shapes": gemm_shapes, "gemm2": gemm2, "gemm2_k": gemm2_k, "conv": conv}[sys.argv[1]]()
