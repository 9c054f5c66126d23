"""Tensor-level wrappers over the C ABI: validate, pass data_ptr()s + strides, launch on the
current torch stream.  No arithmetic happens in Python/torch here; torch only owns memory."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import (EPI_ADD_BF16, EPI_BF16, EPI_F32, EPI_F32_RAW,  # noqa: F401
                   EPI_GATE_RESIDUAL_F32, EPI_GELU_ERF, EPI_GELU_TANH)

Tensor = torch.Tensor
BF16 = torch.bfloat16


class _Stats:
    """Launch accounting for bench.py: every C-ABI compute call is exactly one kernel launch."""
    launches = 0
    timed = None          # None, or {"attention": [(start_evt, end_evt, flops), ...]}


def launches() -> int:
    return _Stats.launches


def start_kernel_timing(classes=("attention",)) -> None:
    """Bracket the launches of the given kernel classes with CUDA events on the launching stream:
    "attention" (self-attention, Lq == Lk), "cross_attention", "gemm", "rows" (LayerNorm / RMSNorm /
    RoPE passes).  bench.py uses "attention" inside the timed region (40 launches per forward) and
    all classes in a separate, untimed profiling pass."""
    _Stats.timed = {c: [] for c in classes}


def stop_kernel_timing():
    t, _Stats.timed = _Stats.timed, None
    return t


def _t0(cls: str):
    if _Stats.timed is None or cls not in _Stats.timed:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _t1(cls: str, ev0, work: float) -> None:
    if ev0 is not None:
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        _Stats.timed[cls].append((ev0, ev1, work))


def _stream() -> int:
    _Stats.launches += 1
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"more4d_b200: `{name}` must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"more4d_b200: `{name}` must be {dtype}, got {t.dtype}")


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, epilogue: int = EPI_BF16,
           out: Optional[Tensor] = None, residual: Optional[Tensor] = None,
           gate: Optional[Tensor] = None, gate_batch_stride: int = 0,
           rows_per_batch: int = 0) -> Tensor:
    """out = epilogue(x @ weight.T + bias).  x: [..., K] bf16 (last dim contiguous, uniform row
    stride); weight: [N, K] bf16 (an nn.Linear weight, consumed in place)."""
    _lib.require_device()
    _req(x, BF16, "x")
    _req(weight, BF16, "weight")
    K = x.shape[-1]
    N = weight.shape[0]
    if weight.dim() != 2:
        weight = weight.reshape(N, -1)
    if weight.shape[1] != K:
        raise ValueError(f"more4d_b200.linear: K mismatch {weight.shape[1]} vs {K}")
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    if weight.stride(-1) != 1:
        raise ValueError("more4d_b200.linear: weight rows must be contiguous")
    M = x2.shape[0]
    f32_out = epilogue in (EPI_F32, EPI_GATE_RESIDUAL_F32, EPI_F32_RAW)
    if out is None:
        out = torch.empty(*x.shape[:-1], N, device=x.device, dtype=torch.float32 if f32_out else BF16)
    _req(out, torch.float32 if f32_out else BF16, "out")
    out2 = out.reshape(-1, N) if out.is_contiguous() else out
    if out2.dim() != 2 or out2.shape[0] != M or out2.stride(-1) != 1:
        raise ValueError("more4d_b200.linear: `out` must be viewable as [M, N] with unit col stride")
    res2 = None
    if epilogue == EPI_GATE_RESIDUAL_F32:
        if residual is None:
            raise ValueError("more4d_b200.linear: residual required")
        _req(residual, torch.float32, "residual")
        res2 = residual.reshape(-1, N)
        if gate is not None:
            _req(gate, torch.float32, "gate")
    elif epilogue == EPI_ADD_BF16:
        if residual is None:
            raise ValueError("more4d_b200.linear: residual required")
        _req(residual, BF16, "residual")
        res2 = residual.reshape(-1, N)
    if bias is not None:
        _req(bias, BF16, "bias")
    ev = _t0("gemm")
    rc = _lib.lib().m4d_gemm_bf16(
        x2.data_ptr(), x2.stride(0), weight.data_ptr(), weight.stride(0), _ptr(bias),
        out2.data_ptr(), out2.stride(0), M, N, K, epilogue,
        _ptr(res2), 0 if res2 is None else res2.stride(0), _ptr(gate), gate_batch_stride,
        rows_per_batch, _stream())
    _lib.check(rc, "m4d_gemm_bf16")
    _t1("gemm", ev, 2.0 * M * N * K)
    return out


def attention(q: Tensor, k: Tensor, v: Tensor, k_lens: Optional[Tensor] = None,
              softmax_scale: Optional[float] = None, out: Optional[Tensor] = None,
              accumulate: bool = False) -> Tensor:
    """softmax(q k^T * scale) v.  q: [B, Lq, N, 128], k/v: [B, Lk, N, 128] bf16; the head and
    channel dims must be contiguous, batch/token strides are free.  k_lens: int32 CUDA [B]."""
    _lib.require_device()
    for n, t in (("q", q), ("k", k), ("v", v)):
        _req(t, BF16, n)
        if t.dim() != 4 or t.stride(3) != 1 or t.stride(2) != t.shape[3]:
            raise ValueError(f"more4d_b200.attention: `{n}` must be [B, L, N, D] with contiguous (N, D)")
    B, Lq, N, D = q.shape
    Lk = k.shape[1]
    if k.shape != v.shape or k.shape[0] != B or k.shape[2] != N or k.shape[3] != D:
        raise ValueError("more4d_b200.attention: q/k/v shape mismatch")
    if k.stride() != v.stride():
        v = v.contiguous()
        k = k.contiguous()
    if out is None:
        out = torch.empty(B, Lq, N, D, device=q.device, dtype=BF16)
    _req(out, BF16, "out")
    if k_lens is not None:
        _req(k_lens, torch.int32, "k_lens")
    cls = "attention" if Lq == Lk else "cross_attention"
    ev = _t0(cls)
    rc = _lib.lib().m4d_attention_fwd(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, Lq, Lk, N, D,
        q.stride(0), q.stride(1), k.stride(0), k.stride(1), out.stride(0), out.stride(1),
        _ptr(k_lens), float(softmax_scale) if softmax_scale else 0.0, int(accumulate), _stream())
    _lib.check(rc, "m4d_attention_fwd")
    _t1(cls, ev, 4.0 * B * N * Lq * Lk * D)
    return out


def attention_seg2(q: Tensor, k: Tensor, v: Tensor, seg_len: int,
                   softmax_scale: Optional[float] = None, out: Optional[Tensor] = None) -> Tensor:
    """attention(q, k[:, :seg_len], v[:, :seg_len]) + attention(q, k[:, seg_len:], v[:, seg_len:]) (each
    rounded to bf16, summed in bf16) in ONE launch — the image+text cross-attention, t4d:533-552.
    seg_len: positive multiple of 128, < Lk."""
    _lib.require_device()
    for n, t in (("q", q), ("k", k), ("v", v)):
        _req(t, BF16, n)
        if t.dim() != 4 or t.stride(3) != 1 or t.stride(2) != t.shape[3]:
            raise ValueError(f"more4d_b200.attention_seg2: `{n}` must be [B, L, N, D] with contiguous (N, D)")
    B, Lq, N, D = q.shape
    Lk = k.shape[1]
    if k.shape != v.shape or k.stride() != v.stride() or k.shape[0] != B or k.shape[2] != N or k.shape[3] != D:
        raise ValueError("more4d_b200.attention_seg2: q/k/v shape or stride mismatch")
    if seg_len <= 0 or seg_len % 128 or seg_len >= Lk:
        raise ValueError("more4d_b200.attention_seg2: seg_len must be a positive multiple of 128 below Lk")
    if out is None:
        out = torch.empty(B, Lq, N, D, device=q.device, dtype=BF16)
    _req(out, BF16, "out")
    ev = _t0("cross_attention")
    rc = _lib.lib().m4d_attention_fwd_seg2(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, Lq, Lk, seg_len, N, D,
        q.stride(0), q.stride(1), k.stride(0), k.stride(1), out.stride(0), out.stride(1),
        float(softmax_scale) if softmax_scale else 0.0, _stream())
    _lib.check(rc, "m4d_attention_fwd_seg2")
    _t1("cross_attention", ev, 4.0 * B * N * Lq * Lk * D)
    return out


def attention_scatter(q: Tensor, k: Tensor, v: Tensor, outs, k_lens: Optional[Tensor] = None) -> None:
    """attention(q, k, v) with query rows scattered over `outs`: rows [i*R, (i+1)*R) go to outs[i]
    ([B, R, N, 128] views with equal strides — typically peer GPUs' buffers), one launch."""
    import ctypes
    _lib.require_device()
    for n, t in (("q", q), ("k", k), ("v", v)):
        _req(t, BF16, n)
        if t.dim() != 4 or t.stride(3) != 1 or t.stride(2) != t.shape[3]:
            raise ValueError(f"more4d_b200.attention_scatter: `{n}` must be [B, L, N, D] with contiguous (N, D)")
    B, Lq, N, D = q.shape
    Lk = k.shape[1]
    if k.stride() != v.stride():
        raise ValueError("more4d_b200.attention_scatter: k and v must share strides")
    R = outs[0].shape[1]
    for o in outs:
        _req(o, BF16, "out")
        if o.shape != (B, R, N, D) or o.stride() != outs[0].stride() or o.stride(3) != 1 or o.stride(2) != D:
            raise ValueError("more4d_b200.attention_scatter: outs must be equally strided [B, R, N, D] views")
    if len(outs) * R < Lq:
        raise ValueError("more4d_b200.attention_scatter: outs do not cover the query rows")
    ptrs = (ctypes.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
    if k_lens is not None:
        _req(k_lens, torch.int32, "k_lens")
    ev = _t0("attention")
    rc = _lib.lib().m4d_attention_fwd_scatter(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), ptrs, len(outs), R, B, Lq, Lk, N, D, q.stride(0), q.stride(1),
        k.stride(0), k.stride(1), outs[0].stride(0), outs[0].stride(1), _ptr(k_lens), 0.0, _stream())
    _lib.check(rc, "m4d_attention_fwd_scatter")
    _t1("attention", ev, 4.0 * B * N * Lq * Lk * D)


def layernorm_modulate(x: Tensor, weight: Optional[Tensor] = None, bias: Optional[Tensor] = None,
                       shift: Optional[Tensor] = None, scale: Optional[Tensor] = None,
                       mod_batch_stride: int = 0, rows_per_batch: Optional[int] = None,
                       eps: float = 1e-6, out_dtype=BF16, guidance: Optional[Tensor] = None,
                       guidance_gate: Optional[Tensor] = None) -> Tensor:
    """LayerNorm over the last dim (+affine) (+AdaLN `*(1+scale)+shift`) in one pass.
    shift/scale: fp32 views whose element [b, :] sits at data_ptr + b*mod_batch_stride."""
    _lib.require_device()
    if not x.is_contiguous():
        x = x.contiguous()
    C = x.shape[-1]
    rows = x.numel() // C
    if x.dtype not in (torch.float32, BF16):
        raise TypeError("more4d_b200.layernorm_modulate: x must be fp32 or bf16")
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    sg_rows, sg_stride = 0, 0
    if guidance is not None:
        _req(guidance, BF16, "guidance")
        guidance = guidance.contiguous()
        n_batches = rows // (rows_per_batch or rows)
        if guidance.dim() != 3 or guidance.shape[0] != n_batches or guidance.shape[2] != 2 * C:
            raise ValueError(f"more4d_b200.layernorm_modulate: guidance must be [{n_batches}, rows, {2 * C}] "
                             f"(one slab per batch element), got {tuple(guidance.shape)}")
        sg_rows = guidance.shape[1]
        sg_stride = guidance.stride(0)
    ev = _t0("rows")
    rc = _lib.lib().m4d_layernorm_modulate(
        x.data_ptr(), int(x.dtype == BF16), _ptr(weight), _ptr(bias), _ptr(shift), _ptr(scale),
        mod_batch_stride, rows, rows_per_batch or rows, C, eps, out.data_ptr(),
        int(out_dtype == torch.float32), _ptr(guidance), sg_stride, sg_rows, _ptr(guidance_gate),
        _stream())
    _lib.check(rc, "m4d_layernorm_modulate")
    _t1("rows", ev, float(x.numel() * x.element_size() + out.numel() * out.element_size()))
    return out


def rmsnorm_rope_(x: Tensor, weight: Optional[Tensor], heads: int, eps: float = 1e-6,
                  rope_cos: Optional[Tensor] = None, rope_sin: Optional[Tensor] = None,
                  grid_fhw: Optional[Tensor] = None) -> Tensor:
    """In place: WanRMSNorm over all channels, then 3-axis RoPE.  x: [B, L, heads*D] bf16."""
    _lib.require_device()
    _req(x, BF16, "x")
    B, L, C = x.shape
    if x.stride(2) != 1 or x.stride(0) != L * x.stride(1):
        raise ValueError("more4d_b200.rmsnorm_rope_: x rows must be uniformly strided")
    if rope_cos is not None:
        _req(rope_cos, torch.float32, "rope_cos")
        _req(grid_fhw, torch.int32, "grid_fhw")
    ev = _t0("rows")
    rc = _lib.lib().m4d_rmsnorm_rope(x.data_ptr(), x.stride(1), _ptr(weight), _ptr(rope_cos),
                                     _ptr(rope_sin), _ptr(grid_fhw), B, L, heads, C // heads, eps,
                                     _stream())
    _lib.check(rc, "m4d_rmsnorm_rope")
    _t1("rows", ev, 4.0 * B * L * C)
    return x


def rmsnorm_scatter(x: Tensor, weight: Optional[Tensor], dst, dst_row0: int, eps: float = 1e-6) -> None:
    """WanRMSNorm of x [B, L_local, C] bf16, scattered by head group: dst[g] ([B, L, C/P] bf16, this
    or a peer GPU's memory) receives channels [g*C/P, (g+1)*C/P) at rows dst_row0 + l."""
    import ctypes
    _lib.require_device()
    _req(x, BF16, "x")
    B, Ll, C = x.shape
    P = len(dst)
    if x.stride(2) != 1 or x.stride(0) != Ll * x.stride(1):
        raise ValueError("more4d_b200.rmsnorm_scatter: x rows must be uniformly strided")
    for d in dst:
        _req(d, BF16, "dst")
        if d.dim() != 3 or d.shape[0] != B or d.shape[2] != C // P or not d.is_contiguous():
            raise ValueError("more4d_b200.rmsnorm_scatter: dst[g] must be contiguous [B, L, C/P]")
    ptrs = (ctypes.c_void_p * P)(*[d.data_ptr() for d in dst])
    rc = _lib.lib().m4d_rmsnorm_scatter(x.data_ptr(), x.stride(1), _ptr(weight), B, Ll, C, eps, ptrs, P,
                                        dst[0].stride(0), dst_row0, _stream())
    _lib.check(rc, "m4d_rmsnorm_scatter")


def small_linear_f32(x: Tensor, weight: Tensor, bias: Optional[Tensor], silu_in: bool = False,
                     silu_out: bool = False) -> Tensor:
    _lib.require_device()
    _req(x, torch.float32, "x")
    _req(weight, BF16, "weight")
    x = x.contiguous()
    M, K = x.shape
    N = weight.shape[0]
    y = torch.empty(M, N, device=x.device, dtype=torch.float32)
    rc = _lib.lib().m4d_small_linear_f32(x.data_ptr(), weight.data_ptr(), _ptr(bias), y.data_ptr(),
                                         M, N, K, int(silu_in), int(silu_out), _stream())
    _lib.check(rc, "m4d_small_linear_f32")
    return y


def timestep_embedding(t: Tensor, dim: int) -> Tensor:
    _lib.require_device()
    t = t.to(device="cuda", dtype=torch.float32).contiguous()
    out = torch.empty(t.shape[0], dim, device=t.device, dtype=torch.float32)
    rc = _lib.lib().m4d_timestep_embedding(t.data_ptr(), t.shape[0], dim, out.data_ptr(), _stream())
    _lib.check(rc, "m4d_timestep_embedding")
    return out


def add_bcast(a_bf16: Tensor, e: Tensor) -> Tensor:
    """out[b, i] = a[i] + e[b, i % m]  (a: bf16 with n = r*m elements, e: fp32 [B, m]) -> fp32
    [B, n]."""
    _lib.require_device()
    _req(a_bf16, BF16, "a")
    _req(e, torch.float32, "e")
    e = e.contiguous()
    B = e.shape[0]
    m = e.numel() // B
    n = a_bf16.numel()
    if n % m != 0:
        raise ValueError("more4d_b200.add_bcast: size mismatch")
    out = torch.empty(B, n, device=e.device, dtype=torch.float32)
    rc = _lib.lib().m4d_add_bcast_f32(a_bf16.data_ptr(), e.data_ptr(), out.data_ptr(), B, n, m,
                                      _stream())
    _lib.check(rc, "m4d_add_bcast_f32")
    return out


def patchify(x: Tensor, y: Optional[Tensor] = None) -> Tensor:
    """[B, Cx, T, H, W] (+ [B, Cy, T, H, W]) -> im2col rows [B, T*(H/2)*(W/2), (Cx+Cy)*4]."""
    _lib.require_device()
    _req(x, BF16, "x")
    x = x.contiguous()
    B, Cx, T, H, W = x.shape
    Cy = 0
    if y is not None:
        _req(y, BF16, "y")
        y = y.contiguous()
        Cy = y.shape[1]
    out = torch.empty(B, T * (H // 2) * (W // 2), (Cx + Cy) * 4, device=x.device, dtype=BF16)
    rc = _lib.lib().m4d_patchify(x.data_ptr(), _ptr(y), B, Cx, Cy, T, H, W, out.data_ptr(), _stream())
    _lib.check(rc, "m4d_patchify")
    return out


def unpatchify(tokens: Tensor, skip_tokens: int, cout: int, T: int, H: int, W: int) -> Tensor:
    """tokens [B, L, 4*cout] bf16 -> [B, cout, T, H, W] (H, W = latent size)."""
    _lib.require_device()
    _req(tokens, BF16, "tokens")
    tokens = tokens.contiguous()
    B = tokens.shape[0]
    out = torch.empty(B, cout, T, H, W, device=tokens.device, dtype=BF16)
    rc = _lib.lib().m4d_unpatchify(tokens.data_ptr(), tokens.stride(0), skip_tokens, B, cout, T, H,
                                   W, out.data_ptr(), _stream())
    _lib.check(rc, "m4d_unpatchify")
    return out


def widen_rows(src: Tensor, dst: Tensor, dst_row0: int) -> None:
    """dst[b, dst_row0:dst_row0+rows, :] (fp32) = src[b, :, :] (bf16)."""
    _lib.require_device()
    _req(src, BF16, "src")
    _req(dst, torch.float32, "dst")
    src = src.contiguous()
    B, rows, C = src.shape
    rc = _lib.lib().m4d_widen_rows(src.data_ptr(), dst.data_ptr(), B, rows, C, dst.stride(0),
                                   dst_row0, _stream())
    _lib.check(rc, "m4d_widen_rows")


def cfg_euler_step_(latents: Tensor, uncond: Tensor, text: Tensor, guidance: float, dt: float) -> Tensor:
    _lib.require_device()
    for n, t in (("latents", latents), ("uncond", uncond), ("text", text)):
        _req(t, BF16, n)
        if not t.is_contiguous():
            raise ValueError(f"more4d_b200.cfg_euler_step_: `{n}` must be contiguous")
    rc = _lib.lib().m4d_cfg_euler_step(uncond.data_ptr(), text.data_ptr(), latents.data_ptr(),
                                       guidance, dt, latents.numel(), _stream())
    _lib.check(rc, "m4d_cfg_euler_step")
    return latents


def silu_bf16(x: Tensor) -> Tensor:
    """bf16(SiLU(x)) for fp32 x."""
    _lib.require_device()
    _req(x, torch.float32, "x")
    x = x.contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=BF16)
    rc = _lib.lib().m4d_silu_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    _lib.check(rc, "m4d_silu_bf16")
    return out


def im2col3x3_cl(x: Tensor, silu: bool = False) -> Tensor:
    """x [F, H, W, C] bf16 channels-last -> im2col rows [F*H*W, 9*C] of a 3x3 / padding-1
    convolution (tap-major, matching `pack_conv_weight`), optionally of SiLU(x)."""
    _lib.require_device()
    _req(x, BF16, "x")
    if not x.is_contiguous() or x.dim() != 4:
        raise ValueError("more4d_b200.im2col3x3_cl: x must be contiguous [F, H, W, C]")
    F_, H, W, C = x.shape
    rows = torch.empty(F_ * H * W, 9 * C, device=x.device, dtype=BF16)
    rc = _lib.lib().m4d_im2col3x3_cl(x.data_ptr(), rows.data_ptr(), F_, H, W, C, int(silu), _stream())
    _lib.check(rc, "m4d_im2col3x3_cl")
    return rows


def bilinear_repeat_cl(x: Tensor, T: int, H: int, W: int, want_raw: bool = True, want_silu: bool = True):
    """x [B, h, w, C] bf16 -> bilinear (align_corners=False) resize to (H, W), repeated over T
    frames: returns (raw, silu), each [B, T*H*W, C] bf16 or None."""
    _lib.require_device()
    _req(x, BF16, "x")
    if not x.is_contiguous() or x.dim() != 4:
        raise ValueError("more4d_b200.bilinear_repeat_cl: x must be contiguous [B, h, w, C]")
    B, h, w, C = x.shape
    raw = torch.empty(B, T * H * W, C, device=x.device, dtype=BF16) if want_raw else None
    act = torch.empty(B, T * H * W, C, device=x.device, dtype=BF16) if want_silu else None
    rc = _lib.lib().m4d_bilinear_repeat_cl(x.data_ptr(), _ptr(raw), _ptr(act), B, h, w, C, T, H, W, _stream())
    _lib.check(rc, "m4d_bilinear_repeat_cl")
    return raw, act


# ======================================================================================
# Motion-Sensitive VAE ops (channels-last bf16 activations [T, H, W, C])
# ======================================================================================
def pack_conv_weight(w: Tensor, cin_multiple: int = 32) -> Tensor:
    """Reference conv weight [Cout, Cin, (kt,) kh, kw] -> implicit-GEMM operand
    [Cout_pad16, taps * Cin_pad] (tap-major, channel-minor, zero padded; Cin padded to
    `cin_multiple`: 32, or 16 for the thin 3-channel inputs of the 3x3 halo kernel).  One-time
    weight preprocessing, not on the per-call path."""
    if w.dim() == 4:
        w = w.unsqueeze(2)
    cout, cin, kt, kh, kw = w.shape
    cin_p, cout_p = (cin + cin_multiple - 1) // cin_multiple * cin_multiple, (cout + 15) // 16 * 16
    p = torch.zeros(cout_p, kt, kh, kw, cin_p, device=w.device, dtype=BF16)
    p[:cout, :, :, :, :cin] = w.permute(0, 2, 3, 4, 1).to(BF16)
    return p.reshape(cout_p, kt * kh * kw * cin_p).contiguous()


def conv_cl(x: Tensor, w_packed: Tensor, bias: Optional[Tensor], cout: int, kernel, stride=(1, 1, 1),
            pad=(0, 0, 0), t_out: Optional[int] = None, out: Optional[Tensor] = None,
            t_mul: int = 1, t_off: int = 0, n_split: Optional[int] = None,
            residual: Optional[Tensor] = None, planar_out: Optional[Tensor] = None, act: int = 0,
            skip: Optional[Tensor] = None) -> Tensor:
    """Implicit-GEMM convolution over a channels-last sequence x [T, H, W, Cin] (Cin % 32 == 0;
    Cin % 16 == 0 for 3x3 stride-1 convolutions).
    `pad` = (front frames, top/left rows, cols); right/bottom/behind padding is implied by the
    output size.  Returns the channels-last output (or `planar_out` [Cout, T, H, W])."""
    _lib.require_device()
    _req(x, BF16, "x")
    if not x.is_contiguous():
        raise ValueError("more4d_b200.conv_cl: x must be contiguous [T, H, W, C]")
    T, H, W, Cin = x.shape
    kt, kh, kw = kernel
    st, sh, sw = stride
    pt, ph, pw = pad
    if w_packed.shape[1] != kt * kh * kw * Cin:
        raise ValueError("more4d_b200.conv_cl: packed weight does not match (taps, Cin)")
    H_out = (H + (2 * ph if sh == 1 else 1) - kh) // sh + 1
    W_out = (W + (2 * pw if sw == 1 else 1) - kw) // sw + 1
    if t_out is None:
        t_out = (T + pt - kt) // st + 1
    out_mode = 0
    if planar_out is not None:
        _req(planar_out, BF16, "planar_out")
        out_t, out_mode, out_C = planar_out, 1, cout
        n_split = n_split or max(cout, 32)
    else:
        if out is None:
            out = torch.empty(t_out, H_out, W_out, cout, device=x.device, dtype=BF16)
        _req(out, BF16, "out")
        out_t, out_C = out, out.shape[-1]
        n_split = n_split or max(out_C, cout)
    rc = _lib.lib().m4d_conv_cl(
        x.data_ptr(), T, H, W, Cin, w_packed.data_ptr(), cout, w_packed.shape[0], _ptr(bias),
        kt, kh, kw, st, sh, sw, pt, ph, pw, t_out, H_out, W_out, out_t.data_ptr(), out_C, t_mul, t_off,
        n_split, _ptr(residual), out_mode, act, _ptr(skip), _stream())
    _lib.check(rc, "m4d_conv_cl")
    return out_t


def conv3x3_rmsnorm_cl(x: Tensor, w_packed: Tensor, bias: Optional[Tensor], cout: int, kt: int,
                       gamma: Tensor, silu: bool = True, want_raw: bool = True,
                       residual: Optional[Tensor] = None):
    """3x3 (x kt causal) stride-1 conv whose epilogue also emits RMS_norm[+SiLU] of its output
    (the next layer's first op).  Returns (raw or None, normalised), both [T, H, W, cout]."""
    _lib.require_device()
    _req(x, BF16, "x")
    if not x.is_contiguous():
        raise ValueError("more4d_b200.conv3x3_rmsnorm_cl: x must be contiguous [T, H, W, C]")
    T, H, W, Cin = x.shape
    if w_packed.shape != (cout, kt * 9 * Cin):
        raise ValueError("more4d_b200.conv3x3_rmsnorm_cl: packed weight does not match (cout, taps, Cin)")
    raw = torch.empty(T, H, W, cout, device=x.device, dtype=BF16) if want_raw else None
    normed = torch.empty(T, H, W, cout, device=x.device, dtype=BF16)
    rc = _lib.lib().m4d_conv3x3_rmsnorm_cl(
        x.data_ptr(), T, H, W, Cin, w_packed.data_ptr(), cout, _ptr(bias), kt, _ptr(raw), _ptr(residual),
        gamma.data_ptr(), normed.data_ptr(), int(silu), _stream())
    _lib.check(rc, "m4d_conv3x3_rmsnorm_cl")
    return raw, normed


FUSED_NORM_CHANNELS = (96, 192)     # Cout values m4d_conv3x3_rmsnorm_cl supports


def rmsnorm_silu_cl(x: Tensor, gamma: Tensor, silu: bool = True, inplace: bool = False) -> Tensor:
    _lib.require_device()
    _req(x, BF16, "x")
    C = x.shape[-1]
    out = x if inplace else torch.empty_like(x)
    rc = _lib.lib().m4d_rmsnorm_silu_cl(x.data_ptr(), gamma.data_ptr(), out.data_ptr(), x.numel() // C,
                                        C, int(silu), _stream())
    _lib.check(rc, "m4d_rmsnorm_silu_cl")
    return out


def upsample2x_cl(x: Tensor) -> Tensor:
    _lib.require_device()
    _req(x, BF16, "x")
    T, H, W, C = x.shape
    out = torch.empty(T, 2 * H, 2 * W, C, device=x.device, dtype=BF16)
    rc = _lib.lib().m4d_upsample2x_cl(x.data_ptr(), out.data_ptr(), T, H, W, C, _stream())
    _lib.check(rc, "m4d_upsample2x_cl")
    return out


def planar_to_cl(x: Tensor, cpad: int, div: Optional[Tensor] = None, add: Optional[Tensor] = None) -> Tensor:
    """[C, T, H, W] -> [T, H, W, cpad] (zero-padded channels), optional per-channel x/div + add."""
    _lib.require_device()
    _req(x, BF16, "x")
    x = x.contiguous()
    C, T, H, W = x.shape
    out = torch.empty(T, H, W, cpad, device=x.device, dtype=BF16)
    rc = _lib.lib().m4d_planar_to_cl(x.data_ptr(), out.data_ptr(), T * H * W, C, cpad, _ptr(div), _ptr(add),
                                     _stream())
    _lib.check(rc, "m4d_planar_to_cl")
    return out


def cl_to_planar(x: Tensor, n_affine: int = 0, sub: Optional[Tensor] = None,
                 mul: Optional[Tensor] = None) -> Tensor:
    """[T, H, W, C] -> [C, T, H, W]; the first n_affine channels get (v - sub) * mul."""
    _lib.require_device()
    _req(x, BF16, "x")
    T, H, W, C = x.shape
    out = torch.empty(C, T, H, W, device=x.device, dtype=BF16)
    rc = _lib.lib().m4d_cl_to_planar(x.data_ptr(), out.data_ptr(), T * H * W, C, n_affine, _ptr(sub),
                                     _ptr(mul), _stream())
    _lib.check(rc, "m4d_cl_to_planar")
    return out


GN_SLICES = 8                          # M4D_GN_SLICES of include/more4d_b200.h


def conv3x3_gnstats_cl(x: Tensor, w_packed: Tensor, bias: Optional[Tensor], cout: int,
                       residual: Optional[Tensor] = None):
    """3x3 Conv2d per frame ([T, H, W, Cin] -> [T, H, W, 128]) whose epilogue also leaves the
    GroupNorm(32) statistics of its output (`stats` for groupnorm_swish_cl): returns (out, stats)."""
    _lib.require_device()
    _req(x, BF16, "x")
    if not x.is_contiguous():
        raise ValueError("more4d_b200.conv3x3_gnstats_cl: x must be contiguous [T, H, W, C]")
    T, H, W, Cin = x.shape
    if w_packed.shape != (cout, 9 * Cin):
        raise ValueError("more4d_b200.conv3x3_gnstats_cl: packed weight does not match (cout, 9 * Cin)")
    out = torch.empty(T, H, W, cout, device=x.device, dtype=BF16)
    tiles = ((H + 15) // 16) * ((W + 15) // 16)
    ws = torch.empty(T * tiles * 64, device=x.device, dtype=torch.float32)
    stats = torch.empty(T * GN_SLICES * 64, device=x.device, dtype=torch.float32)
    _Stats.launches += 1                                   # conv + the partial-sum reduction
    rc = _lib.lib().m4d_conv3x3_gnstats_cl(x.data_ptr(), T, H, W, Cin, w_packed.data_ptr(), cout, _ptr(bias),
                                           out.data_ptr(), _ptr(residual), ws.data_ptr(), stats.data_ptr(),
                                           _stream())
    _lib.check(rc, "m4d_conv3x3_gnstats_cl")
    return out, stats


def groupnorm_swish_cl(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-6, groups: int = 32,
                       inplace: bool = False, stats: Optional[Tensor] = None) -> Tensor:
    """GroupNorm + swish on [F, H, W, C] with per-frame statistics; `stats` (from
    conv3x3_gnstats_cl, the producer of x) skips the statistics pass."""
    _lib.require_device()
    _req(x, BF16, "x")
    F_, H, W, C = x.shape
    out = x if inplace else torch.empty_like(x)
    if stats is not None:
        _req(stats, torch.float32, "stats")
        if stats.numel() != F_ * GN_SLICES * 64:
            raise ValueError("more4d_b200.groupnorm_swish_cl: stats do not belong to this tensor")
        rc = _lib.lib().m4d_groupnorm_apply_cl(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                               stats.data_ptr(), GN_SLICES, F_, H * W, C, groups, eps, _stream())
        _lib.check(rc, "m4d_groupnorm_apply_cl")
        return out
    ws = torch.empty(64 * F_, device=x.device, dtype=torch.float32)
    _Stats.launches += 1                                   # stats + apply = two kernels
    rc = _lib.lib().m4d_groupnorm_swish_cl(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                           ws.data_ptr(), F_, H * W, C, groups, eps, _stream())
    _lib.check(rc, "m4d_groupnorm_swish_cl")
    return out


def softmax_rows(s: Tensor, scale: float, out: Optional[Tensor] = None) -> Tensor:
    """p = softmax(s * scale) row-wise; `out` may be a wider (zero-initialised) bf16 buffer view."""
    _lib.require_device()
    _req(s, torch.float32, "s")
    R, N = s.shape
    p = torch.empty(R, N, device=s.device, dtype=BF16) if out is None else out
    _req(p, BF16, "out")
    if p.shape != (R, N) or p.stride(1) != 1 or s.stride(1) != 1:
        raise ValueError("more4d_b200.softmax_rows: shape / stride mismatch")
    rc = _lib.lib().m4d_softmax_rows(s.data_ptr(), p.data_ptr(), R, N, s.stride(0), p.stride(0), scale,
                                     _stream())
    _lib.check(rc, "m4d_softmax_rows")
    return p


def transpose_bf16(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """[R, C] (row stride free) -> [C, R] (contiguous, or the given view with a free row stride)."""
    _lib.require_device()
    _req(x, BF16, "x")
    R, C = x.shape
    if out is None:
        out = torch.empty(C, R, device=x.device, dtype=BF16)
    _req(out, BF16, "out")
    if out.shape != (C, R) or out.stride(1) != 1:
        raise ValueError("more4d_b200.transpose_bf16: `out` must be a [C, R] view with unit column stride")
    rc = _lib.lib().m4d_transpose_bf16(x.data_ptr(), out.data_ptr(), R, C, x.stride(0), out.stride(0), _stream())
    _lib.check(rc, "m4d_transpose_bf16")
    return out
