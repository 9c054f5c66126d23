"""Host-side mirror of the reference's Wan2.1-DiT (4D-STraG denoiser) over the sm_100a kernels.

Class names, constructor arguments, forward signatures and state-dict keys follow
MoRe4D/models/wan_transformer4d.py ("t4d") so that reference checkpoints load with
``load_state_dict`` and reference call sites (pipeline_wan_fun_control.py:796-817,
train_wan.py:1940-1950) work unchanged.  The nn.Modules below only *hold* parameters; all
arithmetic is in libmore4d_sm100.so (more4d_b200/ops.py).  Inference only: under
``torch.is_grad_enabled()`` the forward raises (backward is out of scope).

Data flow of one block (t4d:633-688), 9 GEMMs + 3 attention launches + 6 row kernels:
  e  = modulation + e0                                   add_bcast          fp32 [B,6,C]
  t  = LN(x)*(1+e1)+e0 -> bf16                           layernorm_modulate
  q,k,v = t Wq^T, t Wk^T, t Wv^T (+bias) -> bf16         gemm x3
  q,k = RoPE(RMSNorm(.))  in place                       rmsnorm_rope x2
  o  = softmax(q k^T/sqrt(d)) v                          attention (tcgen05/TMEM)
  x += bf16(o Wo^T + b) * e2           fp32 in place     gemm, gate-residual epilogue
  n  = LN_affine(x) -> bf16                              layernorm_modulate
  x += bf16(cross_attn(n, context))                      gemm x(2+4), rmsnorm x3, attention x2
  t  = LN(x)*(1+e4)+e3 -> bf16 ; h = GELU(t W0^T + b)    layernorm_modulate, gemm (GELU epilogue)
  x += bf16(h W2^T + b) * e5                             gemm, gate-residual epilogue
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from .cache_utils import teacache_decide, teacache_step_done
from .config import DiTConfig

Tensor = torch.Tensor
BF16 = torch.bfloat16


def _no_grad_only(what: str) -> None:
    if torch.is_grad_enabled():
        raise RuntimeError(f"more4d_b200.{what}: forward-only kernels; wrap the call in "
                           "torch.no_grad() (the reference's training path is out of scope)")


class _Param(nn.Module):
    """Parameter container with nn.Linear / nn.LayerNorm / conv key names (weight, bias)."""

    def __init__(self, wshape, bias=True, device=None, dtype=BF16):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*wshape, device=device, dtype=dtype),
                                   requires_grad=False)
        self.bias = nn.Parameter(torch.empty(wshape[0], device=device, dtype=dtype),
                                 requires_grad=False) if bias else None


def _seq(*mods) -> nn.Sequential:
    """Sequential whose integer keys match the reference (e.g. ffn.0 / ffn.2)."""
    return nn.Sequential(*mods)


class _Slot(nn.Module):
    """Parameter-less placeholder (activation) that keeps Sequential indices aligned."""


# --------------------------------------------------------------------------------------
def rope_params(max_seq_len: int, dim: int, theta: float = 10000.0) -> Tensor:
    """Complex128 table [max_seq_len, dim/2] — same definition as t4d:252-260."""
    ang = torch.outer(torch.arange(max_seq_len, dtype=torch.float64),
                      1.0 / torch.pow(theta, torch.arange(0, dim, 2, dtype=torch.float64) / dim))
    return torch.polar(torch.ones_like(ang), ang)


def rope_params_riflex(max_seq_len: int, dim: int, k: Optional[int] = None, L_test: Optional[int] = None,
                       L_test_scale: Optional[float] = None, theta: float = 10000.0) -> Tensor:
    """RIFLEx frame-axis table (t4d:264-320, `get_1d_rotary_pos_embed_riflex(..., use_real=False)`):
    the k-th intrinsic frequency is lowered to 0.9 * 2 pi / L_test (/ L_test_scale) so that the
    extrapolated clip stays inside one period.  Complex128 [max_seq_len, dim/2]."""
    freqs = 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim))
    if k is not None:
        freqs[k - 1] = 0.9 * 2 * torch.pi / L_test
    if L_test_scale is not None:
        freqs[k - 1] = freqs[k - 1] / L_test_scale
    ang = torch.outer(torch.arange(max_seq_len), freqs)
    return torch.polar(torch.ones_like(ang), ang)


def build_freqs(head_dim: int) -> Tensor:
    """The model's `freqs` buffer (t4d:928-935): frame | row | col tables along the pair axis."""
    d = head_dim
    return torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                      rope_params(1024, 2 * (d // 6))], dim=1)


class _RopeCache:
    """fp32 cos/sin device tables derived from a complex `freqs` tensor."""

    def __init__(self):
        self.key = None
        self.cos = self.sin = None

    def get(self, freqs: Tensor, device):
        # the content tag guards against a re-allocated table (enable_riflex) landing on the same address
        key = (freqs.data_ptr(), tuple(freqs.shape), str(device), float(freqs.real[-1].sum()))
        if key != self.key:
            f = freqs.to("cpu")
            self.cos = f.real.to(torch.float32).contiguous().to(device)
            self.sin = f.imag.to(torch.float32).contiguous().to(device)
            self.key = key
        return self.cos, self.sin


_rope_cache = _RopeCache()


def rope_apply_qk(q: Tensor, k: Tensor, grid_sizes: Tensor, freqs: Tensor):
    """Seam-compatible `rope_apply_qk(q, k, grid_sizes, freqs)` (t4d:372-375): q, k
    [B, L, N, D] bf16, rotated in place and returned."""
    _no_grad_only("rope_apply_qk")
    cos, sin = _rope_cache.get(freqs, q.device)
    grid = torch.as_tensor(grid_sizes).to(device=q.device, dtype=torch.int32).contiguous()
    B, L, N, D = q.shape
    for t in (q, k):
        ops.rmsnorm_rope_(t.view(B, L, N * D), None, N, 0.0, cos, sin, grid)
    return q, k


def attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None,
              causal=False, window_size=(-1, -1), deterministic=False, dtype=BF16,
              fa_version=None, attention_type=None):
    """Seam-compatible `attention()` (t4d:175-236), layout [B, L, N, D].  Only the configuration
    the 4D-STraG path uses is implemented; anything else raises instead of falling back."""
    _no_grad_only("attention")
    if causal or dropout_p != 0.0 or tuple(window_size) != (-1, -1) or q_lens is not None:
        raise NotImplementedError("more4d_b200.attention: only non-causal, full-window, "
                                  "dropout-free attention without q_lens is on the hot path")
    if q_scale is not None:
        raise NotImplementedError("more4d_b200.attention: q_scale is not used by the hot path")
    q, k, v = (t if t.dtype == BF16 else t.to(BF16) for t in (q, k, v))
    kl = None
    if k_lens is not None:
        kl = torch.as_tensor(k_lens).to(device=q.device, dtype=torch.int32)
    return ops.attention(q, k, v, kl, softmax_scale)


# --------------------------------------------------------------------------------------
class WanRMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-5, device=None, dtype=BF16):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.empty(dim, device=device, dtype=dtype), requires_grad=False)


class WanLayerNorm(nn.Module):
    def __init__(self, dim, eps=1e-6, elementwise_affine=False, device=None, dtype=BF16):
        super().__init__()
        self.dim, self.eps = dim, eps
        if elementwise_affine:
            self.weight = nn.Parameter(torch.empty(dim, device=device, dtype=dtype), requires_grad=False)
            self.bias = nn.Parameter(torch.empty(dim, device=device, dtype=dtype), requires_grad=False)
        else:
            self.weight = self.bias = None


class WanSelfAttention(nn.Module):
    """t4d:409-466."""

    def __init__(self, dim, num_heads, window_size=(-1, -1), qk_norm=True, eps=1e-6, device=None):
        super().__init__()
        assert dim % num_heads == 0
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.window_size, self.qk_norm, self.eps = window_size, qk_norm, eps
        self.q = _Param((dim, dim), device=device)
        self.k = _Param((dim, dim), device=device)
        self.v = _Param((dim, dim), device=device)
        self.o = _Param((dim, dim), device=device)
        self.norm_q = WanRMSNorm(dim, eps, device=device) if qk_norm else None
        self.norm_k = WanRMSNorm(dim, eps, device=device) if qk_norm else None

    def attend_sp(self, x: Tensor, k_lens: Optional[Tensor], grid_i32: Tensor, cos: Tensor, sin: Tensor,
                  sp) -> Tensor:
        """Ulysses form of `attend` (the slot of the reference's usp_attn_forward, t4d:1042-1044):
        x is this rank's token shard [B, L/P, C].  q/k/v projections and the full-channel RMSNorm
        are token-local; ONE all-to-all hands every rank all L tokens of heads/P heads, where RoPE
        (needs global token positions) and attention run; one all-to-all brings the output back."""
        B, n_loc, C = x.shape
        n, d = self.num_heads, self.head_dim
        if sp.peer_memory and x.device.type != "cpu":          # gloo / CPU tests take the collective form
            return self._attend_sp_peer(x, k_lens, grid_i32, cos, sin, sp)
        qkv = torch.empty(3, B, n_loc, C, device=x.device, dtype=BF16)
        for i, lin in enumerate((self.q, self.k, self.v)):
            ops.linear(x, lin.weight, lin.bias, out=qkv[i])
        if self.qk_norm:
            ops.rmsnorm_rope_(qkv[0], self.norm_q.weight, n, self.eps)
            ops.rmsnorm_rope_(qkv[1], self.norm_k.weight, n, self.eps)
        full = sp.seq_to_heads(qkv.view(3, B, n_loc, n, d))                 # [3, B, L, n/P, d]
        h = n // sp.world
        L = full.shape[2]
        q, k, v = (full[i].reshape(B, L, h * d) for i in range(3))
        ops.rmsnorm_rope_(q, None, h, self.eps, cos, sin, grid_i32)
        ops.rmsnorm_rope_(k, None, h, self.eps, cos, sin, grid_i32)
        o = ops.attention(q.view(B, L, h, d), k.view(B, L, h, d), v.view(B, L, h, d), k_lens)
        return sp.heads_to_seq(o).reshape(B, n_loc, C)

    def _attend_sp_peer(self, x: Tensor, k_lens: Optional[Tensor], grid_i32: Tensor, cos: Tensor,
                        sin: Tensor, sp) -> Tensor:
        """The same exchange with NO pass of its own: producers store straight into the consumer
        rank's buffer over NVLink (dist.PeerBuffers) —
          v        : the projection GEMM's epilogue writes head group g into rank g's buffer,
          q, k     : the RMSNorm kernel (full-channel statistics need the whole local row, so the
                     GEMM stays local) scatters its output by head group (m4d_rmsnorm_scatter),
          attention: ONE launch whose epilogue routes every query row to the rank that owns its token
                     chunk, into that rank's `o` buffer at this rank's head columns
                     (m4d_attention_fwd_scatter) —
        and each all-to-all shrinks to a ~7 us barrier.  Bit-identical to the unsharded forward."""
        B, n_loc, C = x.shape
        n, d, P, r = self.num_heads, self.head_dim, sp.world, sp.rank
        h, gc = n // P, C // P
        L = n_loc * P
        pb = sp.peer_buffers(B, L, C, x.device)
        rows = slice(r * n_loc, (r + 1) * n_loc)
        for g in range(P):
            wv, bv = self.v.weight[g * gc:(g + 1) * gc], self.v.bias[g * gc:(g + 1) * gc]
            for b in range(B):
                ops.linear(x[b], wv, bv, out=pb.qkv_peers[g][2, b, rows])
        q = ops.linear(x, self.q.weight, self.q.bias)
        k = ops.linear(x, self.k.weight, self.k.bias)
        ops.rmsnorm_scatter(q, self.norm_q.weight if self.qk_norm else None, [p_[0] for p_ in pb.qkv_peers],
                            r * n_loc, self.eps)
        ops.rmsnorm_scatter(k, self.norm_k.weight if self.qk_norm else None, [p_[1] for p_ in pb.qkv_peers],
                            r * n_loc, self.eps)
        pb.barrier()                                   # every rank's pieces have landed in pb.qkv
        qf, kf, vf = pb.qkv[0], pb.qkv[1], pb.qkv[2]   # [B, L, h*d]: all tokens, my heads
        ops.rmsnorm_rope_(qf, None, h, self.eps, cos, sin, grid_i32)
        ops.rmsnorm_rope_(kf, None, h, self.eps, cos, sin, grid_i32)
        k4, v4 = kf.view(B, L, h, d), vf.view(B, L, h, d)
        # ONE launch over all L queries; the epilogue stores token chunk s into rank s's buffer
        outs = [pb.o_peers[s_].view(B, n_loc, n, d)[:, :, r * h:(r + 1) * h] for s_ in range(P)]
        ops.attention_scatter(qf.view(B, L, h, d), k4, v4, outs, k_lens)
        pb.barrier()                                   # every rank's head columns have landed in pb.o
        return pb.o

    def attend(self, x: Tensor, k_lens: Optional[Tensor], grid_i32: Tensor, cos: Tensor,
               sin: Tensor, sp=None) -> Tensor:
        """x bf16 [B, L, C] -> attention output before the `o` projection, bf16 [B, L, C]."""
        if sp is not None and sp.world > 1:
            return self.attend_sp(x, k_lens, grid_i32, cos, sin, sp)
        B, L, C = x.shape
        n, d = self.num_heads, self.head_dim
        q = ops.linear(x, self.q.weight, self.q.bias)
        k = ops.linear(x, self.k.weight, self.k.bias)
        v = ops.linear(x, self.v.weight, self.v.bias)
        ops.rmsnorm_rope_(q, self.norm_q.weight if self.qk_norm else None, n, self.eps, cos, sin, grid_i32)
        ops.rmsnorm_rope_(k, self.norm_k.weight if self.qk_norm else None, n, self.eps, cos, sin, grid_i32)
        o = ops.attention(q.view(B, L, n, d), k.view(B, L, n, d), v.view(B, L, n, d), k_lens)
        return o.view(B, L, C)

    def forward(self, x, seq_lens, grid_sizes, freqs, dtype=BF16, t=0):
        _no_grad_only("WanSelfAttention")
        x = x.to(BF16)
        cos, sin = _rope_cache.get(freqs, x.device)
        grid = torch.as_tensor(grid_sizes).to(device=x.device, dtype=torch.int32).contiguous()
        kl = torch.as_tensor(seq_lens).to(device=x.device, dtype=torch.int32)
        o = self.attend(x, kl, grid, cos, sin)
        return ops.linear(o, self.o.weight, self.o.bias)


class WanI2VCrossAttention(WanSelfAttention):
    """t4d:499-554 (image + text context, two attentions summed); without the *_img
    parameters it degenerates to WanT2VCrossAttention t4d:469-497."""

    def __init__(self, dim, num_heads, window_size=(-1, -1), qk_norm=True, eps=1e-6, device=None,
                 image_branch=True, clip_tokens=257):
        super().__init__(dim, num_heads, window_size, qk_norm, eps, device=device)
        self.image_branch = image_branch
        self.clip_tokens = clip_tokens
        if image_branch:
            self.k_img = _Param((dim, dim), device=device)
            self.v_img = _Param((dim, dim), device=device)
            self.norm_k_img = WanRMSNorm(dim, eps, device=device) if qk_norm else None

    def project_context(self, context: Tensor):
        """K/V of the text (and image) context, t4d:481-483,527-531 — independent of x and of the
        timestep, i.e. identical on all 50 steps of the loop.  Returns (k, v, k_img, v_img)."""
        n = self.num_heads
        ctx_txt = context[:, self.clip_tokens:].contiguous() if self.image_branch else context
        k = ops.linear(ctx_txt, self.k.weight, self.k.bias)
        ops.rmsnorm_rope_(k, self.norm_k.weight if self.qk_norm else None, n, self.eps)
        v = ops.linear(ctx_txt, self.v.weight, self.v.bias)
        ki = vi = None
        if self.image_branch:
            ctx_img = context[:, :self.clip_tokens].contiguous()
            ki = ops.linear(ctx_img, self.k_img.weight, self.k_img.bias)
            ops.rmsnorm_rope_(ki, self.norm_k_img.weight if self.qk_norm else None, n, self.eps)
            vi = ops.linear(ctx_img, self.v_img.weight, self.v_img.bias)
            if k.shape[1] % 128 == 0:
                # text keys then image keys in ONE buffer (views returned), so `attend` can run both
                # softmaxes in one launch (ops.attention_seg2)
                Lt = k.shape[1]
                kc, vc = torch.cat([k, ki], dim=1), torch.cat([v, vi], dim=1)
                k, ki, v, vi = kc[:, :Lt], kc[:, Lt:], vc[:, :Lt], vc[:, Lt:]
        return k, v, ki, vi

    @staticmethod
    def _adjacent(a: Tensor, b: Tensor) -> bool:
        """b's rows start where a's end, inside one [B, La + Lb, C] buffer."""
        return (a.stride() == b.stride() and a.stride(1) == a.shape[2] and a.shape[1] % 128 == 0 and
                b.data_ptr() == a.data_ptr() + a.shape[1] * a.stride(1) * a.element_size() and
                (a.shape[0] == 1 or a.stride(0) == (a.shape[1] + b.shape[1]) * a.stride(1)))

    def attend(self, x: Tensor, context: Optional[Tensor], kv=None) -> Tensor:   # type: ignore[override]
        B, L, C = x.shape
        n, d = self.num_heads, self.head_dim
        q = ops.linear(x, self.q.weight, self.q.bias)
        ops.rmsnorm_rope_(q, self.norm_q.weight if self.qk_norm else None, n, self.eps)
        k, v, ki, vi = kv if kv is not None else self.project_context(context)
        q4 = q.view(B, L, n, d)
        if ki is not None and self._adjacent(k, ki) and self._adjacent(v, vi):
            Lt, Lc = k.shape[1], k.shape[1] + ki.shape[1]
            kc = k.as_strided((B, Lc, n, d), (k.stride(0), k.stride(1), d, 1))
            vc = v.as_strided((B, Lc, n, d), (v.stride(0), v.stride(1), d, 1))
            return ops.attention_seg2(q4, kc, vc, Lt).view(B, L, C)
        o = ops.attention(q4, k.view(B, -1, n, d), v.view(B, -1, n, d))
        if ki is not None:
            ops.attention(q4, ki.view(B, -1, n, d), vi.view(B, -1, n, d), out=o, accumulate=True)
        return o.view(B, L, C)

    def forward(self, x, context, context_lens=None, dtype=BF16, t=0):   # type: ignore[override]
        _no_grad_only("WanI2VCrossAttention")
        if context_lens is not None:
            raise NotImplementedError("context_lens is always None on the reference path (t4d:1174)")
        o = self.attend(x.to(BF16), context.to(BF16))
        return ops.linear(o, self.o.weight, self.o.bias)


class SpatialGuidanceModule(nn.Module):
    """Motion-Perception-Module injection, t4d:739-783 (parameters only; fused into the
    AdaLN row kernel)."""

    def __init__(self, dim, dino_feature_dim=768, device=None):
        super().__init__()
        self.dim = dim
        self.spatial_guide = _seq(_Slot(), _Param((dim * 2, dino_feature_dim), device=device))
        self.gate = nn.Parameter(torch.empty(dim, device=device, dtype=BF16), requires_grad=False)

    def project(self, feats_silu: Tensor) -> Tensor:
        """feats_silu: bf16 SiLU(features) [B, L0, 768] -> (scale | shift) bf16 [B, L0, 2C]."""
        lin = self.spatial_guide[1]
        return ops.linear(feats_silu, lin.weight, lin.bias)


class WanAttentionBlock(nn.Module):
    """t4d:585-688."""

    def __init__(self, cross_attn_type, dim, ffn_dim, num_heads, window_size=(-1, -1),
                 qk_norm=True, cross_attn_norm=False, eps=1e-6, use_spatial_guidance=True,
                 device=None):
        super().__init__()
        self.dim, self.ffn_dim, self.num_heads, self.eps = dim, ffn_dim, num_heads, eps
        self.cross_attn_norm = cross_attn_norm
        self.norm1 = WanLayerNorm(dim, eps)
        self.self_attn = WanSelfAttention(dim, num_heads, window_size, qk_norm, eps, device=device)
        self.norm3 = WanLayerNorm(dim, eps, elementwise_affine=True, device=device) \
            if cross_attn_norm else None
        self.cross_attn = WanI2VCrossAttention(dim, num_heads, (-1, -1), qk_norm, eps, device=device,
                                               image_branch=(cross_attn_type == "i2v_cross_attn"))
        self.norm2 = WanLayerNorm(dim, eps)
        self.ffn = _seq(_Param((ffn_dim, dim), device=device), _Slot(),
                        _Param((dim, ffn_dim), device=device))
        self.modulation = nn.Parameter(torch.empty(1, 6, dim, device=device, dtype=BF16),
                                       requires_grad=False)
        if use_spatial_guidance:
            self.spatial_guidance_self = SpatialGuidanceModule(dim, device=device)
            self.spatial_guidance_ffn = SpatialGuidanceModule(dim, device=device)
        else:
            self.spatial_guidance_self = self.spatial_guidance_ffn = None

    def forward(self, x, e, seq_lens, grid_sizes, freqs, context, context_lens=None, dtype=BF16,
                t=0, dino_features=None, use_cls_token=False, cross_kv=None, sp=None,
                guidance_silu=None):
        """x: [B, L, C] fp32 (bf16 accepted and widened); e: [B, 6, C] fp32.  Returns the fp32
        residual stream.  A contiguous fp32 `x` is updated in place and returned."""
        _no_grad_only("WanAttentionBlock")
        if context_lens is not None:
            raise NotImplementedError("context_lens is always None on the reference path")
        if e.dim() > 3:
            raise NotImplementedError("per-token timesteps (t.dim() != 1) are not on the hot path")
        B, L, C = x.shape
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.to(torch.float32).contiguous()
        dev = x.device
        cos, sin = _rope_cache.get(freqs, dev)
        grid = torch.as_tensor(grid_sizes).to(device=dev, dtype=torch.int32).contiguous()
        k_lens = torch.as_tensor(seq_lens).to(device=dev, dtype=torch.int32)
        context = None if context is None else context.to(BF16)

        feats = guidance_silu        # bf16 SiLU(features), hoisted by the model (identical for all blocks)
        if feats is None and dino_features is not None and dino_features[0] is not None and \
                self.spatial_guidance_self is not None:
            f, cls = dino_features
            src = cls.expand(-1, f.size(1), -1) if (use_cls_token and cls is not None) else f
            feats = ops.silu_bf16(src.float().contiguous())

        em = ops.add_bcast(self.modulation, e.reshape(B, 6 * C).float()).view(B, 6, C)
        ms = 6 * C

        def adaln(shift_i, scale_i, sgm):
            sg = sgm.project(feats) if (feats is not None and sgm is not None) else None
            return ops.layernorm_modulate(x, None, None, em[:, shift_i], em[:, scale_i], ms, L,
                                          self.eps, guidance=sg,
                                          guidance_gate=None if sg is None else sgm.gate)

        # self-attention, x += y * e2
        t1 = adaln(0, 1, self.spatial_guidance_self)
        o = self.self_attn.attend(t1, k_lens, grid, cos, sin, sp)
        ops.linear(o, self.self_attn.o.weight, self.self_attn.o.bias, ops.EPI_GATE_RESIDUAL_F32,
                   out=x, residual=x, gate=em[:, 2], gate_batch_stride=ms, rows_per_batch=L)
        # cross-attention, x += y
        if self.norm3 is not None:
            n3 = ops.layernorm_modulate(x, self.norm3.weight, self.norm3.bias, eps=self.eps)
        else:
            n3 = x.to(BF16)
        o = self.cross_attn.attend(n3, context, cross_kv)
        ops.linear(o, self.cross_attn.o.weight, self.cross_attn.o.bias, ops.EPI_GATE_RESIDUAL_F32,
                   out=x, residual=x)
        # FFN, x += y * e5
        t2 = adaln(3, 4, self.spatial_guidance_ffn)
        h = ops.linear(t2, self.ffn[0].weight, self.ffn[0].bias, ops.EPI_GELU_TANH)
        ops.linear(h, self.ffn[2].weight, self.ffn[2].bias, ops.EPI_GATE_RESIDUAL_F32,
                   out=x, residual=x, gate=em[:, 5], gate_batch_stride=ms, rows_per_batch=L)
        return x


class Head(nn.Module):
    """t4d:691-721."""

    def __init__(self, dim, out_dim, patch_size, eps=1e-6, device=None):
        super().__init__()
        self.dim, self.out_dim, self.patch_size, self.eps = dim, out_dim, patch_size, eps
        self.norm = WanLayerNorm(dim, eps)
        self.head = _Param((math.prod(patch_size) * out_dim, dim), device=device)
        self.modulation = nn.Parameter(torch.empty(1, 2, dim, device=device, dtype=BF16),
                                       requires_grad=False)

    def forward(self, x, e):
        _no_grad_only("Head")
        B, L, C = x.shape
        em = ops.add_bcast(self.modulation, e.float()).view(B, 2, C)
        t = ops.layernorm_modulate(x, None, None, em[:, 0], em[:, 1], 2 * C, L, self.eps)
        return ops.linear(t, self.head.weight, self.head.bias)


class MLPProj(nn.Module):
    """t4d:724-736: LayerNorm, Linear, GELU(erf), Linear, LayerNorm over CLIP tokens."""

    def __init__(self, in_dim, out_dim, device=None):
        super().__init__()
        self.proj = _seq(WanLayerNorm(in_dim, 1e-5, True, device=device),
                         _Param((in_dim, in_dim), device=device), _Slot(),
                         _Param((out_dim, in_dim), device=device),
                         WanLayerNorm(out_dim, 1e-5, True, device=device))

    def forward(self, image_embeds):
        _no_grad_only("MLPProj")
        p = self.proj
        c = ops.layernorm_modulate(image_embeds, p[0].weight, p[0].bias, eps=1e-5)
        c = ops.linear(c, p[1].weight, p[1].bias, ops.EPI_GELU_ERF)
        c = ops.linear(c, p[3].weight, p[3].bias)
        return ops.layernorm_modulate(c, p[4].weight, p[4].bias, eps=1e-5)


class _Config(dict):
    """Minimal stand-in for diffusers' FrozenDict config (`transformer.config.patch_size`,
    `.get("add_ref_conv")` are read by the pipeline, pctl:703,737)."""
    __getattr__ = dict.get


class Conditioning:
    """Step-invariant products of `WanTransformer4DModel.precompute_conditioning`."""

    def __init__(self, context: Tensor, cross_kv):
        self.context = context            # [B, 257 + text_len, C] bf16
        self.cross_kv = cross_kv          # per block: (k, v, k_img, v_img), each [B, Lctx, C] bf16

    def tail(self, h: int) -> "Conditioning":
        """The conditional half of a CFG batch (cfg_skip, cfg_optimization.py:9-35)."""
        cut = lambda t: None if t is None else t[h:]
        return Conditioning(self.context[h:], [tuple(cut(t) for t in kv) for kv in self.cross_kv])


class WanTransformer4DModel(nn.Module):
    """t4d:785-1534 — forward-only, B200-native."""

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=2048,
                 ffn_dim=8192, freq_dim=256, text_dim=4096, out_dim=16, num_heads=16,
                 num_layers=32, window_size=(-1, -1), qk_norm=True, cross_attn_norm=True, eps=1e-6,
                 in_channels=16, hidden_size=2048, add_control_adapter=False,
                 in_dim_control_adapter=24, add_ref_conv=False, in_dim_ref_conv=16,
                 cross_attn_type=None, use_dino_guidance=False, use_omnimae_guidance=False,
                 use_depth_guidance=False, use_cls_token=False, use_spatial_guidance=None,
                 device=None):
        super().__init__()
        assert model_type in ("t2v", "i2v", "ti2v")
        if add_control_adapter:
            raise NotImplementedError("add_control_adapter: SimpleAdapter is undefined in the "
                                      "reference as well (SURVEY.md F4)")
        self.config = _Config(model_type=model_type, patch_size=tuple(patch_size), text_len=text_len,
                              in_dim=in_dim, dim=dim, ffn_dim=ffn_dim, freq_dim=freq_dim,
                              text_dim=text_dim, out_dim=out_dim, num_heads=num_heads,
                              num_layers=num_layers, add_ref_conv=add_ref_conv,
                              in_dim_ref_conv=in_dim_ref_conv, eps=eps)
        self.model_type, self.patch_size, self.text_len = model_type, tuple(patch_size), text_len
        self.in_dim, self.dim, self.ffn_dim, self.freq_dim = in_dim, dim, ffn_dim, freq_dim
        self.text_dim, self.out_dim, self.num_heads, self.num_layers = text_dim, out_dim, num_heads, num_layers
        self.eps, self.use_cls_token = eps, use_cls_token
        if use_spatial_guidance is None:
            use_spatial_guidance = bool(use_dino_guidance or use_omnimae_guidance)
        pt, ph, pw = self.patch_size
        if (pt, ph, pw) != (1, 2, 2):
            raise NotImplementedError("patch_size must be (1, 2, 2)")
        self.patch_embedding = _Param((dim, in_dim, pt, ph, pw), device=device)
        self.text_embedding = _seq(_Param((dim, text_dim), device=device), _Slot(),
                                   _Param((dim, dim), device=device))
        self.time_embedding = _seq(_Param((dim, freq_dim), device=device), _Slot(),
                                   _Param((dim, dim), device=device))
        self.time_projection = _seq(_Slot(), _Param((dim * 6, dim), device=device))
        if cross_attn_type is None:
            cross_attn_type = "t2v_cross_attn" if model_type == "t2v" else "i2v_cross_attn"
        self.blocks = nn.ModuleList([
            WanAttentionBlock(cross_attn_type, dim, ffn_dim, num_heads, window_size, qk_norm,
                              cross_attn_norm, eps, use_spatial_guidance=use_spatial_guidance,
                              device=device) for _ in range(num_layers)])
        self.head = Head(dim, out_dim, self.patch_size, eps, device=device)
        d = dim // num_heads
        assert dim % num_heads == 0 and d % 2 == 0
        self.d = d
        self.freqs = build_freqs(d)                      # complex128, like the reference (F8)
        if model_type == "i2v":
            self.img_emb = MLPProj(1280, dim, device=device)
        self.ref_conv = _Param((dim, in_dim_ref_conv, ph, pw), device=device) if add_ref_conv else None
        # Motion-Perception front end (t4d:883-892): the frozen OmniMAE trunk is out of scope — any
        # module with `.trunk.forward_patch_features(img, None) -> ([1, 196, 768], cls)` can be
        # attached as `omnimae_extractor` (install() attaches the reference's own); the trainable
        # feature_adapter, the bilinear resize and the temporal repeat run on the kernels here.
        self.use_omnimae_guidance = bool(use_omnimae_guidance)
        self.omnimae_extractor = None
        self.dino_dim = 768
        if use_omnimae_guidance:
            self.feature_adapter = _seq(_Param((768, 768, 3, 3), device=device), _Slot(),
                                        _Param((768, 768, 3, 3), device=device))
        self._adapter_packed = {}
        self.control_adapter = None
        self.teacache = None
        self.cfg_skip_ratio = None
        self.current_steps = 0
        self.num_inference_steps = None
        self.sp_world_size, self.sp_world_rank = 1, 0
        self.sp = None                # dist.SequenceParallel after enable_multi_gpus_inference()

    @classmethod
    def from_config(cls, cfg: DiTConfig, device=None) -> "WanTransformer4DModel":
        return cls(model_type=cfg.model_type, patch_size=cfg.patch_size, text_len=cfg.text_len,
                   in_dim=cfg.in_dim, dim=cfg.dim, ffn_dim=cfg.ffn_dim, freq_dim=cfg.freq_dim,
                   text_dim=cfg.text_dim, out_dim=cfg.out_dim, num_heads=cfg.num_heads,
                   num_layers=cfg.num_layers, qk_norm=cfg.qk_norm,
                   cross_attn_norm=cfg.cross_attn_norm, eps=cfg.eps, add_ref_conv=cfg.add_ref_conv,
                   in_dim_ref_conv=cfg.in_dim_ref_conv, use_omnimae_guidance=cfg.use_omnimae_guidance,
                   use_spatial_guidance=cfg.use_spatial_guidance or cfg.use_omnimae_guidance, device=device)

    @classmethod
    def from_pretrained(cls, pretrained_model_path, subfolder=None, transformer_additional_kwargs=None,
                        low_cpu_mem_usage=False, torch_dtype=BF16, device=None):
        """Checkpoint loading with the reference's rules (t4d:1393-1534; callers infer.py:537-565):
        `config.json` supplies the constructor arguments (overridden / extended by
        `transformer_additional_kwargs`, incl. its `dict_mapping` renames); weights come from
        `diffusion_pytorch_model.bin|.safetensors` or every `*.safetensors` shard in the folder; a
        `patch_embedding.weight` with fewer input channels than the model (the released 48-channel
        Control checkpoint vs the 64-channel depth-conditioned model) is zero-padded; tensors whose
        shape still differs are skipped; loading is non-strict.  `low_cpu_mem_usage` is accepted for
        signature compatibility (parameters are created directly on `device`)."""
        import glob
        import inspect
        import json
        import os
        if torch_dtype != BF16:
            raise NotImplementedError("the B200 kernels run the reference's bf16 inference path only")
        path = pretrained_model_path if subfolder is None else os.path.join(pretrained_model_path, subfolder)
        cfg_file = os.path.join(path, "config.json")
        if not os.path.isfile(cfg_file):
            raise RuntimeError(f"{cfg_file} does not exist")
        with open(cfg_file) as f:
            config = json.load(f)
        extra = dict(transformer_additional_kwargs or {})
        for src, dst in extra.pop("dict_mapping", {}).items():
            extra[dst] = config[src]
        accepted = set(inspect.signature(cls.__init__).parameters) - {"self", "device"}
        kwargs = {k: v for k, v in {**config, **extra}.items() if k in accepted}
        model = cls(**kwargs, device=device)
        files = [os.path.join(path, "diffusion_pytorch_model.bin"),
                 os.path.join(path, "diffusion_pytorch_model.safetensors")]
        if os.path.exists(files[0]):
            state = torch.load(files[0], map_location="cpu")
        else:
            from safetensors.torch import load_file
            shards = [files[1]] if os.path.exists(files[1]) else sorted(glob.glob(os.path.join(path, "*.safetensors")))
            if not shards:
                raise RuntimeError(f"no weights found under {path}")
            state = {}
            for shard in shards:
                state.update(load_file(shard))
        own = model.state_dict()
        pe = state.get("patch_embedding.weight")
        if pe is not None and pe.shape != own["patch_embedding.weight"].shape and \
                pe.shape[1] < own["patch_embedding.weight"].shape[1] and pe.shape[0] == own["patch_embedding.weight"].shape[0]:
            wide = torch.zeros(own["patch_embedding.weight"].shape, dtype=pe.dtype)
            wide[:, :pe.shape[1]] = pe
            state["patch_embedding.weight"] = wide
        state = {k: v for k, v in state.items() if k in own and own[k].shape == v.shape}
        missing, unexpected = model.load_state_dict(state, strict=False)
        model.load_report = {"missing": list(missing), "unexpected": list(unexpected)}
        return model

    @property
    def dtype(self):
        return self.patch_embedding.weight.dtype

    # cfg_skip bookkeeping of the reference (t4d:986-1008, cfg_optimization.py:5-39)
    def enable_riflex(self, k=6, L_test=66, L_test_scale=4.886):
        """t4d:1011-1025: swap the frame-axis RoPE table for its RIFLEx variant (length extrapolation).
        The kernels read whatever table `self.freqs` holds."""
        d = self.d
        self.freqs = torch.cat([rope_params_riflex(1024, d - 4 * (d // 6), k=k, L_test=L_test,
                                                   L_test_scale=L_test_scale),
                                rope_params(1024, 2 * (d // 6)), rope_params(1024, 2 * (d // 6))], dim=1)

    def disable_riflex(self):
        """t4d:1027-1036."""
        self.freqs = build_freqs(self.d)

    def enable_multi_gpus_inference(self, group=None):
        """t4d:1038-1044: shard ONE sample's sequence over the ranks of `group` (Ulysses).  Needs an
        initialised torch.distributed process group; num_heads must be a multiple of its size."""
        from .dist import SequenceParallel
        sp = SequenceParallel(group)
        if self.num_heads % sp.world:
            raise ValueError(f"{self.num_heads} heads cannot be split over {sp.world} ranks")
        self.sp, self.sp_world_size, self.sp_world_rank = sp, sp.world, sp.rank

    def disable_multi_gpus_inference(self):
        self.sp, self.sp_world_size, self.sp_world_rank = None, 1, 0

    def enable_cfg_skip(self, cfg_skip_ratio, num_steps):
        if cfg_skip_ratio != 0:
            self.cfg_skip_ratio, self.current_steps, self.num_inference_steps = cfg_skip_ratio, 0, num_steps
        else:
            self.disable_cfg_skip()

    def disable_cfg_skip(self):
        self.cfg_skip_ratio, self.current_steps, self.num_inference_steps = None, 0, None

    # TeaCache hooks of the reference (t4d:961-978)
    def enable_teacache(self, coefficients, num_steps: int, rel_l1_thresh: float,
                        num_skip_start_steps: int = 0, offload: bool = True):
        from .cache_utils import TeaCache
        self.teacache = TeaCache(coefficients, num_steps, rel_l1_thresh=rel_l1_thresh,
                                 num_skip_start_steps=num_skip_start_steps, offload=offload)

    def share_teacache(self, transformer=None):
        self.teacache = transformer.teacache

    def disable_teacache(self):
        self.teacache = None

    # ---------------------------------------------------------------------------------
    def embed_time(self, t: Tensor):
        """e [B, C], e0 [B, 6, C] in fp32 (t4d:1160-1171)."""
        s = ops.timestep_embedding(t, self.freq_dim)
        te = self.time_embedding
        h = ops.small_linear_f32(s, te[0].weight, te[0].bias, silu_out=True)
        e = ops.small_linear_f32(h, te[2].weight, te[2].bias)
        tp = self.time_projection[1]
        e0 = ops.small_linear_f32(e, tp.weight, tp.bias, silu_in=True)
        return e, e0.view(-1, 6, self.dim)

    def embed_context(self, context: Sequence[Tensor], clip_fea: Optional[Tensor]) -> Tensor:
        """[B, 257 + text_len, C] bf16 (t4d:1175-1184).  Step-invariant."""
        dev = self.patch_embedding.weight.device
        B = len(context)
        ctx = torch.zeros(B, self.text_len, self.text_dim, device=dev, dtype=BF16)
        for i, u in enumerate(context):
            ctx[i, :u.size(0)].copy_(u)
        te = self.text_embedding
        h = ops.linear(ctx, te[0].weight, te[0].bias, ops.EPI_GELU_TANH)
        ctx = ops.linear(h, te[2].weight, te[2].bias)
        if clip_fea is not None and self.model_type == "i2v":
            ctx = torch.cat([self.img_emb(clip_fea.to(device=dev, dtype=BF16)), ctx], dim=1)
        return ctx

    def precompute_conditioning(self, context: Sequence[Tensor], clip_fea: Optional[Tensor]) -> "Conditioning":
        """Everything in the forward that depends only on the prompt / CLIP tokens: the context
        embedding (t4d:1175-1184) and every block's cross-attention K/V (t4d:481-483,527-531).
        The reference recomputes them on each of the 50 steps (pctl:796 calls the whole forward);
        they are bit-identical across steps, so `forward(..., conditioning=c)` reuses them
        (SURVEY §8f rank 1).  ~31 MB per block at 14B dims, batch 2."""
        _no_grad_only("WanTransformer4DModel")
        dev = self.patch_embedding.weight.device
        ctx = self.embed_context([c.to(device=dev, dtype=BF16) for c in context], clip_fea)
        return Conditioning(ctx, [blk.cross_attn.project_context(ctx) for blk in self.blocks])

    def forward(self, x, t, context, seq_len, clip_fea=None, y=None, y_camera=None, full_ref=None,
                subject_ref=None, cond_flag=True, first_frame=None, guidance_features=None,
                conditioning: Optional["Conditioning"] = None):
        """Same call surface as t4d:1046-1060.  x: [B, 16, T, h, w] (tensor or list of
        [16, T, h, w]); y: [B, 48, T, h, w]; t: [B]; context: list of [Lt, text_dim];
        clip_fea: [B, 257, 1280]; full_ref: [B, 16, h, w].  Returns [B, 16, T, h, w] bf16.

        `first_frame` (the OmniMAE front end, t4d:1127-1156) is outside the hot path; its
        *output* can be supplied as guidance_features=(feats [B, L0, 768], cls [B, 1, 768])."""
        _no_grad_only("WanTransformer4DModel")
        if subject_ref is not None or y_camera is not None:
            raise NotImplementedError("subject_ref / y_camera are not used by the 4D-STraG path")
        if first_frame is not None and guidance_features is not None:
            raise ValueError("pass either first_frame or guidance_features, not both")
        # @cfg_skip() wrapper semantics (cfg_optimization.py:5-39)
        bs = len(x)
        skip = (bs >= 2 and self.cfg_skip_ratio is not None and
                self.current_steps >= self.num_inference_steps * (1 - self.cfg_skip_ratio))
        if skip:
            h = bs // 2
            if conditioning is not None:
                conditioning = conditioning.tail(h)
            x, t, context = x[h:], t[h:], context[h:]
            clip_fea = None if clip_fea is None else clip_fea[h:]
            y = None if y is None else y[h:]
            full_ref = None if full_ref is None else full_ref[h:]
            # the reference's wrapper slices EVERY tensor / tuple kwarg (cfg_optimization.py:18-25)
            first_frame = None if first_frame is None else first_frame[h:]
            if guidance_features is not None:
                guidance_features = tuple(None if g is None else g[h:] for g in guidance_features)
        out = self._forward(x, t, context, seq_len, clip_fea, y, full_ref, guidance_features, cond_flag,
                            conditioning, first_frame)
        if skip:
            out = torch.cat([out, out], dim=0)
        return out

    def motion_perception_features(self, first_frame: Tensor, target_hw, latent_T: int):
        """t4d:1127-1156: first_frame [B, 3, H, W] in [0, 1] -> ((features [B, T*h*w, 768] bf16,
        cls [B, 1, 768]), bf16 SiLU(features)).  ImageNet normalisation + the frozen OmniMAE trunk
        run as the attached torch module (once per forward, out of scope); feature_adapter
        (im2col + tcgen05 GEMM, SiLU fused into the second gather), bilinear resize and the
        temporal repeat are kernels."""
        if not self.use_omnimae_guidance or self.omnimae_extractor is None:
            raise RuntimeError("first_frame needs use_omnimae_guidance=True and an attached "
                               "`omnimae_extractor` (the OmniMAE trunk is not part of more4d_b200)")
        dev = self.patch_embedding.weight.device
        ff = first_frame.to(dev)
        mean = torch.tensor([0.485, 0.456, 0.406], device=dev, dtype=ff.dtype).view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225], device=dev, dtype=ff.dtype).view(1, 3, 1, 1)
        ff = (ff - mean) / std                                            # transforms.Normalize, t4d:1134-1135
        feats, cls = [], []
        for i in range(ff.shape[0]):                                      # per-sample, like t4d:1138-1145
            f, c = self.omnimae_extractor.trunk.forward_patch_features(ff[i:i + 1], None)
            feats.append(f)
            cls.append(c)
        feats, cls = torch.cat(feats), torch.cat(cls)
        B = ff.shape[0]
        tok = feats.reshape(B, 14, 14, self.dino_dim).to(BF16).contiguous()   # channels-last already
        fa = self.feature_adapter
        h = tok
        for idx, silu in ((0, False), (2, True)):
            w = fa[idx].weight
            key = (idx, w.data_ptr(), w._version)
            if self._adapter_packed.get(idx, (None,))[0] != key:
                self._adapter_packed[idx] = (key, ops.pack_conv_weight(w))
            rows = ops.im2col3x3_cl(h, silu=silu)
            h = ops.linear(rows, self._adapter_packed[idx][1], fa[idx].bias).view(B, 14, 14, self.dino_dim)
        raw, act = ops.bilinear_repeat_cl(h, latent_T, int(target_hw[0]), int(target_hw[1]))
        return (raw, cls.reshape(B, 1, self.dino_dim)), act

    def _forward(self, x, t, context, seq_len, clip_fea, y, full_ref, guidance_features, cond_flag=True,
                 conditioning=None, first_frame=None):
        dev = self.patch_embedding.weight.device
        if isinstance(x, (list, tuple)):
            x = torch.stack(list(x))
        if isinstance(y, (list, tuple)):
            y = torch.stack(list(y))
        x = x.to(device=dev, dtype=BF16)
        y = None if y is None else y.to(device=dev, dtype=BF16)
        B, _, T, H, W = x.shape
        C = self.dim
        grid = [T, H // 2, W // 2]
        L0 = grid[0] * grid[1] * grid[2]
        ref_len = 0
        if self.ref_conv is not None and full_ref is not None:
            ref_len = grid[1] * grid[2]
            grid[0] += 1
            seq_len = seq_len + ref_len
        n_tok = ref_len + L0
        if self.sp_world_size > 1:
            seq_len = int(math.ceil(seq_len / self.sp_world_size)) * self.sp_world_size
        assert n_tok <= seq_len                                   # t4d:1102
        L = seq_len
        # residual stream, fp32 from the start (exact widening of the bf16 embeddings; F7)
        xs = torch.zeros(B, L, C, device=dev, dtype=torch.float32) if L > n_tok else \
            torch.empty(B, L, C, device=dev, dtype=torch.float32)
        cols = ops.patchify(x, y)                                  # [B, L0, in_dim*4]
        pw_ = self.patch_embedding.weight.view(C, -1)
        for b in range(B):
            ops.linear(cols[b], pw_, self.patch_embedding.bias, ops.EPI_F32,
                       out=xs[b, ref_len:ref_len + L0])
        if ref_len:
            rcols = ops.patchify(full_ref.to(device=dev, dtype=BF16).unsqueeze(2))
            rw = self.ref_conv.weight.view(C, -1)
            for b in range(B):
                ops.linear(rcols[b], rw, self.ref_conv.bias, ops.EPI_F32, out=xs[b, :ref_len])
        e, e0 = self.embed_time(t.to(dev))
        if conditioning is not None:
            ctx = None
        else:
            ctx = self.embed_context([c.to(device=dev, dtype=BF16) for c in context], clip_fea)
        seq_lens = torch.full((B,), n_tok, device=dev, dtype=torch.int32)
        grid_sizes = torch.tensor([grid] * B, device=dev, dtype=torch.int32)
        guidance_silu = None
        if first_frame is not None:                                # t4d:1127-1156
            guidance_features, guidance_silu = self.motion_perception_features(first_frame, grid[1:], T)
            if self.use_cls_token:
                guidance_silu = None                               # blocks then project the CLS token instead
        sp = self.sp if self.sp_world_size > 1 else None
        if sp is not None:                                         # context parallel, t4d:1187-1198
            if guidance_features is not None:
                raise NotImplementedError("Motion-Perception guidance under sequence parallelism")
            xs = sp.shard_tokens(xs)
        run_blocks = True
        tc = self.teacache
        if tc is not None:                                         # t4d:1200-1270
            run_blocks = teacache_decide(tc, e0, cond_flag)
            if not run_blocks:
                prev = tc.previous_residual_cond if cond_flag else tc.previous_residual_uncond
                xs = xs + prev.to(xs.device)[-xs.size(0):]         # cache hit: block stack skipped
            else:
                ori = xs.clone().cpu() if tc.offload else xs.clone()
        if run_blocks:
            for i, blk in enumerate(self.blocks):
                xs = blk(xs, e0, seq_lens, grid_sizes, self.freqs, ctx, None, BF16, t,
                         dino_features=guidance_features, use_cls_token=self.use_cls_token,
                         cross_kv=None if conditioning is None else conditioning.cross_kv[i], sp=sp,
                         guidance_silu=guidance_silu)
            if tc is not None:
                res = (xs.cpu() - ori) if tc.offload else (xs - ori)
                if cond_flag:
                    tc.previous_residual_cond = res
                else:
                    tc.previous_residual_uncond = res
        tok = self.head(xs, e)                                     # [B, L, 64] bf16
        if sp is not None:
            tok = sp.gather_tokens(tok)                            # t4d:1320-1321
        out = ops.unpatchify(tok, ref_len, self.out_dim, T, H, W)
        if tc is not None:
            teacache_step_done(tc, cond_flag)
        return out


class WanTransformer3DModel(WanTransformer4DModel):
    """MoRe4D/models/wan_transformer3d.py:725-1511 — the Wan2.1-Fun-InP backbone of the 4D-ViSM
    stage (scripts/inference/infer.py:935-994).  File-level diff against wan_transformer4d.py:
    the same blocks, embeddings, RoPE, head and state-dict keys, without the Motion-Perception
    branch (no `spatial_guidance_*` parameters, no `first_frame`); the oracle reproduces the real
    class to 2e-7 (tests/test_oracle_vs_golden.py).  Same constructor arguments as t3d:735-759."""

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=2048,
                 ffn_dim=8192, freq_dim=256, text_dim=4096, out_dim=16, num_heads=16,
                 num_layers=32, window_size=(-1, -1), qk_norm=True, cross_attn_norm=True, eps=1e-6,
                 in_channels=16, hidden_size=2048, add_control_adapter=False,
                 in_dim_control_adapter=24, add_ref_conv=False, in_dim_ref_conv=16,
                 cross_attn_type=None, device=None):
        super().__init__(model_type=model_type, patch_size=patch_size, text_len=text_len, in_dim=in_dim,
                         dim=dim, ffn_dim=ffn_dim, freq_dim=freq_dim, text_dim=text_dim, out_dim=out_dim,
                         num_heads=num_heads, num_layers=num_layers, window_size=window_size,
                         qk_norm=qk_norm, cross_attn_norm=cross_attn_norm, eps=eps,
                         in_channels=in_channels, hidden_size=hidden_size,
                         add_control_adapter=add_control_adapter,
                         in_dim_control_adapter=in_dim_control_adapter, add_ref_conv=add_ref_conv,
                         in_dim_ref_conv=in_dim_ref_conv, cross_attn_type=cross_attn_type,
                         use_spatial_guidance=False, device=device)

    @classmethod
    def from_config(cls, cfg: DiTConfig, device=None) -> "WanTransformer3DModel":
        return cls(model_type=cfg.model_type, patch_size=cfg.patch_size, text_len=cfg.text_len,
                   in_dim=cfg.in_dim, dim=cfg.dim, ffn_dim=cfg.ffn_dim, freq_dim=cfg.freq_dim,
                   text_dim=cfg.text_dim, out_dim=cfg.out_dim, num_heads=cfg.num_heads,
                   num_layers=cfg.num_layers, qk_norm=cfg.qk_norm,
                   cross_attn_norm=cfg.cross_attn_norm, eps=cfg.eps, add_ref_conv=cfg.add_ref_conv,
                   in_dim_ref_conv=cfg.in_dim_ref_conv, device=device)

    def forward(self, x, t, context, seq_len, clip_fea=None, y=None, y_camera=None, full_ref=None,
                subject_ref=None, cond_flag=True, conditioning=None):                 # t3d:966-978
        return super().forward(x, t, context, seq_len, clip_fea=clip_fea, y=y, y_camera=y_camera,
                               full_ref=full_ref, subject_ref=subject_ref, cond_flag=cond_flag,
                               conditioning=conditioning)

